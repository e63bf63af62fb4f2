// Engine variants for N = 26 000 = 16*125*13 (Urban front end, 26 MHz).
#include "gnss_kernels.cuh"
namespace gnss {
const VariantOps* gnss_variants_q13(int* count) {
    static const VariantOps v[] = {
        Variant<13, 8, 128, 4>::ops(),   // default (first match): fastest measured, profiles/r01
        Variant<13, 2, 512, 1>::ops(),
        Variant<13, 4, 256, 2>::ops(),
        Variant<13, 4, 512, 1>::ops(),
    };
    *count = (int)(sizeof(v) / sizeof(v[0]));
    return v;
}
}  // namespace gnss
