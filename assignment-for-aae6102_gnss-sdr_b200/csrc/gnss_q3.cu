// Engine variants for N = 6 000 = 16*125*3 (small test front end, 6 MHz).
#include "gnss_kernels.cuh"
namespace gnss {
const VariantOps* gnss_variants_q3(int* count) {
    static const VariantOps v[] = {
        Variant<3, 2, 128, 4>::ops(),
        Variant<3, 1, 256, 2>::ops(),
    };
    *count = (int)(sizeof(v) / sizeof(v[0]));
    return v;
}
}  // namespace gnss
