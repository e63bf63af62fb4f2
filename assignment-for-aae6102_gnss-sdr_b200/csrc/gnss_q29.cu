// Engine variants for N = 58 000 = 16*125*29 (Opensky front end, initParameters.m:42,46).
#include "gnss_kernels.cuh"
namespace gnss {
const VariantOps* gnss_variants_q29(int* count) {
    static const VariantOps v[] = {
        Variant<29, 16, 128, 4>::ops(),  // default (first match): fastest measured, profiles/r01
        Variant<29, 8, 256, 2>::ops(),
        Variant<29, 4, 512, 1>::ops(),
    };
    *count = (int)(sizeof(v) / sizeof(v[0]));
    return v;
}
}  // namespace gnss
