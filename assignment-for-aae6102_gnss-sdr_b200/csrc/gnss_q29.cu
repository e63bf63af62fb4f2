// Engine variants for N = 58 000 = 16*125*29 (Opensky front end, initParameters.m:42,46).  The first entry is the default
// (fastest measured, profiles/); each variant is its own translation unit under variants/.
#include "gnss_internal.h"
namespace gnss {
extern const VariantOps gnss_variant_29_16_128;
extern const VariantOps gnss_variant_29_8_256;
extern const VariantOps gnss_variant_29_4_512;
const VariantOps* gnss_variants_q29(int* count) {
    static const VariantOps* const p[] = {&gnss_variant_29_16_128, &gnss_variant_29_8_256, &gnss_variant_29_4_512};
    static VariantOps v[sizeof(p) / sizeof(p[0])];
    *count = (int)(sizeof(p) / sizeof(p[0]));
    for (int i = 0; i < *count; ++i) v[i] = *p[i];
    return v;
}
}  // namespace gnss
