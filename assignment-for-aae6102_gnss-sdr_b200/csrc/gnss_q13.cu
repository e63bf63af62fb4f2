// Engine variants for N = 26 000 = 16*125*13 (Urban front end, 26 MHz).  The first entry is the default
// (fastest measured, profiles/); each variant is its own translation unit under variants/.
#include "gnss_internal.h"
namespace gnss {
extern const VariantOps gnss_variant_13_8_128;
extern const VariantOps gnss_variant_13_2_512;
extern const VariantOps gnss_variant_13_4_256;
extern const VariantOps gnss_variant_13_4_512;
const VariantOps* gnss_variants_q13(int* count) {
    static const VariantOps* const p[] = {&gnss_variant_13_8_128, &gnss_variant_13_2_512, &gnss_variant_13_4_256, &gnss_variant_13_4_512};
    static VariantOps v[sizeof(p) / sizeof(p[0])];
    *count = (int)(sizeof(p) / sizeof(p[0]));
    for (int i = 0; i < *count; ++i) v[i] = *p[i];
    return v;
}
}  // namespace gnss
