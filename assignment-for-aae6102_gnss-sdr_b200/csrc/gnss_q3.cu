// Engine variants for N = 6 000 = 16*125*3 (small test front end, 6 MHz).  The first entry is the default
// (fastest measured, profiles/); each variant is its own translation unit under variants/.
#include "gnss_internal.h"
namespace gnss {
extern const VariantOps gnss_variant_3_2_128;
extern const VariantOps gnss_variant_3_1_256;
const VariantOps* gnss_variants_q3(int* count) {
    static const VariantOps* const p[] = {&gnss_variant_3_2_128, &gnss_variant_3_1_256};
    static VariantOps v[sizeof(p) / sizeof(p[0])];
    *count = (int)(sizeof(p) / sizeof(p[0]));
    for (int i = 0; i < *count; ++i) v[i] = *p[i];
    return v;
}
}  // namespace gnss
