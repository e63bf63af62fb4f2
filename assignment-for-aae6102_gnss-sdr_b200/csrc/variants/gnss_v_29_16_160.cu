// One engine variant per translation unit: Q = 29, 16 CTAs x 160 threads per transform, 3 CTAs per SM.  Five warps fit the
// task counts of an a-row (125 / 145 / 115 tasks of passes 1 / 3 / 4, 25 pass-2 columns = 5 rounds of 5) in ONE round each;
// the price is 27 resident CTA groups instead of 37.  Measured, not the default (profiles/r02/ab_t160_v16.txt).
#include "../gnss_kernels.cuh"
namespace gnss {
extern const VariantOps gnss_variant_29_16_160 = Variant<29, 16, 160, 3>::ops();
}  // namespace gnss
