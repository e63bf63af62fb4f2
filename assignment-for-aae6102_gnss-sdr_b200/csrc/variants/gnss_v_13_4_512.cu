// One engine variant per translation unit (they compile in parallel): Q = 13, 4 CTAs x 512 threads per transform.
#include "../gnss_kernels.cuh"
namespace gnss {
extern const VariantOps gnss_variant_13_4_512 = Variant<13, 4, 512, 1>::ops();
}  // namespace gnss
