// One engine variant per translation unit (they compile in parallel): Q = 13, 4 CTAs x 256 threads per transform.
#include "../gnss_kernels.cuh"
namespace gnss {
extern const VariantOps gnss_variant_13_4_256 = Variant<13, 4, 256, 2>::ops();
}  // namespace gnss
