// One engine variant per translation unit (they compile in parallel): Q = 3, 1 CTAs x 256 threads per transform.
#include "../gnss_kernels.cuh"
namespace gnss {
extern const VariantOps gnss_variant_3_1_256 = Variant<3, 1, 256, 2>::ops();
}  // namespace gnss
