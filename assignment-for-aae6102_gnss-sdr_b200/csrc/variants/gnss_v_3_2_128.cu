// One engine variant per translation unit (they compile in parallel): Q = 3, 2 CTAs x 128 threads per transform.
#include "../gnss_kernels.cuh"
namespace gnss {
extern const VariantOps gnss_variant_3_2_128 = Variant<3, 2, 128, 4>::ops();
}  // namespace gnss
