// One engine variant per translation unit (they compile in parallel): Q = 13, 8 CTAs x 128 threads per transform.
#include "../gnss_kernels.cuh"
namespace gnss {
extern const VariantOps gnss_variant_13_8_128 = Variant<13, 8, 128, 4>::ops();
}  // namespace gnss
