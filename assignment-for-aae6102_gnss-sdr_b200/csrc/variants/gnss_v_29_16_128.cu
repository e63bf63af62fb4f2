// One engine variant per translation unit (they compile in parallel): Q = 29, 16 CTAs x 128 threads per transform.
#include "../gnss_kernels.cuh"
namespace gnss {
extern const VariantOps gnss_variant_29_16_128 = Variant<29, 16, 128, 4>::ops();
}  // namespace gnss
