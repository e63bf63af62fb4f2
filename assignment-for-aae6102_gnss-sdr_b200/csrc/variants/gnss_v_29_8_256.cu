// One engine variant per translation unit (they compile in parallel): Q = 29, 8 CTAs x 256 threads per transform.
#include "../gnss_kernels.cuh"
namespace gnss {
extern const VariantOps gnss_variant_29_8_256 = Variant<29, 8, 256, 2>::ops();
}  // namespace gnss
