// fused3_kernel (item-owner training kernel) instantiations.
#include "vibo_fused3_kernel.cuh"
namespace vibo {
VIBO_FUSED3_INSTANTIATE(1, 1)
VIBO_FUSED3_INSTANTIATE(2, 1)
VIBO_FUSED3_INSTANTIATE(2, 2)
}  // namespace vibo
