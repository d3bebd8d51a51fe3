// fused_uncond_kernel instantiations for the 2PL model, ability_dim 2.
#include "vibo_fused_kernel.cuh"
namespace vibo {
VIBO_FUSED_INSTANTIATE(2, 2)
}  // namespace vibo
