// fused_uncond_kernel instantiations for the 3PL model, ability_dim 1.
#include "vibo_fused_kernel.cuh"
namespace vibo {
VIBO_FUSED_INSTANTIATE(3, 1)
}  // namespace vibo
