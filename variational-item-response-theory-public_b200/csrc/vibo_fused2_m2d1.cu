// fused2_kernel (two-phase) instantiations for the 2PL model, ability_dim 1.
#include "vibo_fused2_kernel.cuh"
namespace vibo {
VIBO_FUSED2_INSTANTIATE(2, 1)
}  // namespace vibo
