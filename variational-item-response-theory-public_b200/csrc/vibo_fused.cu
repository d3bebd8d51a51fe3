// Single-pass fused ELBO kernel (placeholder until the TMA pipeline lands).
#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

bool fused_supported(const vibo_desc&, const float*, const uint8_t*) { return false; }
size_t fused_workspace_bytes(const vibo_desc&) { return 0; }
cudaError_t launch_fused(const vibo_desc&, const float*, const uint8_t*, const float*, const float*,
                         const float*, uint64_t, float, double*, float*, float*, float*, float*, float*,
                         void*, size_t, bool, cudaStream_t) {
  return cudaErrorNotSupported;
}

}  // namespace vibo
