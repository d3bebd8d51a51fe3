// encode_bwd_mma_kernel instantiations (own translation unit: compile time).
#include "vibo_stream_kernel.cuh"

namespace vibo {

namespace {
template <int D, int MT>
cudaError_t run(int grid, size_t smem, const StreamParams& p, float* part, cudaStream_t st) {
  if constexpr (MT * ((4 * D + 7) / 8) * 8 <= 96) {   // accumulator registers (mirrors bwd_mma_plan)
    auto k = encode_bwd_mma_kernel<D, MT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, 512, smem, st>>>(p, part);
    return cudaGetLastError();
  } else {
    return cudaErrorInvalidValue;
  }
}
template <int D>
cudaError_t run_d(int MT, int grid, size_t smem, const StreamParams& p, float* part, cudaStream_t st) {
  switch (MT) {
    case 1: return run<D, 1>(grid, smem, p, part, st);
    case 2: return run<D, 2>(grid, smem, p, part, st);
    case 4: return run<D, 4>(grid, smem, p, part, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace

cudaError_t stream_encode_bwd_mma_run(int D, int MT, int grid, size_t smem, const StreamParams& p, float* part,
                                      cudaStream_t st) {
  switch (D) {
    case 1: return run_d<1>(MT, grid, smem, p, part, st);
    case 2: return run_d<2>(MT, grid, smem, p, part, st);
    case 3: return run_d<3>(MT, grid, smem, p, part, st);
    case 4: return run_d<4>(MT, grid, smem, p, part, st);
    case 5: return run_d<5>(MT, grid, smem, p, part, st);
    case 6: return run_d<6>(MT, grid, smem, p, part, st);
    case 7: return run_d<7>(MT, grid, smem, p, part, st);
    case 8: return run_d<8>(MT, grid, smem, p, part, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace vibo
