// encode_mma_kernel instantiations (own translation unit: compile time).
#include "vibo_stream_kernel.cuh"

namespace vibo {

namespace {
template <int D, int KS, bool EVEN>
cudaError_t run(int grid, size_t smem, const StreamParams& p, int missing_policy, const float* table, float* mu,
                float* lv, float* S, cudaStream_t st) {
  auto k = encode_mma_kernel<D, KS, EVEN>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<grid, kMmaWarps * 32, smem, st>>>(p, missing_policy, table, mu, lv, S);
  return cudaGetLastError();
}
template <int D, int KS>
cudaError_t run_e(bool even, int grid, size_t smem, const StreamParams& p, int missing_policy, const float* table,
                  float* mu, float* lv, float* S, cudaStream_t st) {
  return even ? run<D, KS, true>(grid, smem, p, missing_policy, table, mu, lv, S, st)
              : run<D, KS, false>(grid, smem, p, missing_policy, table, mu, lv, S, st);
}
template <int D>
cudaError_t run_d(int KS, bool even, int grid, size_t smem, const StreamParams& p, int missing_policy,
                  const float* table, float* mu, float* lv, float* S, cudaStream_t st) {
  switch (KS) {
    case 1: return run_e<D, 1>(even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 2: return run_e<D, 2>(even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 4: return run_e<D, 4>(even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 8: return run_e<D, 8>(even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace

cudaError_t stream_encode_mma_run(int D, int KS, bool even, int grid, size_t smem, const StreamParams& p,
                                  int missing_policy, const float* table, float* mu, float* lv, float* S,
                                  cudaStream_t st) {
  switch (D) {
    case 1: return run_d<1>(KS, even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 2: return run_d<2>(KS, even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 3: return run_d<3>(KS, even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 4: return run_d<4>(KS, even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 5: return run_d<5>(KS, even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 6: return run_d<6>(KS, even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 7: return run_d<7>(KS, even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    case 8: return run_d<8>(KS, even, grid, smem, p, missing_policy, table, mu, lv, S, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace vibo
