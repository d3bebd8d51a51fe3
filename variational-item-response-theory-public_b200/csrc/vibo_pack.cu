// Packed row format: ONE signed byte per response cell,
//   -1 = missing (the reference's MISSING_DATA marker, src/config.py:14), 0 / 1 = observed response,
// i.e. the (response, mask) pair of src/datasets.py:928-940 (4 + 1 bytes per cell) in 1 byte.
// It is the transfer format of the host-buffer entry (vibo_fused_elbo_host_packed: PCIe carries
// 1 B/cell instead of 5) and a compact resident format; the row kernels read the unpacked pair.
#include "vibo_kernels.h"

namespace vibo {

__global__ void __launch_bounds__(256) pack_kernel(int64_t n, const float* __restrict__ resp,
                                                   const uint8_t* __restrict__ mask, int8_t* __restrict__ out) {
  const int64_t n4 = n >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(resp) & 15) | (reinterpret_cast<uintptr_t>(mask) & 3) |
                    (reinterpret_cast<uintptr_t>(out) & 3)) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    for (int64_t k = t0; k < n4; k += stride) {
      const float4 x = reinterpret_cast<const float4*>(resp)[k];
      const uchar4 m = reinterpret_cast<const uchar4*>(mask)[k];
      char4 o;
      o.x = m.x ? (x.x > 0.5f ? 1 : 0) : -1;
      o.y = m.y ? (x.y > 0.5f ? 1 : 0) : -1;
      o.z = m.z ? (x.z > 0.5f ? 1 : 0) : -1;
      o.w = m.w ? (x.w > 0.5f ? 1 : 0) : -1;
      reinterpret_cast<char4*>(out)[k] = o;
    }
  }
  for (int64_t k = (vec ? n4 * 4 : 0) + t0; k < n; k += stride) out[k] = mask[k] ? (resp[k] > 0.5f ? 1 : 0) : -1;
}

__global__ void __launch_bounds__(256) unpack_kernel(int64_t n, const int8_t* __restrict__ in,
                                                     float* __restrict__ resp, uint8_t* __restrict__ mask) {
  const int64_t n4 = n >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(resp) & 15) | (reinterpret_cast<uintptr_t>(mask) & 3) |
                    (reinterpret_cast<uintptr_t>(in) & 3)) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    for (int64_t k = t0; k < n4; k += stride) {
      const char4 c = reinterpret_cast<const char4*>(in)[k];
      reinterpret_cast<float4*>(resp)[k] = make_float4((float)c.x, (float)c.y, (float)c.z, (float)c.w);
      reinterpret_cast<uchar4*>(mask)[k] = make_uchar4(c.x >= 0, c.y >= 0, c.z >= 0, c.w >= 0);
    }
  }
  for (int64_t k = (vec ? n4 * 4 : 0) + t0; k < n; k += stride) {
    resp[k] = (float)in[k];
    mask[k] = in[k] >= 0;
  }
}

static int pack_grid(int64_t n) {
  int64_t b = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

cudaError_t launch_pack(int64_t n, const float* resp, const uint8_t* mask, int8_t* out, cudaStream_t st) {
  pack_kernel<<<pack_grid(n), 256, 0, st>>>(n, resp, mask, out);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_unpack(int64_t n, const int8_t* in, float* resp, uint8_t* mask, cudaStream_t st) {
  unpack_kernel<<<pack_grid(n), 256, 0, st>>>(n, in, resp, mask);
  note_launch();
  return cudaGetLastError();
}

}  // namespace vibo
