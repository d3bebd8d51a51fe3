// One-shot all-reduce over NVLink peer memory for the per-step exchange of the
// person-sharded ELBO step (ONE sum of [loss | parameter
// gradients], 26-110 KB, per step).
//
// Why not NCCL: the buffer is tiny, so the cost is pure launch + protocol
// latency, and the call has to live INSIDE the CUDA graph of the step.  Here
// every rank owns one cudaMalloc'ed region (exported to its peers with CUDA
// IPC) holding two send buffers and a pad of arrival flags; the kernel
//   1. copies the rank's slice into its send buffer (parity = epoch & 1),
//   2. publishes the epoch into every peer's flag pad (fence.sys + st.release.sys)
//      and spins until every peer's epoch has arrived in its own pad,
//   3. sums the W peer buffers with direct NVLink loads in RANK ORDER, so every
//      rank obtains bit-identical results (replicated Adam stays in lock-step).
// Two send buffers are enough: a rank can only overwrite parity p again after
// passing the barrier of the step in between, which every peer enters only
// after it has finished reading the earlier step.
// The kernel is a plain launch with fixed pointers: capturable in a CUDA graph;
// the epoch lives in device memory and advances on every execution.
#include <cstdio>
#include <cstring>
#include <new>

#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

constexpr int kCommMaxWorld = 16;
constexpr int kCommMaxBlocks = 32;
constexpr int kCommThreads = 256;
constexpr size_t kCommFlagBytes = (size_t)kCommMaxBlocks * kCommMaxWorld * sizeof(uint32_t);  // arrival pad
constexpr size_t kCommEpochOff = kCommFlagBytes;                                              // local epochs
constexpr size_t kCommErrorOff = kCommEpochOff + 256;                                         // timeout flag
constexpr size_t kCommSendOff = 4096;

// Optional Adam step on the reduced gradients (vibo_comm_allreduce_adam): element i >= skip of the
// vector is the gradient of parameter i - skip.
struct CommAdam {
  float *param, *m, *v;
  const int64_t* step;
  float lr, b1, b2, eps;
  size_t skip;
};

struct CommDev {
  char* base[kCommMaxWorld];  // every rank's region (base[rank] is local)
  size_t cap_floats;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float4 ld_relaxed_sys4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// Every block owns one slice of the vector and runs the whole protocol for it (its own row of
// the flag pad), so the latency is one flag round trip plus ONE batch of W independent peer loads
// per thread: the grid is sized so that a thread reduces at most one float4.  All blocks of a call
// share one epoch (the call count): it selects the send buffer, and every slot of the epoch table
// is advanced on every call, so calls of different lengths can be mixed freely.
template <bool ADAM>
__global__ void __launch_bounds__(kCommThreads) comm_allreduce_kernel(const __grid_constant__ CommDev c,
                                                                      float* __restrict__ data, size_t n,
                                                                      size_t per, const CommAdam a) {
  __shared__ uint32_t s_epoch;
  __shared__ float s_bc[2];
  if (ADAM && threadIdx.x == 32) adam_bias_corrections(a.step, a.b1, a.b2, &s_bc[0], &s_bc[1]);
  const int b = blockIdx.x, t = threadIdx.x;
  char* mine = c.base[c.rank];
  uint32_t* epochs = reinterpret_cast<uint32_t*>(mine + kCommEpochOff);
  if (t == 0) s_epoch = epochs[b] + 1u;
  __syncthreads();
  const uint32_t e = s_epoch;
  const size_t par_off = kCommSendOff + (size_t)(e & 1u) * c.cap_floats * sizeof(float);
  const size_t lo = (size_t)b * per, hi = lo + per < n ? lo + per : n;   // per is a multiple of 4
  const bool vec = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
  const size_t hi4 = vec ? lo + ((hi - lo) & ~(size_t)3) : lo;
  // 1. my slice -> my send buffer
  float* send = reinterpret_cast<float*>(mine + par_off);
  for (size_t i = lo + 4 * (size_t)t; i < hi4; i += 4 * kCommThreads)
    *reinterpret_cast<float4*>(send + i) = *reinterpret_cast<const float4*>(data + i);
  for (size_t i = hi4 + t; i < hi; i += kCommThreads) send[i] = data[i];
  __syncthreads();
  // 2. arrive at every peer, wait for every peer
  if (t < c.world) {
    __threadfence_system();
    uint32_t* peer_pad = reinterpret_cast<uint32_t*>(c.base[t]) + b * kCommMaxWorld + c.rank;
    st_release_sys(peer_pad, e);
    const uint32_t* my_pad = reinterpret_cast<const uint32_t*>(mine) + b * kCommMaxWorld + t;
    // bounded spin (20 s): a peer that died must not leave this GPU hung
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(ld_acquire_sys(my_pad) - e) < 0) {
      __nanosleep(20);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) {
        *reinterpret_cast<uint32_t*>(mine + kCommErrorOff) = 1u;
        break;
      }
    }
  }
  __syncthreads();
  // 3. sum the peers' slices in rank order (identical result on every rank); the W loads of a
  //    thread are independent, so they are all in flight together
  for (size_t i = lo + 4 * (size_t)t; i < hi4; i += 4 * kCommThreads) {
    float4 v[kCommMaxWorld];
#pragma unroll
    for (int r = 0; r < kCommMaxWorld; ++r)
      if (r < c.world) v[r] = ld_relaxed_sys4(reinterpret_cast<const float*>(c.base[r] + par_off) + i);
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int r = 0; r < kCommMaxWorld; ++r)
      if (r < c.world) {
        acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w;
      }
    *reinterpret_cast<float4*>(data + i) = acc;
    if (ADAM) {
      const float g[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (i + q >= a.skip) {
          const size_t k = i + q - a.skip;
          adam_update(g[q], a.param + k, a.m + k, a.v + k, a.lr / s_bc[0], s_bc[1], a.b1, a.b2, a.eps);
        }
    }
  }
  for (size_t i = hi4 + t; i < hi; i += kCommThreads) {
    float acc = 0.0f;
    for (int r = 0; r < c.world; ++r)
      acc += ld_relaxed_sys(reinterpret_cast<const float*>(c.base[r] + par_off) + i);
    data[i] = acc;
    if (ADAM && i >= a.skip) {
      const size_t k = i - a.skip;
      adam_update(acc, a.param + k, a.m + k, a.v + k, a.lr / s_bc[0], s_bc[1], a.b1, a.b2, a.eps);
    }
  }
  // advance the epoch: this block's slot, and (block 0) the slots of the blocks this call did not use
  if (t == 0) epochs[b] = e;
  if (b == 0 && t >= (int)gridDim.x && t < kCommMaxBlocks) epochs[t] = e;
}

}  // namespace vibo

struct vibo_comm {
  vibo::CommDev dev;
  void* local = nullptr;
  size_t bytes = 0;
  bool connected = false;
};

namespace {
thread_local char g_comm_error[256];
int comm_fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  if (e != cudaSuccess) snprintf(g_comm_error, sizeof g_comm_error, "%s: %s", what, cudaGetErrorString(e));
  else snprintf(g_comm_error, sizeof g_comm_error, "%s", what);
  return code;
}
}  // namespace

extern "C" {

const char* vibo_comm_last_error(void) { return g_comm_error; }

int vibo_comm_create(int rank, int world_size, size_t max_floats, vibo_comm** out, void* handle_out) {
  if (out == nullptr || handle_out == nullptr || rank < 0 || rank >= world_size ||
      world_size > vibo::kCommMaxWorld || max_floats == 0)
    return comm_fail(VIBO_ERR_BAD_ARGUMENT, "vibo_comm_create: bad argument (world_size <= 16)");
  static_assert(sizeof(cudaIpcMemHandle_t) == VIBO_COMM_HANDLE_BYTES, "handle size");
  vibo_comm* c = new (std::nothrow) vibo_comm();
  if (c == nullptr) return comm_fail(VIBO_ERR_CUDA, "out of host memory");
  const size_t cap = (max_floats + 63) / 64 * 64;
  c->bytes = vibo::kCommSendOff + 2 * cap * sizeof(float);
  cudaError_t e = cudaMalloc(&c->local, c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
  if (e != cudaSuccess) {
    if (c->local) cudaFree(c->local);
    delete c;
    return comm_fail(VIBO_ERR_CUDA, "vibo_comm_create", e);
  }
  memcpy(handle_out, &h, sizeof h);
  memset(&c->dev, 0, sizeof c->dev);
  c->dev.cap_floats = cap;
  c->dev.rank = rank;
  c->dev.world = world_size;
  c->dev.base[rank] = static_cast<char*>(c->local);
  *out = c;
  return VIBO_OK;
}

int vibo_comm_connect(vibo_comm* c, const void* all_handles) {
  if (c == nullptr || all_handles == nullptr) return comm_fail(VIBO_ERR_BAD_ARGUMENT, "vibo_comm_connect: NULL");
  const char* hs = static_cast<const char*>(all_handles);
  for (int r = 0; r < c->dev.world; ++r) {
    if (r == c->dev.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + (size_t)r * sizeof h, sizeof h);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return comm_fail(VIBO_ERR_CUDA, "cudaIpcOpenMemHandle (peer access over NVLink)", e);
    c->dev.base[r] = static_cast<char*>(p);
  }
  c->connected = true;
  return VIBO_OK;
}

namespace {
int comm_launch(vibo_comm* c, float* data, size_t n, const vibo::CommAdam* adam, void* stream, const char* who) {
  if (c == nullptr || data == nullptr) return comm_fail(VIBO_ERR_BAD_ARGUMENT, "vibo_comm_allreduce: NULL");
  if (!c->connected && c->dev.world > 1) return comm_fail(VIBO_ERR_BAD_ARGUMENT, "vibo_comm_allreduce: not connected");
  if (n > c->dev.cap_floats) return comm_fail(VIBO_ERR_WORKSPACE, "vibo_comm_allreduce: n exceeds the capacity given at create");
  if (n == 0) return VIBO_OK;
  // one float4 per thread while the blocks last, equal 4-aligned slices beyond that
  int blocks = (int)((n + 4 * vibo::kCommThreads - 1) / (4 * vibo::kCommThreads));
  if (blocks > vibo::kCommMaxBlocks) blocks = vibo::kCommMaxBlocks;
  const size_t per = ((n + blocks - 1) / blocks + 3) / 4 * 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (adam != nullptr)
    vibo::comm_allreduce_kernel<true><<<blocks, vibo::kCommThreads, 0, st>>>(c->dev, data, n, per, *adam);
  else
    vibo::comm_allreduce_kernel<false><<<blocks, vibo::kCommThreads, 0, st>>>(c->dev, data, n, per, vibo::CommAdam{});
  vibo::note_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return comm_fail(VIBO_ERR_CUDA, who, e);
  return VIBO_OK;
}
}  // namespace

int vibo_comm_allreduce(vibo_comm* c, float* data, size_t n, void* stream) {
  return comm_launch(c, data, n, nullptr, stream, "comm_allreduce_kernel");
}

int vibo_comm_allreduce_adam(vibo_comm* c, float* data, size_t n, size_t skip, float* param, float* exp_avg,
                             float* exp_avg_sq, const int64_t* step, float lr, float beta1, float beta2,
                             float eps, void* stream) {
  if (param == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr || step == nullptr || skip > n)
    return comm_fail(VIBO_ERR_BAD_ARGUMENT, "vibo_comm_allreduce_adam: bad argument");
  const vibo::CommAdam a{param, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, skip};
  return comm_launch(c, data, n, &a, stream, "comm_allreduce_kernel<adam>");
}

int vibo_comm_status(vibo_comm* c) {
  if (c == nullptr) return VIBO_ERR_BAD_ARGUMENT;
  uint32_t flag = 0;
  cudaError_t e = cudaMemcpy(&flag, static_cast<char*>(c->local) + vibo::kCommErrorOff, sizeof flag,
                             cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return comm_fail(VIBO_ERR_CUDA, "vibo_comm_status", e);
  if (flag != 0) return comm_fail(VIBO_ERR_CUDA, "a peer did not arrive within 20 s (all-reduce timed out)");
  return VIBO_OK;
}

int vibo_comm_destroy(vibo_comm* c) {
  if (c == nullptr) return VIBO_OK;
  for (int r = 0; r < c->dev.world; ++r)
    if (r != c->dev.rank && c->dev.base[r] != nullptr) cudaIpcCloseMemHandle(c->dev.base[r]);
  if (c->local) cudaFree(c->local);
  delete c;
  return VIBO_OK;
}

}  // extern "C"
