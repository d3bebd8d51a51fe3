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

#include "vibo_kernels.h"

namespace vibo {

constexpr int kCommMaxWorld = 16;
constexpr int kCommMaxBlocks = 16;
constexpr int kCommThreads = 512;
constexpr size_t kCommFlagBytes = (size_t)kCommMaxBlocks * kCommMaxWorld * sizeof(uint32_t);  // arrival pad
constexpr size_t kCommEpochOff = kCommFlagBytes;                                              // local epochs
constexpr size_t kCommErrorOff = kCommEpochOff + 256;                                         // timeout flag
constexpr size_t kCommSendOff = 4096;

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

__global__ void __launch_bounds__(kCommThreads) comm_allreduce_kernel(const __grid_constant__ CommDev c,
                                                                      float* __restrict__ data, size_t n) {
  __shared__ uint32_t s_epoch;
  const int b = blockIdx.x, t = threadIdx.x;
  char* mine = c.base[c.rank];
  uint32_t* my_epoch = reinterpret_cast<uint32_t*>(mine + kCommEpochOff) + b;
  if (t == 0) s_epoch = *my_epoch + 1u;
  __syncthreads();
  const uint32_t e = s_epoch;
  const size_t par_off = kCommSendOff + (size_t)(e & 1u) * c.cap_floats * sizeof(float);
  const size_t per = (n + gridDim.x - 1) / gridDim.x;
  const size_t lo = (size_t)b * per, hi = lo + per < n ? lo + per : n;
  // 1. my slice -> my send buffer
  float* send = reinterpret_cast<float*>(mine + par_off);
  for (size_t i = lo + t; i < hi; i += kCommThreads) send[i] = data[i];
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
      __nanosleep(32);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) {
        *reinterpret_cast<uint32_t*>(mine + kCommErrorOff) = 1u;
        break;
      }
    }
  }
  __syncthreads();
  // 3. sum the peers' slices in rank order (identical result on every rank)
  for (size_t i = lo + t; i < hi; i += kCommThreads) {
    float acc = 0.0f;
    for (int r = 0; r < c.world; ++r)
      acc += ld_relaxed_sys(reinterpret_cast<const float*>(c.base[r] + par_off) + i);
    data[i] = acc;
  }
  if (t == 0) *my_epoch = e;
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

int vibo_comm_allreduce(vibo_comm* c, float* data, size_t n, void* stream) {
  if (c == nullptr || data == nullptr) return comm_fail(VIBO_ERR_BAD_ARGUMENT, "vibo_comm_allreduce: NULL");
  if (!c->connected && c->dev.world > 1) return comm_fail(VIBO_ERR_BAD_ARGUMENT, "vibo_comm_allreduce: not connected");
  if (n > c->dev.cap_floats) return comm_fail(VIBO_ERR_WORKSPACE, "vibo_comm_allreduce: n exceeds the capacity given at create");
  if (n == 0) return VIBO_OK;
  int blocks = (int)((n + 4095) / 4096);
  if (blocks > vibo::kCommMaxBlocks) blocks = vibo::kCommMaxBlocks;
  vibo::comm_allreduce_kernel<<<blocks, vibo::kCommThreads, 0, static_cast<cudaStream_t>(stream)>>>(c->dev, data, n);
  vibo::note_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return comm_fail(VIBO_ERR_CUDA, "comm_allreduce_kernel", e);
  return VIBO_OK;
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
