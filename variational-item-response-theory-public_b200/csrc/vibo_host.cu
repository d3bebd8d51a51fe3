// Host-buffer entry point: streams person chunks host->device on a private
// copy stream, overlapped with the kernels of the previous chunk (double
// buffered), and accumulates the per-chunk results on the device.  This is the
// end-to-end path a caller with the reference's host-resident dataset arrays
// (src/datasets.py:928-940) uses.
#include <string>

#include "vibo_hostpack.h"
#include "vibo_kernels.h"

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct HostPipe {
  cudaStream_t copy = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr};
  cudaEvent_t consumed[2] = {nullptr, nullptr};
  cudaEvent_t start = nullptr;
  // host-compressed route: pinned landing buffers of the CPU packers and the events of their H2D copies
  static constexpr int kPackBufs = 3;
  int8_t* pack_host[kPackBufs] = {nullptr, nullptr, nullptr};
  size_t pack_cap = 0;
  cudaEvent_t pack_sent[kPackBufs] = {nullptr, nullptr, nullptr};
  cudaEvent_t raw_sent = nullptr;   // behind the raw part of the current chunk (share controller)
  double share = -1.0;              // current host-packed share of a chunk (< 0: not initialised)
  int at_floor = 0;                 // consecutive feedbacks with the share at its floor
  int rest_calls = 0;               // calls left during which the route stays off (it did not pay)
  int device = -1;
  cudaError_t ensure() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (copy != nullptr && dev == device) return cudaSuccess;
    device = dev;
    if ((e = cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking)) != cudaSuccess) return e;
    for (int i = 0; i < 2; ++i) {
      if ((e = cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    for (int i = 0; i < kPackBufs; ++i)
      if ((e = cudaEventCreateWithFlags(&pack_sent[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&raw_sent, cudaEventDisableTiming)) != cudaSuccess) return e;
    return cudaEventCreateWithFlags(&start, cudaEventDisableTiming);
  }
  cudaError_t ensure_pack(size_t bytes) {
    if (bytes <= pack_cap) return cudaSuccess;
    for (int i = 0; i < kPackBufs; ++i) {
      if (pack_host[i] != nullptr) cudaFreeHost(pack_host[i]);
      pack_host[i] = nullptr;
    }
    pack_cap = 0;
    for (int i = 0; i < kPackBufs; ++i) {
      cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&pack_host[i]), bytes, cudaHostAllocDefault);
      if (e != cudaSuccess) return e;
    }
    pack_cap = bytes;
    return cudaSuccess;
  }
};

thread_local HostPipe g_pipe;

struct StagingLayout {
  size_t resp[2], mask[2], packed[2], scalars, g_table, g_item, total;
};

StagingLayout staging_layout(const vibo_desc& d, int64_t chunk) {
  StagingLayout L;
  const size_t cells = (size_t)chunk * d.num_item;
  const size_t F = (size_t)vibo::item_width_host(d.irt_model, d.ability_dim);
  size_t off = 0;
  for (int b = 0; b < 2; ++b) {
    L.resp[b] = off; off += align_up(cells * 4, 256);
    L.mask[b] = off; off += align_up(cells, 256);
  }
  for (int b = 0; b < 2; ++b) {
    L.packed[b] = off; off += align_up(cells, 256);  // landing buffers of the packed transfer format
  }
  L.scalars = off; off += 256;
  L.g_table = off; off += align_up(4 * 2 * (size_t)(d.conditional ? d.num_item : 1) * 2 * d.ability_dim, 256);
  L.g_item = off; off += align_up(4 * (size_t)d.num_item * F, 256);
  L.total = off;
  return L;
}

// Share of every chunk's rows that crosses PCIe host-compressed (packed to 1 B/cell by the host thread pool
// while the rest of the chunk is in flight in the reference layout).  The two routes use different resources
// (CPU cores + host DRAM vs the PCIe link), so their rates add: with C GB/s of packing and B GB/s of DMA the
// balance is share = 5 C / (5 B + 4 C) (~0.73 for 90 and 51 GB/s).  The share starts at 0.72 and follows the
// machine: after packing a chunk, if the DMA engine has already drained the chunk's raw part it was starved
// (the CPU route is the slower one: less packing), otherwise it is the bottleneck (more packing).
// VIBO_HOST_PACK_FRACTION fixes the share; 0 = DMA only.
// When the host DRAM, not the link, is the shared bottleneck (several GPUs fed from one host) packing cannot win:
// it reads the same 5 B/cell and adds the packed copy.  The route therefore needs >= 6 pool threads per process,
// and if the controller stays pinned at the floor it switches the route off for the next 16 calls.
constexpr double kShareInit = 0.72, kShareStep = 0.02, kShareMin = 0.2, kShareMax = 0.92;
constexpr int kShareMinThreads = 6, kFloorPatience = 24, kRestCalls = 16;
bool host_pack_fixed(double* f) {
  const char* e = getenv("VIBO_HOST_PACK_FRACTION");
  if (e == nullptr) return false;
  const double v = atof(e);
  *f = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
  return true;
}
double host_pack_fraction(size_t chunk_cells) {
  if (chunk_cells < (1u << 20)) return 0.0;   // tiny chunks: not worth the synchronisation
  double f;
  if (host_pack_fixed(&f)) return f;
  if (vibo::host_pool_threads() < kShareMinThreads) return 0.0;   // too few cores per GPU to help
  if (g_pipe.rest_calls > 0) return 0.0;
  if (g_pipe.share < 0.0) g_pipe.share = kShareInit;
  return g_pipe.share;
}
void host_pack_feedback(bool dma_starved) {
  double f;
  if (host_pack_fixed(&f) || g_pipe.share < 0.0) return;
  g_pipe.share += dma_starved ? -kShareStep : kShareStep;
  if (g_pipe.share > kShareMax) g_pipe.share = kShareMax;
  if (g_pipe.share <= kShareMin) {
    g_pipe.share = kShareMin;
    if (++g_pipe.at_floor >= kFloorPatience) {   // the CPU route keeps losing: stop for a while, then probe again
      g_pipe.at_floor = 0;
      g_pipe.rest_calls = kRestCalls;
      g_pipe.share = 0.3;
    }
  } else {
    g_pipe.at_floor = 0;
  }
}

}  // namespace

extern "C" {

size_t vibo_host_staging_bytes(const vibo_desc* desc, int64_t chunk_person) {
  if (desc == nullptr || chunk_person <= 0) return 0;
  return staging_layout(*desc, chunk_person).total;
}

static int fused_elbo_host_impl(const vibo_desc* desc, const float* response_host,
                                const uint8_t* mask_host, const int8_t* packed_host, const float* table,
                                const float* item_feat, const float* eps_ability, uint64_t seed, float beta,
                                double* out_scalars, double* out_scalars_host, float* g_table,
                                float* g_item, int64_t chunk_person, void* staging,
                                size_t staging_bytes, void* workspace, size_t workspace_bytes,
                                void* stream) {
  const bool packed = packed_host != nullptr;
  if (desc == nullptr || (!packed && (response_host == nullptr || mask_host == nullptr)) || staging == nullptr ||
      out_scalars == nullptr || chunk_person <= 0)
    return vibo::set_last_error(VIBO_ERR_BAD_ARGUMENT,
                                "vibo_fused_elbo_host: desc, response_host, mask_host, staging and out_scalars "
                                "are required and chunk_person must be > 0");
  const vibo_desc& d = *desc;
  const StagingLayout L = staging_layout(d, chunk_person);
  if (staging_bytes < L.total)
    return vibo::set_last_error(VIBO_ERR_WORKSPACE, "staging smaller than vibo_host_staging_bytes(desc, chunk_person)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t ce;
#define VIBO_HOST_CUDA(call, where)                                                                  \
  do {                                                                                               \
    if ((ce = (call)) != cudaSuccess)                                                                \
      return vibo::set_last_error(VIBO_ERR_CUDA, (std::string("vibo_fused_elbo_host: ") + where + ": " + \
                                                  cudaGetErrorString(ce)).c_str());                  \
  } while (0)
  VIBO_HOST_CUDA(g_pipe.ensure(), "copy stream / events");
  char* base = static_cast<char*>(staging);
  double* tmp_scalars = reinterpret_cast<double*>(base + L.scalars);
  float* tmp_g_table = reinterpret_cast<float*>(base + L.g_table);
  float* tmp_g_item = reinterpret_cast<float*>(base + L.g_item);
  const bool grad = g_item != nullptr;
  const size_t F = (size_t)vibo::item_width_host(d.irt_model, d.ability_dim);
  const int n_table = 2 * (d.conditional ? d.num_item : 1) * 2 * d.ability_dim;
  const int n_item = (int)(d.num_item * F);

  VIBO_HOST_CUDA(cudaMemsetAsync(out_scalars, 0, 2 * sizeof(double), st), "memset");
  if (grad) {
    VIBO_HOST_CUDA(cudaMemsetAsync(g_table, 0, sizeof(float) * n_table, st), "memset");
    VIBO_HOST_CUDA(cudaMemsetAsync(g_item, 0, sizeof(float) * n_item, st), "memset");
  }
  // staging buffers may still be read by earlier work on `st`
  VIBO_HOST_CUDA(cudaEventRecord(g_pipe.start, st), "event record");
  VIBO_HOST_CUDA(cudaStreamWaitEvent(g_pipe.copy, g_pipe.start, 0), "stream wait");

  const size_t chunk_cells = (size_t)(d.num_person < chunk_person ? d.num_person : chunk_person) * d.num_item;
  if (!packed && g_pipe.rest_calls > 0 && chunk_cells >= (1u << 20)) --g_pipe.rest_calls;
  const bool pack_route = !packed && host_pack_fraction(chunk_cells) > 0.0;
  if (pack_route)   // sized for the largest share the controller can reach
    VIBO_HOST_CUDA(g_pipe.ensure_pack((size_t)chunk_person * d.num_item), "pinned pack buffers");
  int64_t c = 0;
  for (int64_t r0 = 0; r0 < d.num_person; r0 += chunk_person, ++c) {
    const int b = (int)(c & 1);
    const int64_t n = (d.num_person - r0 < chunk_person) ? d.num_person - r0 : chunk_person;
    if (c >= 2) VIBO_HOST_CUDA(cudaStreamWaitEvent(g_pipe.copy, g_pipe.consumed[b], 0), "stream wait");
    int64_t unpack_cells = 0;
    size_t unpack_first = 0;
    if (packed) {
      VIBO_HOST_CUDA(cudaMemcpyAsync(base + L.packed[b], packed_host + (size_t)r0 * d.num_item,
                                     (size_t)n * d.num_item, cudaMemcpyHostToDevice, g_pipe.copy),
                     "H2D copy of packed rows");
    } else {
      // rows [0, n_raw) of the chunk in the reference layout by DMA; rows [n_raw, n) packed on the host
      // (while that DMA runs) and sent as 1 B/cell
      const double pack_f = pack_route ? host_pack_fraction(chunk_cells) : 0.0;   // follows the controller
      int64_t n_raw = n;
      if (pack_f > 0.0) {   // the raw part is a multiple of 16 rows: the unpacked rows behind it stay 16-byte aligned
        n_raw = (n - (int64_t)((double)n * pack_f) + 15) / 16 * 16;
        if (pack_f >= 1.0) n_raw = 0;
        if (n_raw > n) n_raw = n;
      }
      const int64_t n_pack = n - n_raw;
      if (n_raw > 0) {
        VIBO_HOST_CUDA(cudaMemcpyAsync(base + L.resp[b], response_host + (size_t)r0 * d.num_item,
                                       (size_t)n_raw * d.num_item * sizeof(float), cudaMemcpyHostToDevice,
                                       g_pipe.copy),
                       "H2D copy of response rows");
        VIBO_HOST_CUDA(cudaMemcpyAsync(base + L.mask[b], mask_host + (size_t)r0 * d.num_item,
                                       (size_t)n_raw * d.num_item, cudaMemcpyHostToDevice, g_pipe.copy),
                       "H2D copy of mask rows");
        if (n_pack > 0) VIBO_HOST_CUDA(cudaEventRecord(g_pipe.raw_sent, g_pipe.copy), "event record");
      }
      if (n_pack > 0) {
        const int hb = (int)(c % HostPipe::kPackBufs);
        const size_t cells = (size_t)n_pack * d.num_item, first = (size_t)(r0 + n_raw) * d.num_item;
        if (c >= HostPipe::kPackBufs)   // the copy that last read this pinned buffer
          VIBO_HOST_CUDA(cudaEventSynchronize(g_pipe.pack_sent[hb]), "event synchronize");
        vibo::host_pack_parallel(response_host + first, mask_host + first, g_pipe.pack_host[hb], cells);
        if (n_raw > 0 && c >= 1) host_pack_feedback(cudaEventQuery(g_pipe.raw_sent) == cudaSuccess);
        VIBO_HOST_CUDA(cudaMemcpyAsync(base + L.packed[b], g_pipe.pack_host[hb], cells, cudaMemcpyHostToDevice,
                                       g_pipe.copy),
                       "H2D copy of host-packed rows");
        VIBO_HOST_CUDA(cudaEventRecord(g_pipe.pack_sent[hb], g_pipe.copy), "event record");
        unpack_cells = (int64_t)cells;
        unpack_first = (size_t)n_raw * d.num_item;
      }
    }
    VIBO_HOST_CUDA(cudaEventRecord(g_pipe.copied[b], g_pipe.copy), "event record");
    VIBO_HOST_CUDA(cudaStreamWaitEvent(st, g_pipe.copied[b], 0), "stream wait");
    if (packed)  // 1 B/cell crossed PCIe; expand to the (response, mask) pair the row kernels read
      VIBO_HOST_CUDA(vibo::launch_unpack(n * d.num_item, reinterpret_cast<const int8_t*>(base + L.packed[b]),
                                         reinterpret_cast<float*>(base + L.resp[b]),
                                         reinterpret_cast<uint8_t*>(base + L.mask[b]), st),
                     "unpack");
    else if (unpack_cells > 0)
      VIBO_HOST_CUDA(vibo::launch_unpack(unpack_cells, reinterpret_cast<const int8_t*>(base + L.packed[b]),
                                         reinterpret_cast<float*>(base + L.resp[b]) + unpack_first,
                                         reinterpret_cast<uint8_t*>(base + L.mask[b]) + unpack_first, st),
                     "unpack");
    vibo_desc dc = d;
    dc.num_person = n;
    dc.person_offset = d.person_offset + r0;
    const int rc = vibo_fused_elbo(&dc, reinterpret_cast<const float*>(base + L.resp[b]),
                                   reinterpret_cast<const uint8_t*>(base + L.mask[b]), table, item_feat,
                                   eps_ability ? eps_ability + (size_t)r0 * d.ability_dim : nullptr, seed,
                                   beta, tmp_scalars, nullptr, nullptr, nullptr,
                                   grad ? tmp_g_table : nullptr, grad ? tmp_g_item : nullptr, workspace,
                                   workspace_bytes, st);
    if (rc != VIBO_OK) return rc;  // vibo_fused_elbo recorded the message
    if (grad) {
      VIBO_HOST_CUDA(vibo::launch_accumulate(g_table, tmp_g_table, n_table, out_scalars, tmp_scalars, 2, st),
                     "accumulate");
      VIBO_HOST_CUDA(vibo::launch_accumulate(g_item, tmp_g_item, n_item, nullptr, nullptr, 0, st), "accumulate");
    } else {
      VIBO_HOST_CUDA(vibo::launch_accumulate(nullptr, nullptr, 0, out_scalars, tmp_scalars, 2, st), "accumulate");
    }
    VIBO_HOST_CUDA(cudaEventRecord(g_pipe.consumed[b], st), "event record");
  }
  if (out_scalars_host != nullptr)
    VIBO_HOST_CUDA(cudaMemcpyAsync(out_scalars_host, out_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, st),
                   "D2H copy of the scalars");
  VIBO_HOST_CUDA(cudaStreamSynchronize(st), "stream synchronize");
#undef VIBO_HOST_CUDA
  return VIBO_OK;
}

int vibo_fused_elbo_host(const vibo_desc* desc, const float* response_host,
                         const uint8_t* mask_host, const float* table, const float* item_feat,
                         const float* eps_ability, uint64_t seed, float beta,
                         double* out_scalars, double* out_scalars_host, float* g_table,
                         float* g_item, int64_t chunk_person, void* staging,
                         size_t staging_bytes, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return fused_elbo_host_impl(desc, response_host, mask_host, nullptr, table, item_feat, eps_ability, seed, beta,
                              out_scalars, out_scalars_host, g_table, g_item, chunk_person, staging,
                              staging_bytes, workspace, workspace_bytes, stream);
}

int vibo_fused_elbo_host_packed(const vibo_desc* desc, const int8_t* packed_host, const float* table,
                                const float* item_feat, const float* eps_ability, uint64_t seed,
                                float beta, double* out_scalars, double* out_scalars_host,
                                float* g_table, float* g_item, int64_t chunk_person, void* staging,
                                size_t staging_bytes, void* workspace, size_t workspace_bytes,
                                void* stream) {
  if (packed_host == nullptr)
    return vibo::set_last_error(VIBO_ERR_BAD_ARGUMENT, "vibo_fused_elbo_host_packed: packed_host is NULL");
  return fused_elbo_host_impl(desc, nullptr, nullptr, packed_host, table, item_feat, eps_ability, seed, beta,
                              out_scalars, out_scalars_host, g_table, g_item, chunk_person, staging,
                              staging_bytes, workspace, workspace_bytes, stream);
}

int vibo_pack(const vibo_desc* desc, const float* response, const uint8_t* mask, int8_t* packed, void* stream) {
  if (desc == nullptr || response == nullptr || mask == nullptr || packed == nullptr || desc->num_person < 0 ||
      desc->num_item <= 0)
    return vibo::set_last_error(VIBO_ERR_BAD_ARGUMENT, "vibo_pack: bad argument");
  if (desc->num_person == 0) return VIBO_OK;
  if (vibo::launch_pack(desc->num_person * (int64_t)desc->num_item, response, mask, packed,
                        static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return vibo::set_last_error(VIBO_ERR_CUDA, "vibo_pack: launch failed");
  return VIBO_OK;
}

int vibo_pack_host(const vibo_desc* desc, const float* response_host, const uint8_t* mask_host,
                   int8_t* packed_host) {
  if (desc == nullptr || response_host == nullptr || mask_host == nullptr || packed_host == nullptr ||
      desc->num_person < 0 || desc->num_item <= 0)
    return vibo::set_last_error(VIBO_ERR_BAD_ARGUMENT, "vibo_pack_host: bad argument");
  vibo::host_pack_parallel(response_host, mask_host, packed_host, (size_t)desc->num_person * desc->num_item);
  return VIBO_OK;
}

int vibo_host_threads(void) { return vibo::host_pool_threads(); }

double vibo_host_pack_share(const vibo_desc* desc, int64_t chunk_person) {
  if (desc == nullptr || chunk_person <= 0) return 0.0;
  const int64_t rows = desc->num_person < chunk_person ? desc->num_person : chunk_person;
  return host_pack_fraction((size_t)rows * desc->num_item);
}

int vibo_unpack(const vibo_desc* desc, const int8_t* packed, float* response, uint8_t* mask, void* stream) {
  if (desc == nullptr || response == nullptr || mask == nullptr || packed == nullptr || desc->num_person < 0 ||
      desc->num_item <= 0)
    return vibo::set_last_error(VIBO_ERR_BAD_ARGUMENT, "vibo_unpack: bad argument");
  if (desc->num_person == 0) return VIBO_OK;
  if (vibo::launch_unpack(desc->num_person * (int64_t)desc->num_item, packed, response, mask,
                          static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return vibo::set_last_error(VIBO_ERR_CUDA, "vibo_unpack: launch failed");
  return VIBO_OK;
}

}  // extern "C"
