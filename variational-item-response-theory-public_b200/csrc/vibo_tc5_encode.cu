// Conditional-posterior encode on the 5th-generation tensor cores (tcgen05.mma, TMEM, 2-D TMA).
//
// With --conditional-posterior the expert of a cell depends on the item, and the product-of-experts
// sums of a fully observed row are a matrix product with the 0/1 response matrix
// (models.py:664-710 + :596-629, utils.py:105-113):
//     S_i[d] = sum_j tau^0_jd  +  sum_j x_ij (tau^1_jd - tau^0_jd),      N_i[d] likewise with mu tau.
// This kernel streams the response matrix ONCE through the tensor cores:
//   * A = a 128-person x 32-item tile of the float32 response matrix, brought by TMA
//     (cp.async.bulk.tensor.2d, 128-byte swizzle) straight into the UMMA canonical K-major layout:
//     0 / 1 (and the -1 of missing cells) are exact in TF32, so there is NO conversion pass;
//   * B = the table differences (tau^1 - tau^0 | mu^1 tau^1 - mu^0 tau^0), split into THREE TF32 terms
//     (hi + mid + lo = the fp32 value to ~2^-31) -> 6D <= 30 columns, resident in shared memory for
//     the whole kernel (K-major, swizzled, prepared once per CTA);
//   * D = 128 x 32 fp32 accumulators in TMEM, double buffered: tcgen05.mma.kind::tf32 (M 128, N 32,
//     K 8) issued by ONE thread, 4 per 32-item block; tcgen05.commit releases the shared-memory
//     stage to the TMA producer and, after the last block, hands the accumulator to the epilogue;
//   * epilogue warps (one thread per person): tcgen05.ld of the row, + the baseline sums, posterior
//     mean / log-variance / precision out.  While the tensor cores work they scan the tile's mask
//     bytes; a row with a missing cell (rare in this regime) is recomputed exactly on the CUDA cores
//     (prior experts or --drop-missing), so the kernel is correct for any mask.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = mask scan + epilogue.
// Bound: HBM (4 B/cell response + 1 B/cell mask scan); replaces encode_mma_kernel (mma.sync).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>

#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

namespace {

constexpr int kT5Rows = 128, kT5KB = 32, kT5N = 32, kT5Stages = 5, kT5Threads = 192;
constexpr uint32_t kT5StageBytes = kT5Rows * kT5KB * 4;   // 16 KB
constexpr uint32_t kT5BBlock = kT5N * kT5KB * 4;          // 4 KB of B per 32-item block
constexpr uint32_t kT5TmemCols = 64;                      // two 32-column accumulators

struct T5Params {
  int64_t P;
  int I, D, missing_policy, n_kb;
  const float* resp;
  const uint8_t* mask;
  const float* table;   // (2, I, 2D)
  float* mu;
  float* lv;
  float* S;
};

struct T5Smem {
  uint32_t b_off, stage_off, base_off, flag_off, bar_off, total;
};
__host__ __device__ inline T5Smem t5_layout(int n_kb) {
  T5Smem L;
  L.b_off = 0;
  L.stage_off = ((uint32_t)n_kb * kT5BBlock + 1023u) / 1024u * 1024u;
  L.base_off = L.stage_off + kT5Stages * kT5StageBytes;
  L.flag_off = L.base_off + 2 * VIBO_MAX_ABILITY_DIM * 4;
  L.bar_off = L.flag_off + 2 * kT5Rows;   // row flags, double buffered by tile parity
  L.total = L.bar_off + 8 * (2 * kT5Stages + 4) + 16;
  return L;
}

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "T5_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra T5_DONE;\n"
      "bra T5_WAIT;\n"
      "T5_DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, 128-byte swizzle, 128-byte rows packed densely (see vibo_percell.cu)
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32: D f32, A / B TF32 (format 2), K-major, N = 32, M = 128
constexpr uint32_t kT5Idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kT5Idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(kT5Threads, 1)
tc5_encode_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ T5Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const T5Smem L = t5_layout(p.n_kb);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int I = p.I, D = p.D, D2 = 2 * D;
  float* s_base = reinterpret_cast<float*>(smem + L.base_off);   // base_S[D] | base_N[D]
  uint8_t* s_flags = smem + L.flag_off;
  const uint32_t bars = saddr(smem + L.bar_off);
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kT5Stages + s); };
  auto acc_full = [&](int b) { return bars + 8u * (2 * kT5Stages + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (2 * kT5Stages + 2 + b); };
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * (2 * kT5Stages + 4));
  const uint32_t b_base = saddr(smem + L.b_off), st_base = saddr(smem + L.stage_off);

  // ---- setup ------------------------------------------------------------------------------------
  if (t == 0) {
    for (int s = 0; s < kT5Stages; ++s) {
      bar_init(full(s), 1);
      bar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      bar_init(acc_full(b), 1);
      bar_init(acc_empty(b), 4);   // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(s_tmem)),
                 "n"(kT5TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B: zero, then the three TF32 terms of every table difference, swizzled K-major
  for (uint32_t k = t; k < (uint32_t)p.n_kb * kT5BBlock / 16; k += kT5Threads)
    reinterpret_cast<uint4*>(smem + L.b_off)[k] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int idx = t; idx < I * D2; idx += kT5Threads) {
    const int j = idx / D2, c = idx % D2, d = c < D ? c : c - D;
    const float m0 = p.table[(size_t)j * D2 + d], l0 = p.table[(size_t)j * D2 + D + d];
    const float m1 = p.table[(size_t)(I + j) * D2 + d], l1 = p.table[(size_t)(I + j) * D2 + D + d];
    const float t0 = 1.0f / (expf(l0) + kPoeEps), t1 = 1.0f / (expf(l1) + kPoeEps);
    const float diff = c < D ? t1 - t0 : m1 * t1 - m0 * t0;
    const float hi = tf32_trunc(diff), mid = tf32_trunc(diff - hi), lo = tf32_trunc(diff - hi - mid);
    const int kb = j / kT5KB, kk = j % kT5KB;
    const float parts[3] = {hi, mid, lo};
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int n = s * D2 + c;
      const uint32_t off = (uint32_t)kb * kT5BBlock + (uint32_t)n * 128u +
                           ((((uint32_t)kk >> 2) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)kk & 3u) * 4u;
      *reinterpret_cast<float*>(smem + L.b_off + off) = parts[s];
    }
  }
  // baseline sums over items of the r = 0 experts (fixed order, double)
  for (int c = warp; c < D2; c += kT5Threads / 32) {
    const int d = c < D ? c : c - D;
    double a = 0.0;
    for (int j = lane; j < I; j += 32) {
      const float m0 = p.table[(size_t)j * D2 + d], l0 = p.table[(size_t)j * D2 + D + d];
      const float t0 = 1.0f / (expf(l0) + kPoeEps);
      a += (double)(c < D ? t0 : m0 * t0);
    }
    a = warp_sum(a);
    if (lane == 0) s_base[c] = (float)a;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;
  const int64_t n_tiles = (p.P + kT5Rows - 1) / kT5Rows;

  if (warp == 0) {
    // ===================== TMA producer ===================================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.n_kb; ++kb) {
          bar_wait(empty(s), ph ^ 1u);
          bar_expect_tx(full(s), kT5StageBytes);
          tma_load_2d(st_base + (uint32_t)s * kT5StageBytes, &tmap, full(s), kb * kT5KB, (int)(tile * kT5Rows));
          if (++s == kT5Stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================================================
    int s = 0, b = 0;
    uint32_t ph = 0, aph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      bar_wait(acc_empty(b), aph ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kb = 0; kb < p.n_kb; ++kb) {
        bar_wait(full(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a = st_base + (uint32_t)s * kT5StageBytes, bb = b_base + (uint32_t)kb * kT5BBlock;
#pragma unroll
          for (int ks = 0; ks < kT5KB / 8; ++ks)   // 8 TF32 = 32 bytes of K per instruction
            umma_tf32(tmem + (uint32_t)b * kT5N, desc_k_sw128(a + ks * 32), desc_k_sw128(bb + ks * 32),
                      (kb | ks) != 0 ? 1u : 0u);
          umma_commit(empty(s));                          // stage free once these MMAs have read it
          if (kb == p.n_kb - 1) umma_commit(acc_full(b));  // accumulator complete
        }
        __syncwarp();
        if (++s == kT5Stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      b ^= 1;
      if (b == 0) aph ^= 1u;
    }
  } else {
    // ===================== mask scan + epilogue (one thread per person of the tile) ==========
    const int e = t - 64;                               // 0..127
    const int m = 32 * (warp & 3) + lane;               // this thread's TMEM lane == row of the tile
    const float prior_tau = p.missing_policy == VIBO_MISSING_PRIOR ? 1.0f / (1.0f + kPoeEps) : 0.0f;
    int b = 0;
    uint32_t aph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row0 = tile * kT5Rows;
      const int rows = (int)((p.P - row0 < kT5Rows) ? p.P - row0 : kT5Rows);
      // ---- scan the tile's mask bytes (contiguous block) for missing cells; the flags of tile
      // k + 2 reuse this buffer only after every thread has passed the barriers of tile k + 1
      uint8_t* s_flag = s_flags + b * kT5Rows;
      s_flag[e] = 0;
      asm volatile("bar.sync 2, 128;" ::: "memory");
      {
        const uint8_t* mb = p.mask + row0 * I;
        const int64_t len = (int64_t)rows * I, n16 = len >> 4;
        const uint4* m16 = reinterpret_cast<const uint4*>(mb);
        // 16 independent 16-byte loads in flight per thread (a load-test-branch loop would expose one
        // memory latency per 16 bytes and make the scan, not HBM, the bound)
        constexpr int B = 16;
        for (int64_t k0 = e; k0 < n16; k0 += 128 * B) {
          uint4 w[B];
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const int64_t k = k0 + (int64_t)u * 128;
            w[u] = k < n16 ? __ldg(m16 + k) : make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
          }
          uint32_t z = 0;
#pragma unroll
          for (int u = 0; u < B; ++u)
            z |= ((w[u].x - 0x01010101u) & ~w[u].x) | ((w[u].y - 0x01010101u) & ~w[u].y) |
                 ((w[u].z - 0x01010101u) & ~w[u].z) | ((w[u].w - 0x01010101u) & ~w[u].w);
          if (z & 0x80808080u) {
            for (int u = 0; u < B; ++u) {
              const int64_t k = k0 + (int64_t)u * 128;
              if (k >= n16) break;
              for (int q = 0; q < 16; ++q)
                if (mb[k * 16 + q] == 0) s_flag[(k * 16 + q) / I] = 1;
            }
          }
        }
        for (int64_t k = (n16 << 4) + e; k < len; k += 128)
          if (mb[k] == 0) s_flag[k / I] = 1;
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const bool exact = s_flag[m] != 0;
      // ---- accumulator row
      bar_wait(acc_full(b), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32];
      const uint32_t taddr = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)b * kT5N;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32"
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
          "[%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(acc_empty(b));   // the MMA warp may overwrite this accumulator
      if (m < rows) {
        const int64_t row = row0 + m;
        for (int d = 0; d < D; ++d) {
          float Ssum, Nsum;
          if (!exact) {
            // hi + mid + lo columns of the S and N differences, plus the r = 0 baseline
            Ssum = s_base[d] + (__uint_as_float(r[d]) + __uint_as_float(r[D2 + d]) + __uint_as_float(r[2 * D2 + d]));
            Nsum = s_base[D + d] +
                   (__uint_as_float(r[D + d]) + __uint_as_float(r[D2 + D + d]) + __uint_as_float(r[2 * D2 + D + d]));
          } else {
            // a row with missing cells: exact recomputation (prior experts or --drop-missing,
            // models.py:606-627)
            Ssum = 0.0f;
            Nsum = 0.0f;
            for (int j = 0; j < I; ++j) {
              if (p.mask[row * I + j]) {
                const int x = p.resp[row * I + j] > 0.5f ? 1 : 0;
                const float mu = p.table[((size_t)x * I + j) * D2 + d], lam = p.table[((size_t)x * I + j) * D2 + D + d];
                const float tau = 1.0f / (expf(lam) + kPoeEps);
                Ssum += tau;
                Nsum = fmaf(mu, tau, Nsum);
              } else {
                Ssum += prior_tau;
              }
            }
          }
          p.mu[row * D + d] = Nsum / Ssum;
          p.lv[row * D + d] = -logf(Ssum);
          p.S[row * D + d] = Ssum;
        }
      }
      b ^= 1;
      if (b == 0) aph ^= 1u;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kT5TmemCols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

// cudaErrorNotSupported when the shape / pointers are not covered (caller falls back to the
// mma.sync / slab-stream kernels).
cudaError_t tc5_encode(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* table, float* mu,
                       float* lv, float* S, cudaStream_t st) {
  const char* off = getenv("VIBO_DISABLE_TC5");
  if (off != nullptr && off[0] == '1') return cudaErrorNotSupported;
  const int I = d.num_item, D = d.ability_dim;
  if (!d.conditional || D > 5 || (I & 3) != 0 || I < 32 || I > 1024 || d.num_person < 1) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(resp) & 15) || (reinterpret_cast<uintptr_t>(mask) & 15)) return cudaErrorNotSupported;
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return cudaErrorNotSupported;
  T5Params p;
  p.P = d.num_person; p.I = I; p.D = D; p.missing_policy = d.missing_policy; p.n_kb = (I + kT5KB - 1) / kT5KB;
  p.resp = resp; p.mask = mask; p.table = table; p.mu = mu; p.lv = lv; p.S = S;
  const T5Smem L = t5_layout(p.n_kb);
  if (L.total > 227 * 1024) return cudaErrorNotSupported;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)I, (cuuint64_t)d.num_person};
  const cuuint64_t strides[1] = {(cuuint64_t)I * sizeof(float)};
  const cuuint32_t box[2] = {kT5KB, kT5Rows};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(resp), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorNotSupported;
  const int64_t n_tiles = (d.num_person + kT5Rows - 1) / kT5Rows;
  int grid = sm_count();
  if ((int64_t)grid > n_tiles) grid = (int)n_tiles;
  cudaError_t e = cudaFuncSetAttribute(tc5_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  if (e != cudaSuccess) return e;
  tc5_encode_kernel<<<grid, kT5Threads, L.total, st>>>(tmap, p);
  return cudaGetLastError();
}

}  // namespace vibo
