// Backward of the conditional-posterior encode on the 5th-generation tensor cores (tcgen05.mma,
// TMEM, 2-D TMA): the companion of vibo_tc5_encode.cu.
//
// The gradient of the product-of-experts sums with respect to the expert table is
//     A^r_j[d] = sum_i [o_ij, x_ij = r] GN_i[d],    B^r_j[d] likewise with GS_i,
//     GN = g_mu / S,   GS = -(g_mu mu + g_lv) / S                       (SURVEY.md Appendix A "PoE"),
// which for fully observed persons is the TRANSPOSED product  C1 = X^T G  (items x persons times
// persons x 2D) plus a column sum:  A^1 | B^1 = C1,   A^0 | B^0 = sum_i G_i - C1.
//   * A operand = X^T: the same 2-D TMA boxes of the float32 response matrix as the forward kernel
//     (32 items x 32 persons, 128-byte swizzle with 32-byte atoms); the rows of a box are persons = the K dimension, so
//     the box IS the UMMA canonical MN-major layout (no transposition pass), exact in TF32;
//   * B operand = [GN | GS], computed by the CUDA cores for 128 persons at a time (four 32-person K
//     blocks; one pass hides its global-memory latency behind ~13 us of streaming), split into three
//     TF32 terms (6D <= 30 columns), written K-major / swizzled, double buffered;
//   * D = 8 item tiles x (128 items x 32 columns) fp32 accumulators resident in TMEM (256 columns) for
//     the WHOLE kernel: tcgen05.mma.kind::tf32 (M = 128 items, N = 32, K = 8 persons), one thread
//     issues; a CTA walks its person chunks once and emits one partial at the end.
// Persons with a missing cell are taken out of the tensor-core operand (their G columns are zeroed)
// and accumulated exactly by one warp in row order (deterministic), so any mask is handled.
// Output: the per-CTA partials part[cta][r][j][A(D) | B(D)] that encode_bwd_finalize_kernel sums.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>

#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

namespace {

constexpr int kB5Items = 128, kB5KP = 32, kB5N = 32, kB5Stages = 6, kB5Threads = 192;
constexpr int kB5SP = 128, kB5Sub = kB5SP / kB5KP;   // persons per builder pass ("super-chunk") = 4 TMA / MMA chunks of 32
constexpr uint32_t kB5BoxBytes = 32 * kB5KP * 4;          // 32 items x 32 persons: 4 KB
constexpr uint32_t kB5StageBytes = 4 * kB5BoxBytes;       // 128 items x 32 persons: 16 KB
constexpr uint32_t kB5BTile = kB5N * kB5KP * 4;           // 4 KB
constexpr uint32_t kB5TmemCols = 256;

struct B5Params {
  int64_t P;
  int I, D, n_it;
  const float* resp;
  const uint8_t* mask;
  const float* amu;
  const float* S;
  const float* g_mu;
  const float* g_lv;
  float* part;   // [grid][2][I][2D]
};

// shared memory (bytes): stages | B tiles (2 x 4) | G staging [128][2D] | Gsum [2D] | flags [128] | barriers
constexpr uint32_t kB5OffStage = 0;
constexpr uint32_t kB5OffB = kB5Stages * kB5StageBytes;
constexpr uint32_t kB5OffG = kB5OffB + 2 * kB5Sub * kB5BTile;
constexpr uint32_t kB5OffSum = kB5OffG + kB5SP * 2 * 5 * 4;   // D <= 5
constexpr uint32_t kB5OffFlag = kB5OffSum + 2 * VIBO_MAX_ABILITY_DIM * 4;
constexpr uint32_t kB5OffBar = kB5OffFlag + kB5SP;
constexpr uint32_t kB5Smem = kB5OffBar + 8 * (2 * kB5Stages + 5) + 16;

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
      "B5_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra B5_DONE;\n"
      "bra B5_WAIT;\n"
      "B5_DONE:\n"
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
// B: K-major, 128-byte swizzle, 128-byte rows packed densely (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// A: MN-major TF32.  For 32-bit MN-major operands the only UMMA shared-memory layout is the 128-byte
// swizzle with a 32-BYTE base (layout type SWIZZLE_128B_BASE32B = 1; Swizzle<2,5,2> on byte addresses:
// the 32-byte chunk index of a 128-byte line is XORed with the line index mod 4), which is what TMA
// writes in mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  One K row (person) is a 128-byte line of 32
// consecutive MN elements (items); 4 K rows make a 512-byte swizzle atom.  Leading byte offset =
// distance between the 32-item blocks of the M = 128 tile (one TMA box each: 4096 B), stride byte
// offset = distance between 4-person atoms (512 B)   [cute/atom/mma_traits_sm100.hpp, Major::MN,
// Layout_MN_SW128_32B_Atom; cutlass/gemm/collective/builders/sm100_common.inl:92].
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)(kB5BoxBytes >> 4) << 16) | (32ull << 32) | (1ull << 46) |
         (1ull << 61);
}
// kind::tf32, D f32, A MN-major (bit 15), B K-major, N = 32, M = 128
constexpr uint32_t kB5Idesc =
    (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kB5Idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(kB5Threads, 1)
tc5_encode_bwd_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ B5Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int I = p.I, D = p.D, D2 = 2 * D;
  float* s_g = reinterpret_cast<float*>(smem + kB5OffG);       // [32 persons][2D]: GN | GS (0 for flagged persons)
  float* s_sum = reinterpret_cast<float*>(smem + kB5OffSum);   // [2D] column sums over unflagged persons
  uint8_t* s_flag = smem + kB5OffFlag;                         // [32]
  const uint32_t bars = saddr(smem + kB5OffBar);
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kB5Stages + s); };
  auto b_full = [&](int b) { return bars + 8u * (2 * kB5Stages + b); };
  auto b_empty = [&](int b) { return bars + 8u * (2 * kB5Stages + 2 + b); };
  const uint32_t acc_full = bars + 8u * (2 * kB5Stages + 4);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kB5OffBar + 8 * (2 * kB5Stages + 5));
  const uint32_t st_base = saddr(smem + kB5OffStage), b_base = saddr(smem + kB5OffB);
  float* my_part = p.part + (size_t)blockIdx.x * 2 * I * D2;

  if (t == 0) {
    for (int s = 0; s < kB5Stages; ++s) {
      bar_init(full(s), 1);
      bar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      bar_init(b_full(b), 4);    // one arrive per builder warp
      bar_init(b_empty(b), 1);   // tcgen05.commit
    }
    bar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(s_tmem)),
                 "n"(kB5TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B tiles: zero once (rows n >= 6D stay zero); the partial of this CTA starts at zero (exact rows add into it)
  for (uint32_t k = t; k < 2 * kB5Sub * kB5BTile / 16; k += kB5Threads)
    reinterpret_cast<uint4*>(smem + kB5OffB)[k] = make_uint4(0, 0, 0, 0);
  for (int k = t; k < 2 * I * D2; k += kB5Threads) my_part[k] = 0.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;
  const int64_t n_chunks = (p.P + kB5SP - 1) / kB5SP;   // super-chunks of 128 persons

  if (warp == 0) {
    // ===================== TMA producer ===================================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int64_t left = p.P - c * kB5SP;
        const int n_sub = left >= kB5SP ? kB5Sub : (int)((left + kB5KP - 1) / kB5KP);
        for (int sub = 0; sub < n_sub; ++sub) {
          for (int it = 0; it < p.n_it; ++it) {
            bar_wait(empty(s), ph ^ 1u);
            bar_expect_tx(full(s), kB5StageBytes);
            const uint32_t dst = st_base + (uint32_t)s * kB5StageBytes;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tma_load_2d(dst + q * kB5BoxBytes, &tmap, full(s), it * kB5Items + q * 32,
                          (int)(c * kB5SP + sub * kB5KP));
            if (++s == kB5Stages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================================================
    int s = 0, bb = 0;
    uint32_t ph = 0, bph = 0;
    bool first = true;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      const int64_t left = p.P - c * kB5SP;
      const int n_sub = left >= kB5SP ? kB5Sub : (int)((left + kB5KP - 1) / kB5KP);
      bar_wait(b_full(bb), bph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int sub = 0; sub < n_sub; ++sub) {
        for (int it = 0; it < p.n_it; ++it) {
          bar_wait(full(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (lane == 0) {
            const uint32_t a = st_base + (uint32_t)s * kB5StageBytes;
            const uint32_t b = b_base + (uint32_t)(bb * kB5Sub + sub) * kB5BTile;
#pragma unroll
            for (int ks = 0; ks < kB5KP / 8; ++ks)   // 8 persons per instruction: two 512-byte swizzle atoms of A
              umma_tf32(tmem + (uint32_t)it * kB5N, desc_mn_sw128(a + ks * 1024), desc_k_sw128(b + ks * 32),
                        (first && sub == 0 && ks == 0) ? 0u : 1u);
            umma_commit(empty(s));
            if (sub == n_sub - 1 && it == p.n_it - 1) umma_commit(b_empty(bb));
          }
          __syncwarp();
          if (++s == kB5Stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
      first = false;
      bb ^= 1;
      if (bb == 0) bph ^= 1u;
    }
    if (lane == 0) umma_commit(acc_full);
    __syncwarp();
  } else {
    // ===================== G builder (per chunk) + epilogue (once) =============================
    const int e = t - 64;   // 0..127
    int bb = 0;
    uint32_t bph = 0;
    double gsum = 0.0;      // threads e < 2D: column sum of G over this CTA's unflagged persons
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      const int64_t row0 = c * kB5SP;
      const int rows = (int)((p.P - row0 < kB5SP) ? p.P - row0 : kB5SP);
      s_flag[e] = 0;
      asm volatile("bar.sync 2, 128;" ::: "memory");
      {   // persons with a missing cell (contiguous mask block of the chunk)
        const uint8_t* mb = p.mask + row0 * I;
        const int64_t len = (int64_t)rows * I, n16 = len >> 4;
        const uint4* m16 = reinterpret_cast<const uint4*>(mb);
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
      // G of the super-chunk: one thread per person
      bar_wait(b_empty(bb), bph ^ 1u);   // the MMAs of two super-chunks ago are done with these B tiles
      {
        const int k = e;
        const int64_t row = row0 + k;
        const bool live = k < rows;
        const bool flagged = live && s_flag[k] != 0;
        const int sub = k >> 5, kk = k & 31;
        for (int cc = 0; cc < D2; ++cc) {
          const int d = cc < D ? cc : cc - D;
          float g = 0.0f;
          if (live) {
            const float sv = p.S[row * D + d], gm = p.g_mu[row * D + d];
            g = cc < D ? gm / sv : -(gm * p.amu[row * D + d] + p.g_lv[row * D + d]) / sv;
          }
          s_g[k * D2 + cc] = g;   // exact value, also for flagged persons (used by the exact pass)
          const float gt = flagged ? 0.0f : g;
          const float hi = tf32_trunc(gt), mid = tf32_trunc(gt - hi), lo = tf32_trunc(gt - hi - mid);
          const float parts[3] = {hi, mid, lo};
#pragma unroll
          for (int s3 = 0; s3 < 3; ++s3) {
            const int n = s3 * D2 + cc;
            const uint32_t off = (uint32_t)(bb * kB5Sub + sub) * kB5BTile + (uint32_t)n * 128u +
                                 ((((uint32_t)kk >> 2) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)kk & 3u) * 4u;
            *reinterpret_cast<float*>(smem + kB5OffB + off) = parts[s3];
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(b_full(bb));
      asm volatile("bar.sync 2, 128;" ::: "memory");   // s_g complete
      if (e < D2) {
        for (int k = 0; k < rows; ++k)
          if (s_flag[k] == 0) gsum += (double)s_g[k * D2 + e];
      }
      // exact pass for flagged persons: one warp, persons in order, lanes over items (deterministic)
      if (warp == 2) {
        for (int k = 0; k < rows; ++k) {
          if (s_flag[k] == 0) continue;
          const int64_t row = row0 + k;
          for (int j = lane; j < I; j += 32) {
            if (p.mask[row * I + j] == 0) continue;
            const int x = p.resp[row * I + j] > 0.5f ? 1 : 0;
            float* dst = my_part + ((size_t)x * I + j) * D2;
            for (int cc = 0; cc < D2; ++cc) dst[cc] += s_g[k * D2 + cc];
          }
        }
      }
      bb ^= 1;
      if (bb == 0) bph ^= 1u;
    }
    // ---- epilogue: C1 from TMEM, A^1 | B^1 = C1, A^0 | B^0 = column sum - C1, added to the exact part
    if (e < D2) s_sum[e] = (float)gsum;
    asm volatile("bar.sync 2, 128;" ::: "memory");
    bar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int it = 0; it < p.n_it; ++it) {
      uint32_t r[32];
      const uint32_t taddr = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)it * kB5N;
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
      const int j = it * kB5Items + 32 * (warp & 3) + lane;
      if (j < I) {
        for (int cc = 0; cc < D2; ++cc) {
          const float c1 = __uint_as_float(r[cc]) + __uint_as_float(r[D2 + cc]) + __uint_as_float(r[2 * D2 + cc]);
          my_part[((size_t)I + j) * D2 + cc] += c1;
          my_part[(size_t)j * D2 + cc] += s_sum[cc] - c1;
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kB5TmemCols) : "memory");
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
// mma.sync / slab-stream kernels).  *grid_out = number of per-CTA partials written to `part`.
cudaError_t tc5_encode_bwd(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* amu,
                           const float* S, const float* g_mu, const float* g_lv, float* part, int* grid_out,
                           cudaStream_t st) {
  const char* off = getenv("VIBO_DISABLE_TC5");
  if (off != nullptr && off[0] == '1') return cudaErrorNotSupported;
  const int I = d.num_item, D = d.ability_dim;
  if (!d.conditional || D > 5 || (I & 3) != 0 || I < 32 || I > 1024 || d.num_person < 1) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(resp) & 15) || (reinterpret_cast<uintptr_t>(mask) & 15)) return cudaErrorNotSupported;
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return cudaErrorNotSupported;
  B5Params p;
  p.P = d.num_person; p.I = I; p.D = D; p.n_it = (I + kB5Items - 1) / kB5Items;
  p.resp = resp; p.mask = mask; p.amu = amu; p.S = S; p.g_mu = g_mu; p.g_lv = g_lv; p.part = part;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)I, (cuuint64_t)d.num_person};
  const cuuint64_t strides[1] = {(cuuint64_t)I * sizeof(float)};
  const cuuint32_t box[2] = {32, kB5KP};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(resp), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorNotSupported;
  const int64_t n_chunks = (d.num_person + kB5SP - 1) / kB5SP;
  int grid = sm_count();
  if ((int64_t)grid > n_chunks) grid = (int)n_chunks;
  cudaError_t e = cudaFuncSetAttribute(tc5_encode_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kB5Smem);
  if (e != cudaSuccess) return e;
  tc5_encode_bwd_kernel<<<grid, kB5Threads, kB5Smem, st>>>(tmap, p);
  if (grid_out) *grid_out = grid;
  return cudaGetLastError();
}

}  // namespace vibo
