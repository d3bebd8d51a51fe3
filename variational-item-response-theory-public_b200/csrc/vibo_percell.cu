// Per-cell MLP of the nonlinear generative models on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM): the one GEMM-shaped hot op of this path
// (reference models.py:769-919 LinkedIRT / DeepIRT / ResidualIRT).
//
// For every response cell (person i, item j) the decoders evaluate
//     a1   = u_j + v_i + z_ij * w0                 first-layer pre-activation (rank-1 structured)
//     out  = w4 . ELU( W2 ELU(a1) + c2 ) + c4      hidden width H = 64
// (link: u = c0, v = 0, z = the IRT logit;  deep / residual: u_j = W_item h_item_j + c1,
// v_i = W_ability h_ability_i, no z term).  The reference materialises (P * I, 2H) inputs and runs
// cuBLAS / MKL GEMMs over them; here a tile of 128 cells is
//   1. BUILT on the CUDA cores: ELU(a1) for 128 cells x 64 features, split into bf16 hi + lo
//      (h = hi + lo to ~16 significant bits), written to shared memory in the UMMA canonical
//      K-major 128-byte-swizzled layout;
//   2. MULTIPLIED on the tensor cores: D[128 x 64] = A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T with
//      B = W2 (bf16 hi / lo, resident in shared memory), 12 tcgen05.mma (M 128, N 64, K 16) issued
//      by ONE thread, fp32 accumulation in TMEM, completion signalled by tcgen05.commit on an
//      mbarrier;
//   3. FINISHED on the CUDA cores: tcgen05.ld of the accumulator row, + c2, ELU, dot with w4.
// A cell tile is 8 persons x 16 items, so u / v rows are staged once per tile.
//
// Warp roles (288 threads): warps 0-7 build and finish (warp w owns TMEM lanes 32 (w % 4) .. + 31,
// i.e. cell rows of the tile, and the feature / accumulator-column half w / 4); warp 8 allocates
// TMEM and issues the MMAs.  Two CTAs per SM overlap each other's build / MMA / finish phases.
#include <cuda_bf16.h>

#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

namespace {

constexpr int kPcH = 64;             // hidden width this kernel is built for
constexpr int kPcTileP = 8, kPcTileI = 16, kPcTileM = kPcTileP * kPcTileI;   // 128 cells per tile
constexpr int kPcWorkers = 256, kPcThreads = kPcWorkers + 32;
constexpr int kPcULd = kPcH + 4;     // padded row of the staged u rows (conflict-free float4 reads)
constexpr uint32_t kPcTmemCols = 64;

struct PercellParams {
  int64_t P;
  int I;
  int u_rows, v_rows;   // I or 1 (broadcast row), P or 1
  const float* U;       // (u_rows, 64)
  const float* V;       // (v_rows, 64)
  const float* Z;       // (P, I) or null
  const float* w0;      // (64) or null
  const float* W2;      // (64, 64) row-major [out][in]
  const float* c2;      // (64)
  const float* w4;      // (64)
  float c4;
  float* out;           // (P, I)
};

// shared memory map (bytes); the operand tiles need 1024-byte alignment (128B swizzle atoms)
constexpr int kOffAhi = 0, kOffAlo = 16384, kOffBhi = 32768, kOffBlo = 40960;
constexpr int kOffU = 49152;                                   // [16][68] f32
constexpr int kOffV = kOffU + kPcTileI * kPcULd * 4;           // [8][64] f32
constexpr int kOffZ = kOffV + kPcTileP * kPcH * 4;             // [128] f32
constexpr int kOffPart = kOffZ + kPcTileM * 4;                 // [128] f32
constexpr int kOffVec = kOffPart + kPcTileM * 4;               // c2 | w4 | w0 : 3 x 64 f32
constexpr int kOffBar = kOffVec + 3 * kPcH * 4;                // 2 mbarriers + tmem base
constexpr int kPcSmem = kOffBar + 64;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, rows of 128 bytes packed
// densely (8-row groups 1024 bytes apart).  Field layout as in cute/arch/mma_sm100_desc.hpp:
// start address >> 4 [0,14), leading byte offset >> 4 [16,30), stride byte offset >> 4 [32,46),
// version = 1 [46,48), layout type SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D f32 [4,6) = 1, A bf16 [7,10) = 1, B bf16 [10,13) = 1, both
// K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float elu_fast(float a) { return a > 0.0f ? a : __expf(a) - 1.0f; }

// h -> bf16 hi and bf16 lo (h - hi), two values per 32-bit word
__device__ __forceinline__ void split2(float h0, float h1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(h0, h1);
  const float r0 = h0 - __low2float(b), r1 = h1 - __high2float(b);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  hi = *reinterpret_cast<const uint32_t*>(&b);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(kPcThreads, 2) percell_mlp_kernel(const __grid_constant__ PercellParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  float* s_u = reinterpret_cast<float*>(smem + kOffU);
  float* s_v = reinterpret_cast<float*>(smem + kOffV);
  float* s_z = reinterpret_cast<float*>(smem + kOffZ);
  float* s_part = reinterpret_cast<float*>(smem + kOffPart);
  float* s_c2 = reinterpret_cast<float*>(smem + kOffVec);
  float* s_w4 = s_c2 + kPcH;
  float* s_w0 = s_w4 + kPcH;
  const uint32_t bar_a = smem_addr(smem + kOffBar), bar_acc = bar_a + 8;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffBar + 16);
  const uint32_t a_hi = smem_addr(smem + kOffAhi), a_lo = smem_addr(smem + kOffAlo);
  const uint32_t b_hi = smem_addr(smem + kOffBhi), b_lo = smem_addr(smem + kOffBlo);

  // ---- one-time setup: barriers, TMEM, W2 (bf16 hi / lo, swizzled K-major), vectors ------------
  if (t == 0) {
    mbar_init(bar_a, kPcWorkers / 32);   // one arrive per worker warp
    mbar_init(bar_acc, 1);               // tcgen05.commit
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(s_tmem)),
                 "n"(kPcTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int k = t; k < kPcH * kPcH / 2; k += kPcThreads) {   // pairs (n, kk), (n, kk + 1)
    const int n = k / (kPcH / 2), kk = 2 * (k % (kPcH / 2));
    uint32_t hi, lo;
    split2(p.W2[n * kPcH + kk], p.W2[n * kPcH + kk + 1], hi, lo);
    const uint32_t off = (uint32_t)n * 128u + ((((uint32_t)kk >> 3) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)kk & 7u) * 2u;
    *reinterpret_cast<uint32_t*>(smem + kOffBhi + off) = hi;
    *reinterpret_cast<uint32_t*>(smem + kOffBlo + off) = lo;
  }
  if (t < kPcH) {
    s_c2[t] = p.c2[t];
    s_w4[t] = p.w4[t];
    s_w0[t] = p.w0 != nullptr ? p.w0[t] : 0.0f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // B tiles: generic writes -> tensor-core reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;

  const int n_it = (p.I + kPcTileI - 1) / kPcTileI;
  const int64_t n_pt = (p.P + kPcTileP - 1) / kPcTileP;
  const int64_t n_tiles = n_pt * n_it;
  uint32_t phase = 0;

  if (warp == 8) {
    // ===================== MMA issuer ====================================================
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(bar_a, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {   // hi hi, hi lo, lo hi
          const uint32_t a = prod == 2 ? a_lo : a_hi, b = prod == 1 ? b_lo : b_hi;
#pragma unroll
          for (int ks = 0; ks < kPcH / 16; ++ks)   // 32 bytes of K per instruction inside the swizzle atom
            umma_bf16(tmem, umma_desc_k_sw128(a + ks * 32), umma_desc_k_sw128(b + ks * 32),
                      (prod | ks) != 0 ? 1u : 0u);
        }
        umma_commit(bar_acc);   // arrives when the 12 MMAs have written TMEM (and read shared memory)
      }
      __syncwarp();
      phase ^= 1u;
    }
  } else {
    // ===================== build + finish warps ==========================================
    const int m = t & 127, half = t >> 7;            // cell row of the tile, feature half
    const int pi = m / kPcTileI, ji = m % kPcTileI;
    const uint32_t sw = (uint32_t)m & 7u;
    const uint32_t taddr = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)half * 32u;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t pt = tile / n_it;
      const int it = (int)(tile - pt * n_it);
      const int64_t i0 = pt * kPcTileP;
      const int j0 = it * kPcTileI;
      // ---- stage the tile's u rows, v rows and z values
      {
        const int r = t >> 4, c4 = t & 15;   // 16 rows x 16 float4
        int j = j0 + r;
        j = j < p.I ? j : p.I - 1;
        const float4 uv = *reinterpret_cast<const float4*>(p.U + (size_t)(p.u_rows == 1 ? 0 : j) * kPcH + c4 * 4);
        *reinterpret_cast<float4*>(s_u + r * kPcULd + c4 * 4) = uv;
        if (t < kPcTileP * 16) {
          int64_t i = i0 + r;
          i = i < p.P ? i : p.P - 1;
          const float4 vv = *reinterpret_cast<const float4*>(p.V + (size_t)(p.v_rows == 1 ? 0 : i) * kPcH + c4 * 4);
          *reinterpret_cast<float4*>(s_v + r * kPcH + c4 * 4) = vv;
        }
        if (t < kPcTileM) {
          int64_t i = i0 + pi;
          int j2 = j0 + ji;
          i = i < p.P ? i : p.P - 1;
          j2 = j2 < p.I ? j2 : p.I - 1;
          s_z[t] = p.Z != nullptr ? p.Z[i * p.I + j2] : 0.0f;
        }
      }
      worker_barrier();
      // ---- build A: ELU(u_j + v_i + z w0) for this thread's 32 features, bf16 hi / lo
      {
        const float z = s_z[m];
        const float* ur = s_u + ji * kPcULd + half * 32;
        const float* vr = s_v + pi * kPcH + half * 32;
        const float* wr = s_w0 + half * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // 16-byte chunks of 8 bf16
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 u4 = *reinterpret_cast<const float4*>(ur + c * 8 + q * 4);
            const float4 v4 = *reinterpret_cast<const float4*>(vr + c * 8 + q * 4);
            const float4 w4 = *reinterpret_cast<const float4*>(wr + c * 8 + q * 4);
            const float h0 = elu_fast(fmaf(z, w4.x, u4.x + v4.x)), h1 = elu_fast(fmaf(z, w4.y, u4.y + v4.y));
            const float h2 = elu_fast(fmaf(z, w4.z, u4.z + v4.z)), h3 = elu_fast(fmaf(z, w4.w, u4.w + v4.w));
            split2(h0, h1, hi[2 * q], lo[2 * q]);
            split2(h2, h3, hi[2 * q + 1], lo[2 * q + 1]);
          }
          const uint32_t off = (uint32_t)m * 128u + ((((uint32_t)(half * 4 + c)) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(hi[0]), "r"(hi[1]),
                       "r"(hi[2]), "r"(hi[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(lo[0]), "r"(lo[1]),
                       "r"(lo[2]), "r"(lo[3])
                       : "memory");
        }
      }
      // generic-proxy writes -> async-proxy (tensor core) reads; order the earlier tcgen05.ld of this
      // thread before the MMA that will overwrite the accumulator
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a);
      // ---- finish: accumulator row -> + c2, ELU, dot w4
      mbar_wait(bar_acc, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32];
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
      float acc = 0.0f;
#pragma unroll
      for (int n = 0; n < 32; ++n)
        acc = fmaf(s_w4[half * 32 + n], elu_fast(__uint_as_float(r[n]) + s_c2[half * 32 + n]), acc);
      if (half == 1) s_part[m] = acc;
      worker_barrier();   // partial sums visible; staging buffers free for the next tile
      if (half == 0) {
        const int64_t i = i0 + pi;
        const int j = j0 + ji;
        if (i < p.P && j < p.I) p.out[i * p.I + j] = acc + s_part[m] + p.c4;
      }
      phase ^= 1u;
    }
  }

  // ---- teardown: every tcgen05 operation of this CTA is complete before TMEM is released
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kPcTmemCols) : "memory");
  }
}

}  // namespace

cudaError_t launch_percell_mlp(int64_t P, int I, int u_rows, int v_rows, const float* U, const float* V,
                               const float* Z, const float* w0, const float* W2, const float* c2, const float* w4,
                               float c4, float* out, cudaStream_t st) {
  PercellParams p;
  p.P = P; p.I = I; p.u_rows = u_rows; p.v_rows = v_rows; p.U = U; p.V = V; p.Z = Z; p.w0 = w0; p.W2 = W2;
  p.c2 = c2; p.w4 = w4; p.c4 = c4; p.out = out;
  const int64_t n_tiles = ((P + kPcTileP - 1) / kPcTileP) * ((I + kPcTileI - 1) / kPcTileI);
  int64_t grid = (int64_t)sm_count() * 2;
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(percell_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPcSmem);
  if (e != cudaSuccess) return e;
  percell_mlp_kernel<<<(int)grid, kPcThreads, kPcSmem, st>>>(p);
  note_launch();
  return cudaGetLastError();
}

}  // namespace vibo
