// Single-pass ELBO evaluation for the CONDITIONAL posterior (C3's shape) on tcgen05 / TMEM / 2-D TMA.
//
// One read of the response matrix computes, per person, the product-of-experts posterior
// (models.py:664-710 + :596-629, utils.py:105-113), the reparameterised draw (:506-510), the IRT link
// (:729-766), the masked Bernoulli log-likelihood (utils.py:46-49) and the person-side prior term
// (utils.py:85-88 / models.py:433-435): what round 1 did in two passes (tensor-core encode, then a
// slab-stream link kernel that re-read the matrix).
//
// A CTA walks tiles of 128 persons.  While tile t streams through the tensor cores exactly as in
// vibo_tc5_encode.cu (TMA boxes of the float32 response matrix -> tcgen05.mma.kind::tf32 against the
// resident three-term TF32 table -> S, N in TMEM), four "packer" warps turn every shared-memory stage
// into ONE BIT per cell (warp ballot of x > 1/2: 16 KB per tile instead of 512 KB), so the tile stays
// on chip.  The epilogue warps form the posterior, draw theta (Philox keyed by the global person index,
// or supplied noise), add the person-side term and publish theta; sixteen "link" warps then score
// tile t from the bits while tile t + 1 is already streaming: warp w owns item blocks w and w + 16 with
// the item parameters in registers, lane = item, loop over the 128 persons (theta broadcast from
// shared memory).  Persons with a missing cell are flagged by the mask scan and handled exactly by
// their epilogue thread (posterior and log-likelihood), so the kernel is correct for any mask.
//
// Warps: 0 TMA producer, 1 TMEM allocator + MMA issuer, 2-5 mask scan + epilogue, 6-9 bit packers,
// 10-25 link.  Bound: HBM (4 B/cell + 1 B/cell mask scan), with ~20 CUDA-core instructions per cell of
// link arithmetic overlapped underneath.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>

#include "vibo_stream_kernel.cuh"

namespace vibo {

namespace {

constexpr int kE5Rows = 128, kE5KB = 32, kE5N = 32, kE5Stages = 3;
constexpr int kE5PackWarps = 4, kE5LinkWarps = 16, kE5Threads = (2 + 4 + kE5PackWarps + kE5LinkWarps) * 32;   // 832
constexpr uint32_t kE5StageBytes = kE5Rows * kE5KB * 4;
constexpr uint32_t kE5BBlock = kE5N * kE5KB * 4;
constexpr uint32_t kE5TmemCols = 64;
constexpr int kE5ThLd = 8;   // floats per person in the theta tile

struct E5Params {
  int64_t P;
  int I, n_kb, missing_policy, form;
  int debug;   // VIBO_E5_DEBUG (tool runs only): 1 skip the link arithmetic, 2 skip the bit packing, 4 skip the mask scan;
               // 16: packers hand a stage back after the ballots; 128: no hardening of the hand-offs (A/B runs);
               // 8: two 768-thread barriers per tile put the epilogue / packer warps and the link warps in lock-step --
               // racecheck does not credit an mbarrier arrive / wait pair as ordering ordinary shared-memory accesses
  int64_t person_offset;
  const float* resp;
  const uint8_t* mask;
  const float* table;      // (2, I, 2D)
  const float* item_feat;  // (I, F)
  const float* eps;        // (P, D) or null
  uint64_t seed;
  const uint64_t* seed_dev;
  float* out_mu;
  float* out_lv;
  float* out_theta;
  double* part;            // [grid][2]: LL, person term
};

struct E5Smem {
  uint32_t b_off, stage_off, bits_off, theta_off, flag_off, base_off, red_off, bar_off, total;
};
__host__ __device__ inline E5Smem e5_layout(int n_kb) {
  E5Smem L;
  L.b_off = 0;
  L.stage_off = ((uint32_t)n_kb * kE5BBlock + 1023u) / 1024u * 1024u;
  L.bits_off = L.stage_off + kE5Stages * kE5StageBytes;
  L.theta_off = L.bits_off + 2 * kE5Rows * 32 * 4;
  L.flag_off = L.theta_off + 2 * kE5Rows * kE5ThLd * 4;
  L.base_off = L.flag_off + 2 * kE5Rows;
  L.red_off = L.base_off + 64;
  L.bar_off = L.red_off + 8 * 2 * (kE5Threads / 32);
  L.total = L.bar_off + 8 * (2 * kE5Stages + 8) + 16;
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
      "E5_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra E5_DONE;\n"
      "bra E5_WAIT;\n"
      "E5_DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// The same wait for the roles that wait LONG (a whole tile): back off with nanosleep between polls.  A polling
// warp issues three instructions per trip; sixteen link warps polling for the next tile took a quarter of the
// SM's issue slots away from the packer / epilogue warps they were waiting for.
__device__ __forceinline__ void bar_wait_idle(uint32_t bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(128);
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
constexpr uint32_t kE5Idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kE5Idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Log-likelihood of TWO cells (the warp's two item blocks for one person) in packed f32x2 arithmetic,
// forward only; same clamp semantics as link_cell / cell_logistic_fast (vibo_stream_kernel.cuh): the
// probability is clamped to [eps32, 1 - eps32] (utils.py:46-49 -> torch Bernoulli).  x0 / x1: the cell is 1.
// 3PL: the probabilities of the two observed responses, clamped to [eps32, 1 - eps32] (the value the reference
// takes the logarithm of).  The caller multiplies the probabilities of up to four persons before ONE lg2 per
// item: four factors >= 1.19e-7 stay far above the smallest normal float.
__device__ __forceinline__ f2_t prob_pair_3pl(f2_t z2, bool x0, bool x1, f2_t g2, f2_t omg2) {
  float z0, z1;
  unpack2(z2, z0, z1);
  const f2_t zc2 = pack2(fmaxf(z0, -80.0f), fmaxf(z1, -80.0f));
  float t0, t1;
  unpack2(mul2(zc2, pack2(-kLog2e, -kLog2e)), t0, t1);
  const f2_t e2 = pack2(ex2_approx(t0), ex2_approx(t1));
  float w0, w1;
  unpack2(add2(e2, pack2(1.0f, 1.0f)), w0, w1);
  const f2_t r2 = pack2(rcp_approx(w0), rcp_approx(w1));
  const f2_t pp2 = fma2(omg2, r2, g2);            // p = g + (1 - g) sigmoid(z)
  const f2_t q2 = mul2(omg2, mul2(e2, r2));       // 1 - p = (1 - g) sigmoid(-z), full relative precision
  float p0, p1, q0, q1;
  unpack2(pp2, p0, p1);
  unpack2(q2, q0, q1);
  return pack2(fminf(fmaxf(x0 ? p0 : q0, kEps32), 1.0f - kEps32), fminf(fmaxf(x1 ? p1 : q1, kEps32), 1.0f - kEps32));
}

// Returns (ll0, ll1) in log2 units for the 3PL (caller scales by ln 2) / natural units for 1PL / 2PL.
template <int MODEL>
__device__ __forceinline__ f2_t eval_pair(f2_t z2, bool x0, bool x1, f2_t g2, f2_t omg2) {
  float z0, z1;
  unpack2(z2, z0, z1);
  if (MODEL == 3) {
    const f2_t zc2 = pack2(fmaxf(z0, -80.0f), fmaxf(z1, -80.0f));
    float t0, t1;
    unpack2(mul2(zc2, pack2(-kLog2e, -kLog2e)), t0, t1);
    const f2_t e2 = pack2(ex2_approx(t0), ex2_approx(t1));
    float w0, w1;
    unpack2(add2(e2, pack2(1.0f, 1.0f)), w0, w1);
    const f2_t r2 = pack2(rcp_approx(w0), rcp_approx(w1));
    const f2_t pp2 = fma2(omg2, r2, g2);            // p = g + (1 - g) sigmoid(z)
    const f2_t q2 = mul2(omg2, mul2(e2, r2));       // 1 - p = (1 - g) sigmoid(-z), full relative precision
    float p0, p1, q0, q1;
    unpack2(pp2, p0, p1);
    unpack2(q2, q0, q1);
    const float u0 = fminf(fmaxf(x0 ? p0 : q0, kEps32), 1.0f - kEps32);
    const float u1 = fminf(fmaxf(x1 ? p1 : q1, kEps32), 1.0f - kEps32);
    return pack2(lg2_approx(u0), lg2_approx(u1));
  } else {
    // ll = (x - 1) zc - log(1 + exp(-zc)),  zc = clamp(z, +-15.942385)
    const f2_t zc2 = pack2(fminf(fmaxf(z0, -kLogitClamp), kLogitClamp), fminf(fmaxf(z1, -kLogitClamp), kLogitClamp));
    float t0, t1;
    unpack2(mul2(zc2, pack2(-kLog2e, -kLog2e)), t0, t1);
    const f2_t l2 = pack2(lg2_approx(1.0f + ex2_approx(t0)), lg2_approx(1.0f + ex2_approx(t1)));
    const f2_t xm2 = pack2(x0 ? 0.0f : -1.0f, x1 ? 0.0f : -1.0f);
    return fma2(l2, pack2(-kLn2f, -kLn2f), mul2(xm2, zc2));
  }
}

template <int MODEL, int D>
__global__ void __launch_bounds__(kE5Threads, 1)
tc5_eval_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ E5Params p) {
  constexpr int F = item_width(MODEL, D), D2 = 2 * D;
  extern __shared__ __align__(1024) unsigned char smem[];
  const E5Smem L = e5_layout(p.n_kb);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int I = p.I;
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem + L.bits_off);   // [2][32 item blocks][128 persons]
  float* s_theta = reinterpret_cast<float*>(smem + L.theta_off);      // [2][128][8]
  uint8_t* s_flags = smem + L.flag_off;                               // [2][128]
  float* s_base = reinterpret_cast<float*>(smem + L.base_off);        // base_S[D] | base_N[D]
  double* s_red = reinterpret_cast<double*>(smem + L.red_off);        // [warps][2]
  const uint32_t bars = saddr(smem + L.bar_off);
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kE5Stages + s); };
  auto acc_full = [&](int b) { return bars + 8u * (2 * kE5Stages + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (2 * kE5Stages + 2 + b); };
  auto theta_full = [&](int b) { return bars + 8u * (2 * kE5Stages + 4 + b); };
  auto tile_done = [&](int b) { return bars + 8u * (2 * kE5Stages + 6 + b); };
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * (2 * kE5Stages + 8));
  const uint32_t b_base = saddr(smem + L.b_off), st_base = saddr(smem + L.stage_off);

  // ---- setup ------------------------------------------------------------------------------------
  if (t == 0) {
    for (int s = 0; s < kE5Stages; ++s) {
      bar_init(full(s), 1);
      bar_init(empty(s), 1 + kE5PackWarps);   // tcgen05.commit + the packer warps
    }
    for (int b = 0; b < 2; ++b) {
      bar_init(acc_full(b), 1);
      bar_init(acc_empty(b), 4);    // epilogue warps
      bar_init(theta_full(b), 4 + kE5PackWarps);   // epilogue warps (theta, flags) + packer warps (bits)
      bar_init(tile_done(b), kE5LinkWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(s_tmem)),
                 "n"(kE5TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t k = t; k < (uint32_t)p.n_kb * kE5BBlock / 16; k += kE5Threads)
    reinterpret_cast<uint4*>(smem + L.b_off)[k] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int idx = t; idx < I * D2; idx += kE5Threads) {
    const int j = idx / D2, c = idx % D2, d = c < D ? c : c - D;
    const float m0 = p.table[(size_t)j * D2 + d], l0 = p.table[(size_t)j * D2 + D + d];
    const float m1 = p.table[(size_t)(I + j) * D2 + d], l1 = p.table[(size_t)(I + j) * D2 + D + d];
    const float t0 = 1.0f / (expf(l0) + kPoeEps), t1 = 1.0f / (expf(l1) + kPoeEps);
    const float diff = c < D ? t1 - t0 : m1 * t1 - m0 * t0;
    const float hi = tf32_trunc(diff), mid = tf32_trunc(diff - hi), lo = tf32_trunc(diff - hi - mid);
    const int kb = j / kE5KB, kk = j % kE5KB;
    const float parts[3] = {hi, mid, lo};
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int n = s * D2 + c;
      const uint32_t off = (uint32_t)kb * kE5BBlock + (uint32_t)n * 128u +
                           ((((uint32_t)kk >> 2) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)kk & 3u) * 4u;
      *reinterpret_cast<float*>(smem + L.b_off + off) = parts[s];
    }
  }
  for (int c = warp; c < D2; c += kE5Threads / 32) {
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
  const int64_t n_tiles = (p.P + kE5Rows - 1) / kE5Rows;
  double acc_ll = 0.0, acc_term = 0.0;   // per-thread partial results

  if (warp == 0) {
    // ===================== TMA producer ===================================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.n_kb; ++kb) {
          bar_wait(empty(s), ph ^ 1u);
          bar_expect_tx(full(s), kE5StageBytes);
          tma_load_2d(st_base + (uint32_t)s * kE5StageBytes, &tmap, full(s), kb * kE5KB, (int)(tile * kE5Rows));
          if (++s == kE5Stages) {
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
          const uint32_t a = st_base + (uint32_t)s * kE5StageBytes, bb = b_base + (uint32_t)kb * kE5BBlock;
#pragma unroll
          for (int ks = 0; ks < kE5KB / 8; ++ks)
            umma_tf32(tmem + (uint32_t)b * kE5N, desc_k_sw128(a + ks * 32), desc_k_sw128(bb + ks * 32),
                      (kb | ks) != 0 ? 1u : 0u);
          umma_commit(empty(s));
          if (kb == p.n_kb - 1) umma_commit(acc_full(b));
        }
        __syncwarp();
        if (++s == kE5Stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      b ^= 1;
      if (b == 0) aph ^= 1u;
    }
  } else if (warp < 6) {
    // ===================== mask scan + epilogue (one thread per person of the tile) ==========
    const int e = t - 64;
    const int m = 32 * (warp & 3) + lane;
    const float prior_tau = p.missing_policy == VIBO_MISSING_PRIOR ? 1.0f / (1.0f + kPoeEps) : 0.0f;
    const uint64_t key = p.seed_dev != nullptr ? p.seed_dev[0] + p.seed_dev[1] : p.seed;
    int b = 0;
    uint32_t aph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row0 = tile * kE5Rows;
      const int rows = (int)((p.P - row0 < kE5Rows) ? p.P - row0 : kE5Rows);
      uint8_t* s_flag = s_flags + b * kE5Rows;
      float* th_tile = s_theta + b * kE5Rows * kE5ThLd;
      bar_wait_idle(tile_done(b), aph ^ 1u);   // the link warps are done with this buffer (two tiles ago)
      s_flag[e] = 0;
      asm volatile("bar.sync 2, 128;" ::: "memory");
      {
        const uint8_t* mb = p.mask + row0 * I;
        const int64_t len = (int64_t)rows * I, n16 = len >> 4;
        const uint4* m16 = reinterpret_cast<const uint4*>(mb);
        constexpr int B = 16;
        for (int64_t k0 = e; k0 < ((p.debug & 4) ? 0 : n16); k0 += 128 * B) {
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
      const bool exact = m < rows && s_flag[m] != 0;
      bar_wait_idle(acc_full(b), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32];
      const uint32_t taddr = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)b * kE5N;
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
      if (lane == 0) bar_arrive(acc_empty(b));
      if (m < rows) {
        const int64_t row = row0 + m;
        float th[D], tsum = 0.0f;
        float nrm[4];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float Ssum, Nsum;
          if (!exact) {
            Ssum = s_base[d] + (__uint_as_float(r[d]) + __uint_as_float(r[D2 + d]) + __uint_as_float(r[2 * D2 + d]));
            Nsum = s_base[D + d] +
                   (__uint_as_float(r[D + d]) + __uint_as_float(r[D2 + D + d]) + __uint_as_float(r[2 * D2 + D + d]));
          } else {
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
          const float amu = Nsum / Ssum, alv = -logf(Ssum), sd = rsqrtf(Ssum);
          float ev;
          if (p.eps != nullptr) {
            ev = p.eps[row * D + d];
          } else {
            if ((d & 3) == 0) philox_normal4(key, (uint64_t)(p.person_offset + row), (uint32_t)(d >> 2), nrm);
            ev = nrm[d & 3];
          }
          th[d] = fmaf(ev, sd, amu);
          tsum += th[d];
          if (p.form == VIBO_ELBO_KL) {
            acc_term += (double)(-0.5f * (1.0f + alv - amu * amu - 1.0f / Ssum));
          } else {
            acc_term += (double)(-0.5f * th[d] * th[d] + 0.5f * ev * ev + 0.5f * alv);
          }
          th_tile[m * kE5ThLd + d] = th[d];
          if (p.out_mu != nullptr) {
            p.out_mu[row * D + d] = amu;
            p.out_lv[row * D + d] = alv;
            p.out_theta[row * D + d] = th[d];
          }
        }
        if (MODEL == 1) th_tile[m * kE5ThLd + 7] = tsum;
        if (exact) {
          // log-likelihood of a person with missing cells, on the spot (the link warps skip flagged persons)
          float ll = 0.0f;
          for (int j = 0; j < I; ++j) {
            if (p.mask[row * I + j] == 0) continue;
            float z;
            if (MODEL == 1) {
              z = tsum + p.item_feat[j];
            } else {
              z = p.item_feat[(size_t)j * F + D];
#pragma unroll
              for (int d = 0; d < D; ++d) z = fmaf(-th[d], p.item_feat[(size_t)j * F + d], z);
            }
            const float g = MODEL == 3 ? 1.0f / (1.0f + expf(-p.item_feat[(size_t)j * F + D + 1])) : 0.0f;
            float l1, dz, t0;
            link_cell<MODEL>(z, p.resp[row * I + j], g, 1.0f - g, 1.0f, l1, dz, t0);
            ll += l1;
          }
          acc_ll += (double)ll;
        }
      } else {
        s_flag[m] = 1;   // rows past the end of the matrix: the link warps skip them
      }
      if (!(p.debug & 128)) __threadfence_block();   // theta / flags are performed before the arrive below
      __syncwarp();
      if (lane == 0) bar_arrive(theta_full(b));   // release: theta, flags of this tile
      if (p.debug & 8) {   // lock-step with the link warps (racecheck runs, see E5Params::debug)
        asm volatile("bar.sync 3, 768;" ::: "memory");
        asm volatile("bar.sync 3, 768;" ::: "memory");
      }
      b ^= 1;
      if (b == 0) aph ^= 1u;
    }
  } else if (warp < 6 + kE5PackWarps) {
    // ===================== bit packers: one bit per cell of every stage =======================
    const int pw = warp - 6;   // persons [32 pw, 32 pw + 32) of the tile
    // byte offset of element (row, item = lane) inside an 8-row swizzle atom, for row & 7 = k
    uint32_t off8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      off8[k] = (uint32_t)k * 128u + ((((uint32_t)lane >> 2) ^ (uint32_t)k) << 4) + ((uint32_t)lane & 3u) * 4u;
    int s = 0, b = 0;
    uint32_t ph = 0, aph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      uint32_t* bits = s_bits + b * kE5Rows * 32;
      bar_wait_idle(tile_done(b), aph ^ 1u);
      for (int kb = 0; kb < p.n_kb; ++kb) {
        bar_wait(full(s), ph);
        const uint32_t st = st_base + (uint32_t)s * kE5StageBytes + (uint32_t)pw * 32u * 128u;
        // pull the warp's 32 x 32 cells into registers and hand the stage back at once: the ballots
        // below then run off the ring's critical path
        uint32_t xs[32];
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) xs[rr] = lds32(st + (uint32_t)(rr >> 3) * 1024u + off8[rr & 7]);
        // The ballot of the LAST row loaded is taken first: warps issue in order, so once it has issued every
        // load of this stage has returned its data and the stage can be handed back (an arrive placed right
        // behind the loads would rely on the load and mbarrier pipes staying in order).
        uint32_t w_last = 0;
        if (!(p.debug & 128)) w_last = __ballot_sync(0xffffffffu, __uint_as_float(xs[31]) > 0.5f);
        if (!(p.debug & 16)) {
          __syncwarp();
          if (lane == 0) bar_arrive(empty(s));   // this warp's reads of the stage are complete
        }
        // one ballot per person row; lane 0 stores the word (a per-row "lane == row" select costs more
        // instructions than the store it saves)
        if (!(p.debug & 2)) {
          uint32_t* dst = bits + kb * kE5Rows + pw * 32;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            uint32_t w = w_last;
            if (rr < 31 || (p.debug & 128)) w = __ballot_sync(0xffffffffu, __uint_as_float(xs[rr]) > 0.5f);
            if (lane == 0) dst[rr] = w;   // word of person 32 pw + rr
          }
        }
        if (p.debug & 16) {   // hand the stage back only after the ballots consumed every loaded word
          __syncwarp();
          if (lane == 0) bar_arrive(empty(s));
        }
        if (++s == kE5Stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (!(p.debug & 128)) __threadfence_block();   // the ballot words are performed before the arrive below
      __syncwarp();
      if (lane == 0) bar_arrive(theta_full(b));   // release: the bits of this tile
      if (p.debug & 8) {   // lock-step with the link warps (racecheck runs, see E5Params::debug)
        asm volatile("bar.sync 3, 768;" ::: "memory");
        asm volatile("bar.sync 3, 768;" ::: "memory");
      }
      b ^= 1;
      if (b == 0) aph ^= 1u;
    }
  } else {
    // ===================== link warps: score tile t from the bits while tile t + 1 streams ========
    const int lw = warp - 6 - kE5PackWarps;
    constexpr int NB = 2;   // item blocks per warp: lw and lw + 16
    float a[NB][MODEL == 1 ? 1 : D], bj[NB], gj[NB], omg[NB], wv[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int j = (lw + 16 * q) * kE5KB + lane;
      const bool valid = (lw + 16 * q) < p.n_kb && j < I;
      const int jc = valid ? j : 0;
      if (MODEL == 1) {
        bj[q] = p.item_feat[jc];
        a[q][0] = 0.0f;
        gj[q] = 0.0f;
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) a[q][d] = p.item_feat[(size_t)jc * F + d];
        bj[q] = p.item_feat[(size_t)jc * F + D];
        gj[q] = MODEL == 3 ? 1.0f / (1.0f + expf(-p.item_feat[(size_t)jc * F + D + 1])) : 0.0f;
      }
      omg[q] = 1.0f - gj[q];
      wv[q] = valid ? 1.0f : 0.0f;
    }
    // the two item blocks as f32x2 pairs; the validity weight (and ln 2 for the 3PL) folded into one scale
    f2_t na2[MODEL == 1 ? 1 : D];
#pragma unroll
    for (int d = 0; d < (MODEL == 1 ? 1 : D); ++d) na2[d] = pack2(-a[0][d], -a[1][d]);
    const f2_t b2 = pack2(bj[0], bj[1]), g2 = pack2(gj[0], gj[1]), omg2 = pack2(omg[0], omg[1]);
    const float sc = MODEL == 3 ? kLn2f : 1.0f;
    const f2_t scale2 = pack2(wv[0] * sc, wv[1] * sc);
    const uint32_t lane_bit = 1u << lane;
    const int kb0 = lw & 31, kb1 = (lw + 16) & 31;
    int b = 0;
    uint32_t tph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const uint32_t* bits = s_bits + b * kE5Rows * 32;
      const float* th_tile = s_theta + b * kE5Rows * kE5ThLd;
      const uint8_t* s_flag = s_flags + b * kE5Rows;
      bar_wait_idle(theta_full(b), tph);
      if (p.debug & 8) asm volatile("bar.sync 3, 768;" ::: "memory");
      f2_t ll2 = pack2(0.0f, 0.0f);
      // four persons per trip: one load of their flags and one of each item block's four bit words
#pragma unroll 2
      for (int p4 = 0; p4 < ((p.debug & 1) ? 0 : kE5Rows); p4 += 4) {
        const uint32_t f4 = *reinterpret_cast<const uint32_t*>(s_flag + p4);
        const uint4 w0 = *reinterpret_cast<const uint4*>(bits + kb0 * kE5Rows + p4);
        const uint4 w1 = *reinterpret_cast<const uint4*>(bits + kb1 * kE5Rows + p4);
        const uint32_t wa[4] = {w0.x, w0.y, w0.z, w0.w}, wb[4] = {w1.x, w1.y, w1.z, w1.w};
        f2_t prod2 = pack2(1.0f, 1.0f);   // 3PL: product of the four persons' probabilities per item
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if ((f4 >> (8 * q)) & 0xffu) continue;   // missing cells (handled by the epilogue) or past the end
          const float* tp = th_tile + (p4 + q) * kE5ThLd;
          const float4 t4 = *reinterpret_cast<const float4*>(tp);
          const float4 t5 = *reinterpret_cast<const float4*>(tp + 4);
          const float tv[8] = {t4.x, t4.y, t4.z, t4.w, t5.x, t5.y, t5.z, t5.w};
          f2_t z2 = b2;
          if (MODEL == 1) {
            z2 = add2(z2, pack2(tv[7], tv[7]));
          } else {
#pragma unroll
            for (int d = 0; d < D; ++d) z2 = fma2(pack2(tv[d], tv[d]), na2[d], z2);
          }
          const bool x0 = (wa[q] & lane_bit) != 0, x1 = (wb[q] & lane_bit) != 0;
          if (MODEL == 3) prod2 = mul2(prod2, prob_pair_3pl(z2, x0, x1, g2, omg2));
          else ll2 = fma2(eval_pair<MODEL>(z2, x0, x1, g2, omg2), scale2, ll2);
        }
        if (MODEL == 3) {
          float u0, u1;
          unpack2(prod2, u0, u1);
          ll2 = fma2(pack2(lg2_approx(u0), lg2_approx(u1)), scale2, ll2);
        }
      }
      float l0, l1;
      unpack2(ll2, l0, l1);
      acc_ll += (double)(l0 + l1);
      __syncwarp();
      if (lane == 0) bar_arrive(tile_done(b));
      if (p.debug & 8) asm volatile("bar.sync 3, 768;" ::: "memory");
      b ^= 1;
      if (b == 0) tph ^= 1u;
    }
  }

  // ---- CTA result (fixed order) ---------------------------------------------------------------
  acc_ll = warp_sum(acc_ll);
  acc_term = warp_sum(acc_term);
  if (lane == 0) {
    s_red[warp * 2] = acc_ll;
    s_red[warp * 2 + 1] = acc_term;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (t == 0) {
    double a0 = 0.0, a1 = 0.0;
    for (int w = 0; w < kE5Threads / 32; ++w) {
      a0 += s_red[w * 2];
      a1 += s_red[w * 2 + 1];
    }
    p.part[(size_t)blockIdx.x * 2] = a0;
    p.part[(size_t)blockIdx.x * 2 + 1] = a1;
  }
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kE5TmemCols) : "memory");
  }
}

__global__ void __launch_bounds__(64) e5_sum_kernel(const double* __restrict__ part, int nparts,
                                                    double* __restrict__ out) {
  // two outputs, each summed by one warp in a fixed order
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (int q = lane; q < nparts; q += 32) s += part[(size_t)q * 2 + c];
  s = warp_sum(s);
  if (lane == 0) out[c] = s;
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

template <int MODEL, int D>
cudaError_t launch_e5(const CUtensorMap& tmap, const E5Params& p, int grid, size_t smem, cudaStream_t st) {
  auto k = tc5_eval_kernel<MODEL, D>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<grid, kE5Threads, smem, st>>>(tmap, p);
  return cudaGetLastError();
}

template <int MODEL>
cudaError_t launch_e5_d(int D, const CUtensorMap& tmap, const E5Params& p, int grid, size_t smem, cudaStream_t st) {
  switch (D) {
    case 1: return launch_e5<MODEL, 1>(tmap, p, grid, smem, st);
    case 2: return launch_e5<MODEL, 2>(tmap, p, grid, smem, st);
    case 3: return launch_e5<MODEL, 3>(tmap, p, grid, smem, st);
    case 4: return launch_e5<MODEL, 4>(tmap, p, grid, smem, st);
    default: return launch_e5<MODEL, 5>(tmap, p, grid, smem, st);
  }
}

}  // namespace

size_t tc5_eval_workspace_bytes() { return (size_t)sm_count() * 2 * sizeof(double) + 256; }

// Forward-only fused ELBO of the conditional posterior.  cudaErrorNotSupported when the shape /
// pointers are not covered (the caller composes encode -> person_forward -> link instead).
cudaError_t tc5_eval(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* table,
                     const float* item_feat, const float* eps, uint64_t seed, const uint64_t* seed_dev,
                     double* out_scalars, float* amu, float* alv, float* ability, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  const char* off = getenv("VIBO_DISABLE_TC5");
  if (off != nullptr && off[0] == '1') return cudaErrorNotSupported;
  const char* off2 = getenv("VIBO_DISABLE_TC5_EVAL");
  if (off2 != nullptr && off2[0] == '1') return cudaErrorNotSupported;
  const int I = d.num_item, D = d.ability_dim;
  if (!d.conditional || D > 5 || (I & 3) != 0 || I < 32 || I > 1024 || d.num_person < 1) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(resp) & 15) || (reinterpret_cast<uintptr_t>(mask) & 15)) return cudaErrorNotSupported;
  if (ws_bytes < tc5_eval_workspace_bytes()) return cudaErrorNotSupported;
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return cudaErrorNotSupported;
  E5Params p;
  p.P = d.num_person; p.I = I; p.n_kb = (I + kE5KB - 1) / kE5KB; p.missing_policy = d.missing_policy;
  p.form = d.elbo_form; p.person_offset = d.person_offset;
  { const char* dbg = getenv("VIBO_E5_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; } p.resp = resp; p.mask = mask; p.table = table;
  p.item_feat = item_feat; p.eps = eps; p.seed = seed; p.seed_dev = seed_dev;
  const bool person_out = amu != nullptr && alv != nullptr && ability != nullptr;
  p.out_mu = person_out ? amu : nullptr; p.out_lv = person_out ? alv : nullptr;
  p.out_theta = person_out ? ability : nullptr;
  p.part = static_cast<double*>(ws);
  const E5Smem L = e5_layout(p.n_kb);
  if (L.total > 227 * 1024) return cudaErrorNotSupported;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)I, (cuuint64_t)d.num_person};
  const cuuint64_t strides[1] = {(cuuint64_t)I * sizeof(float)};
  const cuuint32_t box[2] = {kE5KB, kE5Rows};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(resp), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorNotSupported;
  const int64_t n_tiles = (d.num_person + kE5Rows - 1) / kE5Rows;
  int grid = sm_count();
  if ((int64_t)grid > n_tiles) grid = (int)n_tiles;
  cudaError_t e;
  if (d.irt_model == 1) e = launch_e5_d<1>(D, tmap, p, grid, L.total, st);
  else if (d.irt_model == 2) e = launch_e5_d<2>(D, tmap, p, grid, L.total, st);
  else e = launch_e5_d<3>(D, tmap, p, grid, L.total, st);
  if (e != cudaSuccess) return e;
  e5_sum_kernel<<<1, 64, 0, st>>>(p.part, grid, out_scalars);
  note_launch(2);
  return cudaGetLastError();
}

}  // namespace vibo
