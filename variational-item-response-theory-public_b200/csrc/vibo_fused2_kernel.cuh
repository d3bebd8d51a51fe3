// Two-phase single-pass fused ELBO kernel for the unconditional 1PL / 2PL
// encoder (sm_100a): the production kernel for the headline configuration.
//
// Same data movement as vibo_fused_kernel.cuh (per-team rings of 1-D TMA bulk
// copies on mbarriers; each row read from HBM once).  What differs is where
// the cross-person item-gradient sums live.  There, each lane keeps the
// accumulators of the 32 items it owns in registers (64+ registers), which
// caps the CTA at 16 warps and leaves the kernel latency-bound.  Here a team
// of 4 warps processes the rows of a stage in two phases:
//
//   phase A (one warp per row): counts -> posterior -> draw -> link ->
//       log-likelihood; d ll / d z of every cell is written back to shared
//       memory IN PLACE of the response value it was computed from;
//   team barrier;
//   phase B (one thread per item group): each of the team's 128 threads owns
//       one or two 4-item groups for the whole kernel and adds the stage's
//       rows' d ll / d z (and theta-weighted d ll / d z) into 8..24 registers.
//
// Registers drop below 96, so 20 warps (5 teams) are resident per SM.
//
// Clamp handling: the eps32 clamp of torch.distributions only acts where
// |z| > 15.94.  A row whose bound sum_d |theta_d| max_j|a_jd| + max_j|b_j| is
// below that cannot reach the clamp, and takes a path without the clamp and
// range-test instructions; other rows take the exact path.
#pragma once

#include "vibo_fused_kernel.cuh"

namespace vibo {

constexpr int kF2Teams = 5;
constexpr int kF2TeamWarps = 4;
constexpr int kF2Warps = kF2Teams * kF2TeamWarps;
constexpr int kF2Threads = kF2Warps * 32;
constexpr int kF2TeamThreads = kF2TeamWarps * 32;
constexpr int kF2GroupsPerThread = 2;     // phase-B groups per thread: I <= 4 * 128 * 2 = 1024
constexpr int kF2MaxRows = 64;            // rows per stage upper bound (team scratch sizing)
// scratch between the barrier block and the item parameters:
//   [0, 64)        max_j |a_jd| (d < 8), max_j |b_j|, then sum_j a'_jd (d < 2), sum_j b'_j
//   [64, 576)      per-warp double partials of those sums (setup only)
//   [576, ...)     per-team theta of the rows of each ring slot, every value stored
//                  twice (an f32x2 pair for phase B): [team][slot < NS][row < R][d < D][2]
//                  (phase A of the next slot may start while a slower warp is still
//                  in phase B, hence one copy per slot)
constexpr int kF2SumOff = 64, kF2ThetaOff = 576;
__host__ __device__ inline int f2_scratch_bytes(int R, int NS, int D) {
  const int theta = kF2Teams * NS * R * D * 8;
  const int owner = kF2Teams * 768;   // fused3_kernel (same plan): per-team counts, g_theta partials, table sums
  return kF2ThetaOff + (theta > owner ? theta : owner);
}

// ---- packed f32x2 arithmetic (FFMA2 / FADD2 / FMUL2): two floats per issue slot ----
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) {
  f2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) {
  f2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2_t sub2(f2_t a, f2_t b) {
  f2_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) {
  f2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void lds128_f2(uint32_t addr, f2_t& lo, f2_t& hi) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(addr));
}
__device__ __forceinline__ f2_t lds64_f2(uint32_t addr) {
  f2_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128_f2(uint32_t addr, f2_t lo, f2_t hi) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(lo), "l"(hi) : "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void team_barrier(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(kF2TeamThreads) : "memory");
}

// Item parameters live in shared memory PRE-SCALED for the base-2 exponential:
//   a'_jd = log2(e) a_jd,  b'_j = -log2(e) b_j,  so  z'_ij = b'_j + theta_i . a'_j = -log2(e) z_ij
// and E = exp(-z) = ex2(z').  Sums of dz * a' are un-scaled by ln 2 per person.
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2f = 0.6931471805599453f;

// One 4-cell group of phase A, general path (EXACT keeps the clamp and its
// range test; !FULL consults the mask).  Accumulates in natural-log units.
template <int MODEL, int D, bool GRAD, bool FULL, bool EXACT>
__device__ __forceinline__ void f2_group(uint32_t xaddr, uint32_t m4, const float4& b4,
                                         const float4 (&a4)[MODEL == 1 ? 1 : D], const float (&th)[D],
                                         float t1, float (&gth)[D], float& s1, float& s3) {
  constexpr int DA = MODEL == 1 ? 0 : D;
  const float4 x4 = lds128(xaddr);
  const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
  const float bs[4] = {b4.x, b4.y, b4.z, b4.w};
  float dzs[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float av[DA > 0 ? DA : 1];
#pragma unroll
    for (int d = 0; d < DA; ++d) av[d] = c == 0 ? a4[d].x : (c == 1 ? a4[d].y : (c == 2 ? a4[d].z : a4[d].w));
    float zs = bs[c];   // z' = -log2(e) z
    if (MODEL == 1) {
      zs += t1;
    } else {
#pragma unroll
      for (int d = 0; d < DA; ++d) zs = fmaf(th[d], av[d], zs);
    }
    float z = zs * -kLn2f;
    // ll = x zc - softplus(zc) = (x - 1) zc - log(1 + E),  E = exp(-zc);  d ll / d z = x - 1/(1 + E)
    float x = xs[c];
    if (!FULL) {
      const bool o = ((m4 >> (8 * c)) & 0xffu) != 0;
      x = o ? x : 0.5f;   // missing cell -> neutral cell (x = 1/2, z = 0): gradient 0, ll = -log 2
      z = o ? z : 0.0f;
    }
    const float zc = EXACT ? fminf(fmaxf(z, -kLogitClamp), kLogitClamp) : z;
    const float w = 1.0f + ex2_approx(zc * -kLog2e);
    s1 = fmaf(x - 1.0f, zc, s1);
    s3 += lg2_approx(w);
    float dz = 0.0f;
    if (GRAD) {
      dz = x - rcp_approx(w);                  // x - sigmoid(zc)
      if (EXACT) dz = (z == zc) ? dz : 0.0f;  // zero gradient outside the eps32 clamp
      if (MODEL == 1) {
        gth[0] -= dz;
      } else {
#pragma unroll
        for (int d = 0; d < DA; ++d) gth[d] = fmaf(dz, av[d], gth[d]);   // log2(e) * sum dz a
      }
    }
    dzs[c] = dz;
  }
  if (GRAD) sts128(xaddr, make_float4(dzs[0], dzs[1], dzs[2], dzs[3]));
}

template <int MODEL, int D, int LPP, int NG, bool GRAD, bool FULL, bool EXACT>
__device__ __forceinline__ void f2_pass2(uint32_t xp, uint32_t mp, uint32_t pp, int I4, int kfull, bool has_tail,
                                         const float (&th)[D], float t1, float (&gth)[D], float& s1, float& s3) {
  constexpr int DA = MODEL == 1 ? 0 : D;
  float nmiss_lane = 0.0f;
  auto one = [&](int gi) {
    uint32_t m4 = 0x01010101u;
    if (!FULL) {
      m4 = lds32(mp + gi * 4);
      nmiss_lane += (float)(4 - __popc(m4 & 0x01010101u));
    }
    float4 a4[DA > 0 ? DA : 1];
#pragma unroll
    for (int d = 0; d < DA; ++d) a4[d] = lds128(pp + (d * I4 + gi) * 16);
    const float4 b4 = lds128(pp + (DA * I4 + gi) * 16);
    f2_group<MODEL, D, GRAD, FULL, EXACT>(xp + gi * 16, m4, b4, a4, th, t1, gth, s1, s3);
  };
  // no per-group register state is indexed by k, so this is a real loop
  // (warp-uniform trip count), lightly unrolled for instruction-level parallelism
#pragma unroll 2
  for (int k = 0; k < kfull; ++k) one(LPP * k);
  if (has_tail) one(LPP * kfull);
  if (!FULL) s3 -= nmiss_lane;  // each neutral cell added log2(2) = 1
}

// Fast path: every cell of the row observed and no logit can reach the clamp.
// Packed f32x2 arithmetic, and only 6 (GRAD) / 5 MUFU per 4 cells instead of 12 / 8:
//   sum_j log2(1 + E_j) over 4 cells = log2 of the product (each factor < 2^23.1, so
//     the product of four stays far below FLT_MAX), and
//   1 / (1 + E_j) for the 4 cells from ONE reciprocal of that product.
// Accumulates, in log2 units, s1x = sum_j x_j z'_j (pair) and s3 = sum_j log2(1 + E_j);
// the row's  sum_j z'_j  is added in closed form by the caller:
//   ll_row = ln 2 * (sum_j z'_j - s1x - s3).
template <int MODEL, int D, bool GRAD>
__device__ __forceinline__ void f2_group_fast(uint32_t xaddr, uint32_t paddr, int I4, const f2_t (&th2)[D],
                                              f2_t t2, f2_t (&g2)[D], f2_t& s1x, float& s3) {
  constexpr int DA = MODEL == 1 ? 0 : D;
  f2_t x01, x23, z01, z23;
  lds128_f2(xaddr, x01, x23);
  lds128_f2(paddr + DA * I4 * 16, z01, z23);   // b'
  f2_t a01[DA > 0 ? DA : 1], a23[DA > 0 ? DA : 1];
  if (MODEL == 1) {
    z01 = add2(z01, t2);
    z23 = add2(z23, t2);
  } else {
#pragma unroll
    for (int d = 0; d < DA; ++d) {
      lds128_f2(paddr + d * I4 * 16, a01[d], a23[d]);
      z01 = fma2(th2[d], a01[d], z01);
      z23 = fma2(th2[d], a23[d], z23);
    }
  }
  float z0, z1, z2, z3;
  unpack2(z01, z0, z1);
  unpack2(z23, z2, z3);
  const f2_t one2 = pack2(1.0f, 1.0f);
  const f2_t w01 = add2(pack2(ex2_approx(z0), ex2_approx(z1)), one2);
  const f2_t w23 = add2(pack2(ex2_approx(z2), ex2_approx(z3)), one2);
  const f2_t c = mul2(w01, w23);   // (w0 w2, w1 w3)
  float cx, cy;
  unpack2(c, cx, cy);
  const float prod = cx * cy;
  s3 += lg2_approx(prod);
  s1x = fma2(x01, z01, s1x);
  s1x = fma2(x23, z23, s1x);
  if (GRAD) {
    const float r = rcp_approx(prod);
    const f2_t cs = pack2(cy * r, cx * r);           // (1/(w0 w2), 1/(w1 w3))
    const f2_t dz01 = sub2(x01, mul2(w23, cs));      // x - 1/w = x - sigmoid(z)
    const f2_t dz23 = sub2(x23, mul2(w01, cs));
    if (MODEL == 1) {
      g2[0] = add2(g2[0], add2(dz01, dz23));
    } else {
#pragma unroll
      for (int d = 0; d < DA; ++d) {
        g2[d] = fma2(dz01, a01[d], g2[d]);
        g2[d] = fma2(dz23, a23[d], g2[d]);
      }
    }
    sts128_f2(xaddr, dz01, dz23);
  }
}

template <int MODEL, int D, int LPP, int NG, bool GRAD>
__device__ __forceinline__ void f2_pass2_fast(uint32_t xp, uint32_t pp, int I4, int kfull, bool has_tail,
                                              const float (&th)[D], float t1, float (&gth)[D], float& s1x,
                                              float& s3) {
  constexpr int DA = MODEL == 1 ? 0 : D;
  f2_t th2[D], g2[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    th2[d] = pack2(th[d], th[d]);
    g2[d] = pack2(0.0f, 0.0f);
  }
  const f2_t t2 = pack2(t1, t1);
  f2_t s1p = pack2(0.0f, 0.0f);
#pragma unroll 2
  for (int k = 0; k < kfull; ++k)
    f2_group_fast<MODEL, D, GRAD>(xp + LPP * k * 16, pp + LPP * k * 16, I4, th2, t2, g2, s1p, s3);
  if (has_tail)
    f2_group_fast<MODEL, D, GRAD>(xp + LPP * kfull * 16, pp + LPP * kfull * 16, I4, th2, t2, g2, s1p, s3);
  float lo, hi;
  unpack2(s1p, lo, hi);
  s1x = lo + hi;
  if (GRAD) {
#pragma unroll
    for (int d = 0; d < (MODEL == 1 ? 1 : DA); ++d) {
      unpack2(g2[d], lo, hi);
      gth[d] = MODEL == 1 ? -(lo + hi) : lo + hi;
    }
  }
}

template <int MODEL, int D, int LPP, int NG, bool GRAD>
__global__ void __launch_bounds__(kF2Threads, 1) fused2_kernel(const __grid_constant__ FusedParams p) {
  static_assert(MODEL == 1 || MODEL == 2, "two-phase kernel covers 1PL / 2PL");
  constexpr int F = item_width(MODEL, D);
  constexpr int DA = MODEL == 1 ? 0 : D;
  constexpr int TW = kF2TeamWarps, NQ = kF2Teams;
  constexpr int PPW = 32 / LPP;
  constexpr int NGB = kF2GroupsPerThread;
  constexpr float kLn2 = 0.6931471805599453f;

  extern __shared__ __align__(128) unsigned char smem[];
  const int I = p.I, R = p.R, NS = p.nstage;
  const FusedSmem L = fused_smem_layout(I, D, MODEL, R, NS, NQ, f2_scratch_bytes(R, NS, D));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);   // [NQ * NS] (<= 16)
  int* done_cnt = reinterpret_cast<int*>(smem + 128);
  uint64_t* empty_bar = reinterpret_cast<uint64_t*>(smem + 256);   // [NQ * NS]: every warp of a team arrives when it leaves the stage
  float* s_max = reinterpret_cast<float*>(smem + kFusedHdr);      // amax[0..7], bmax, sum a'[0..1], sum b'
  double* s_wsum = reinterpret_cast<double*>(smem + kFusedHdr + kF2SumOff);   // [warp][3] setup partials
  float* s_param = reinterpret_cast<float*>(smem + L.params_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / TW, wt = warp % TW, tt = threadIdx.x - team * kF2TeamThreads;
  const int n_groups = I >> 2;
  const int64_t n_chunks = (p.P + R - 1) / R;
  const int64_t chunk0 = (int64_t)blockIdx.x * NQ + team, chunk_step = (int64_t)gridDim.x * NQ;
  uint64_t* t_full = full_bar + team * NS;
  int* t_done = done_cnt + team * NS;
  uint64_t* t_empty = empty_bar + team * NS;
  unsigned char* t_stage = smem + L.stage_off + (size_t)team * NS * L.stage_bytes;
  static_assert(D <= 2, "setup sums / phase B are sized for D <= 2");
  float* t_theta = reinterpret_cast<float*>(smem + kFusedHdr + kF2ThetaOff) + (size_t)team * NS * R * D * 2;

  // ---- one-time setup ----------------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < NQ * NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], TW);
      done_cnt[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 16) s_max[threadIdx.x] = 0.0f;
  fused_draw_noise<D>(p, NQ);
  __syncthreads();
  {
    float amax[DA > 0 ? DA : 1], bmax = 0.0f;
    double asum[DA > 0 ? DA : 1], bsum = 0.0;
#pragma unroll
    for (int d = 0; d < (DA > 0 ? DA : 1); ++d) {
      amax[d] = 0.0f;
      asum[d] = 0.0;
    }
    for (int j = threadIdx.x; j < I; j += blockDim.x) {
      if (MODEL == 1) {
        const float b = p.item_feat[j], bs = -kLog2e * b;
        s_param[j] = bs;
        bsum += (double)bs;
        bmax = fmaxf(bmax, fabsf(b));
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float a = p.item_feat[(size_t)j * F + d], as = kLog2e * a;
          s_param[(size_t)d * I + j] = as;
          asum[d] += (double)as;
          amax[d] = fmaxf(amax[d], fabsf(a));
        }
        const float b = p.item_feat[(size_t)j * F + D], bs = -kLog2e * b;
        s_param[(size_t)D * I + j] = bs;
        bsum += (double)bs;
        bmax = fmaxf(bmax, fabsf(b));
      }
    }
    // sum_j a'_jd, sum_j b'_j: warp partials, combined in a fixed order below
#pragma unroll
    for (int d = 0; d < DA; ++d) {
      const double v = warp_sum(asum[d]);
      if (lane == 0) s_wsum[warp * 3 + d] = v;
    }
    {
      const double v = warp_sum(bsum);
      if (lane == 0) s_wsum[warp * 3 + 2] = v;
    }
    // non-negative floats order like their bit patterns: atomicMax on ints
#pragma unroll
    for (int d = 0; d < DA; ++d) atomicMax(reinterpret_cast<int*>(&s_max[d]), __float_as_int(amax[d]));
    atomicMax(reinterpret_cast<int*>(&s_max[8]), __float_as_int(bmax));
  }
  // prologue: the first NS chunks of each team
  if (wt == 0) {
    for (int s = 0; s < NS; ++s) {
      const int64_t c = chunk0 + (int64_t)s * chunk_step;
      if (c < n_chunks) fused_issue_chunk<D>(p, L, c, t_stage + (size_t)s * L.stage_bytes, &t_full[s], lane);
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < kF2Warps; ++w) v += s_wsum[w * 3 + threadIdx.x];
    s_max[9 + threadIdx.x] = ((int)threadIdx.x < DA || threadIdx.x == 2) ? (float)v : 0.0f;
  }
  __syncthreads();

  // ---- per-thread state ------------------------------------------------------
  float ll_acc = 0.0f, ll2_acc = 0.0f, term_acc = 0.0f;   // per lane; summed in double at the end
  float tA[2][D], tB[2][D];                // expert-table gradient sums (sub-group leaders)
  // phase-B item-gradient sums of this thread's groups, as f32x2 pairs over cells (0,1) and (2,3):
  // accA[k][d][h] = sum_i dz * theta_d,  accB[k][h] = sum_i dz
  f2_t accA[GRAD ? NGB : 1][DA > 0 ? DA : 1][2], accB[GRAD ? NGB : 1][2];
#pragma unroll
  for (int d = 0; d < D; ++d) tA[0][d] = tA[1][d] = tB[0][d] = tB[1][d] = 0.0f;
#pragma unroll
  for (int k = 0; k < (GRAD ? NGB : 1); ++k) {
    accB[k][0] = accB[k][1] = pack2(0.0f, 0.0f);
#pragma unroll
    for (int d = 0; d < (DA > 0 ? DA : 1); ++d) accA[k][d][0] = accA[k][d][1] = pack2(0.0f, 0.0f);
  }

  const int sub = lane / LPP, q = lane % LPP;
  const uint32_t submask = LPP == 32 ? 0xffffffffu : (((1u << LPP) - 1u) << (sub * LPP));
  float tau[2][D], mt[2][D], amaxv[DA > 0 ? DA : 1];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float mu = p.table[r * 2 * D + d], lam = p.table[r * 2 * D + D + d];
      tau[r][d] = 1.0f / (expf(lam) + kPoeEps);
      mt[r][d] = mu * tau[r][d];
    }
#pragma unroll
  for (int d = 0; d < DA; ++d) amaxv[d] = s_max[d];
  const float bmaxv = s_max[8];
  float asumv[DA > 0 ? DA : 1];
#pragma unroll
  for (int d = 0; d < DA; ++d) asumv[d] = s_max[9 + d];
  const float bsumv = s_max[11];
  const float prior_tau = p.missing_policy == VIBO_MISSING_PRIOR ? 1.0f / (1.0f + kPoeEps) : 0.0f;
  const int kfull = min(n_groups / LPP, NG);
  const bool has_tail = kfull < NG && q < n_groups - kfull * LPP;
  uint32_t pp = smem_u32(s_param) + q * 16;
  asm volatile("mov.u32 %0, %0;" : "+r"(pp));

  const uint32_t stage0 = smem_u32(t_stage), bar0 = smem_u32(t_full), theta0 = smem_u32(t_theta);
  const int first_row = wt * PPW + sub;
  const uint32_t off_x = (uint32_t)first_row * I * 4 + q * 16;
  const uint32_t off_m = (uint32_t)L.mask_off + (uint32_t)first_row * I + q * 4;
  const uint32_t off_e = (uint32_t)L.eps_off + (uint32_t)first_row * D * 4;
  const uint32_t step_x = (uint32_t)TW * PPW * I * 4, step_m = (uint32_t)TW * PPW * I,
                 step_e = (uint32_t)TW * PPW * D * 4;
  int s = 0;
  uint32_t phase = 0;
  // 32-bit loop state: this team runs n_it stages; only the globally last chunk can be ragged
  const int n_it = chunk0 < n_chunks ? (int)((n_chunks - chunk0 + chunk_step - 1) / chunk_step) : 0;
  const int last_rows = (int)(p.P - (n_chunks - 1) * R);
  const bool owns_last = n_it > 0 && chunk0 + (int64_t)(n_it - 1) * chunk_step == n_chunks - 1;
  const uint32_t stage_bytes = (uint32_t)L.stage_bytes;

  for (int it = 0; it < n_it; ++it) {
    mbar_wait_addr(bar0 + (uint32_t)s * 8, phase);
    const uint32_t sb = stage0 + (uint32_t)s * stage_bytes;
    const uint32_t thb = theta0 + (uint32_t)s * (uint32_t)(R * D * 8);  // this slot's theta rows (pairs)
    const int rows = (owns_last && it == n_it - 1) ? last_rows : R;

    // ======================= phase A: one sub-group per row =================
    uint32_t xrow = sb + off_x, mrow = sb + off_m, erow = sb + off_e;
    for (int rbase = wt * PPW; rbase < rows; rbase += TW * PPW, xrow += step_x, mrow += step_m, erow += step_e) {
      const int r = rbase + sub;
      const bool valid = r < rows;
      const uint32_t xp = valid ? xrow : xrow - (uint32_t)sub * I * 4;
      const uint32_t mp = valid ? mrow : mrow - (uint32_t)sub * I;
      const uint32_t ep = valid ? erow : erow - (uint32_t)sub * D * 4;

      // ---- pass 1: counts --------------------------------------------------
      f2_t n1p2 = pack2(0.0f, 0.0f);
      uint32_t mand = 0x01010101u;
#pragma unroll 4
      for (int k = 0; k < kfull; ++k) {
        f2_t x01, x23;
        lds128_f2(xp + LPP * k * 16, x01, x23);
        n1p2 = add2(n1p2, add2(x01, x23));
        mand &= lds32(mp + LPP * k * 4);
      }
      if (has_tail) {
        f2_t x01, x23;
        lds128_f2(xp + LPP * kfull * 16, x01, x23);
        n1p2 = add2(n1p2, add2(x01, x23));
        mand &= lds32(mp + LPP * kfull * 4);
      }
      const bool full_obs = __all_sync(0xffffffffu, mand == 0x01010101u);
      float n1lo, n1hi;
      unpack2(n1p2, n1lo, n1hi);
      int n1 = (int)(n1lo + n1hi), nobs = 0;   // 0/1 responses: the float partial sums are exact small integers
      if (!full_obs) {
        n1 = 0;
        auto count = [&](int gi) {
          const float4 x = lds128(xp + gi * 16);
          const uint32_t m = lds32(mp + gi * 4);
          const bool o0 = (m & 0xffu) != 0, o1 = (m & 0xff00u) != 0, o2 = (m & 0xff0000u) != 0,
                     o3 = (m & 0xff000000u) != 0;
          nobs += (int)o0 + (int)o1 + (int)o2 + (int)o3;
          n1 += (int)(o0 && x.x > 0.5f) + (int)(o1 && x.y > 0.5f) + (int)(o2 && x.z > 0.5f) +
                (int)(o3 && x.w > 0.5f);
        };
#pragma unroll 2
        for (int k = 0; k < kfull; ++k) count(LPP * k);
        if (has_tail) count(LPP * kfull);
      }
      // integer sub-group sums in one REDUX each
      const float n1f = (float)__reduce_add_sync(submask, n1);
      const float nobsf = full_obs ? (float)I : (float)__reduce_add_sync(submask, nobs);
      const float n0f = nobsf - n1f, nmiss = (float)I - nobsf;

      // ---- per-person posterior and draw -----------------------------------
      float amu[D], invS[D], sd[D], th[D], epsv[D], alv[D];
      float tsum = 0.0f, term = 0.0f, bound = bmaxv;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float S = fmaf(n0f, tau[0][d], fmaf(n1f, tau[1][d], nmiss * prior_tau));
        const float N = fmaf(n0f, mt[0][d], n1f * mt[1][d]);
        invS[d] = rcp_approx(S);                 // = exp(logvar)
        amu[d] = N * invS[d];
        alv[d] = -kLn2 * lg2_approx(S);          // log(1 / S)
        sd[d] = rsqrt_approx(S);                 // exp(logvar / 2)
        epsv[d] = __uint_as_float(lds32(ep + d * 4));
        th[d] = fmaf(epsv[d], sd[d], amu[d]);
        tsum += th[d];
        bound = MODEL == 1 ? bound + fabsf(th[d]) : fmaf(fabsf(th[d]), amaxv[MODEL == 1 ? 0 : d], bound);
        if (p.form == VIBO_ELBO_KL) {
          term += -0.5f * (1.0f + alv[d] - amu[d] * amu[d] - invS[d]);
        } else {
          term += -0.5f * th[d] * th[d] + 0.5f * epsv[d] * epsv[d] + 0.5f * alv[d];
        }
      }
      if (valid && q == 0) {
        term_acc += term;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          if (GRAD && DA > 0)
            asm volatile("st.shared.v2.f32 [%0], {%1, %1};" ::"r"(thb + (uint32_t)(r * D + d) * 8), "f"(th[d]) : "memory");
        }
        if (p.out_mu != nullptr) {
          const int64_t row = (chunk0 + (int64_t)it * chunk_step) * R + r;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            p.out_mu[row * D + d] = amu[d];
            p.out_lv[row * D + d] = alv[d];
            p.out_theta[row * D + d] = th[d];
          }
        }
      }
      // a row can reach the eps32 clamp only if its logit bound exceeds it
      // (NaN bounds, e.g. all-missing rows under --drop-missing, go exact)
      const bool exact = !__all_sync(0xffffffffu, bound <= kLogitClamp);

      // ---- pass 2 -----------------------------------------------------------
      float gth[D];
#pragma unroll
      for (int d = 0; d < D; ++d) gth[d] = 0.0f;
      const float t1 = -kLog2e * tsum;   // 1PL: z' = b' + t1
      if (valid) {
        float s1 = 0.0f, s3 = 0.0f;
        if (full_obs && !exact) {
          f2_pass2_fast<MODEL, D, LPP, NG, GRAD>(xp, pp, n_groups, kfull, has_tail, th, t1, gth, s1, s3);
          // log2 units: sum_j z'_j (closed form, once per row) - sum_j x_j z'_j - sum_j log2(1 + E_j)
          float zsum = 0.0f;
          if (q == 0) {
            zsum = bsumv;
            if (MODEL == 1) {
              zsum = fmaf((float)I, t1, zsum);
            } else {
#pragma unroll
              for (int d = 0; d < DA; ++d) zsum = fmaf(th[d], asumv[d], zsum);
            }
          }
          ll2_acc += zsum - s1 - s3;
        } else {
          if (full_obs)
            f2_pass2<MODEL, D, LPP, NG, GRAD, true, true>(xp, mp, pp, n_groups, kfull, has_tail, th, t1, gth, s1, s3);
          else
            f2_pass2<MODEL, D, LPP, NG, GRAD, false, true>(xp, mp, pp, n_groups, kfull, has_tail, th, t1, gth, s1, s3);
          ll_acc += s1 - kLn2 * s3;
        }
      }

      // ---- per-person backward ---------------------------------------------
      if (GRAD) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float gv = MODEL == 1 ? gth[0] : gth[d];
#pragma unroll
          for (int o = LPP / 2; o > 0; o >>= 1) gv += __shfl_xor_sync(0xffffffffu, gv, o);
          if (MODEL != 1) gv *= kLn2;   // the item discriminations in shared memory carry log2(e)
          float g_mu, g_lv;
          if (p.form == VIBO_ELBO_KL) {
            g_mu = fmaf(p.beta, amu[d], gv);
            g_lv = 0.5f * gv * epsv[d] * sd[d] + 0.5f * p.beta * (invS[d] - 1.0f);
          } else {
            gv += th[d];
            g_mu = gv;
            g_lv = 0.5f * gv * epsv[d] * sd[d] - 0.5f;
          }
          const float GN = g_mu * invS[d];
          const float GS = -(g_mu * amu[d] + g_lv) * invS[d];
          if (valid && q == 0) {
            tA[0][d] = fmaf(n0f, GN, tA[0][d]);
            tA[1][d] = fmaf(n1f, GN, tA[1][d]);
            tB[0][d] = fmaf(n0f, GS, tB[0][d]);
            tB[1][d] = fmaf(n1f, GS, tB[1][d]);
          }
        }
      }
    }

    // ======================= phase B: one thread per item group =============
    if (GRAD) {
      team_barrier(team);  // every row's d ll / d z and theta are in shared memory
#pragma unroll
      for (int k = 0; k < NGB; ++k) {
        const int g = tt + kF2TeamThreads * k;
        if (g < n_groups) {
          uint32_t a = sb + (uint32_t)g * 16;
          auto add_row = [&](int r) {
            f2_t dz01, dz23;
            lds128_f2(a, dz01, dz23);
            a += (uint32_t)I * 4;
#pragma unroll
            for (int d = 0; d < DA; ++d) {
              const f2_t tt = lds64_f2(thb + (uint32_t)(r * D + d) * 8);   // (theta_d, theta_d)
              accA[k][d][0] = fma2(dz01, tt, accA[k][d][0]);
              accA[k][d][1] = fma2(dz23, tt, accA[k][d][1]);
            }
            accB[k][0] = add2(accB[k][0], dz01);
            accB[k][1] = add2(accB[k][1], dz23);
          };
          if (rows == R) {  // R is a multiple of 4
            for (int r = 0; r < R; r += 4) {
              add_row(r);
              add_row(r + 1);
              add_row(r + 2);
              add_row(r + 3);
            }
          } else {
#pragma unroll 1
            for (int r = 0; r < rows; ++r) add_row(r);
          }
        }
      }
    }

    // the last warp of the team to leave the stage refills it
    // (VIBO_FUSED_DEBUG=3: a team barrier first -- the hand-off below is an mbarrier arrive by the
    // reading warps + a wait by the refilling warp, which racecheck does not credit as ordering the
    // bulk copy after the reads; with the barrier in place the tool reports no hazard anywhere else)
    if (p.debug == 3) team_barrier(team);
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      // release: this warp's reads of the stage are complete (mbarrier arrive); the counter only
      // ELECTS the warp that arrived last, which then acquires the completed 'empty' phase before
      // it lets the bulk copy overwrite the stage
      mbar_arrive(&t_empty[s]);
      __threadfence_block();
      last = atomicAdd(&t_done[s], 1) == TW - 1;
      if (last) atomicExch(&t_done[s], 0);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last && it + NS < n_it) {
      const int64_t cn = chunk0 + (int64_t)(it + NS) * chunk_step;
      mbar_wait(&t_empty[s], phase);
      fused_issue_chunk<D>(p, L, cn, t_stage + (size_t)s * L.stage_bytes, &t_full[s], lane);
    }
    if (++s == NS) {
      s = 0;
      phase ^= 1u;
    }
  }

  // ---- CTA-level combine (deterministic order) -----------------------------
  __syncthreads();  // every stage consumed; stage memory is free for reuse
  double* s_d = reinterpret_cast<double*>(smem + L.stage_off);          // [warps][2]
  float* s_t = reinterpret_cast<float*>(smem + L.stage_off + 1024);     // [warps][4D]
  float* s_item = reinterpret_cast<float*>(smem + L.stage_off + 4096);  // [I*F]
  {
    const double a = warp_sum((double)ll_acc + 0.6931471805599453 * (double)ll2_acc),
                 b = warp_sum((double)term_acc);
    if (lane == 0) {
      s_d[warp * 2] = a;
      s_d[warp * 2 + 1] = b;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float v0 = warp_sum(tA[0][d]), v1 = warp_sum(tA[1][d]), v2 = warp_sum(tB[0][d]),
                  v3 = warp_sum(tB[1][d]);
      if (lane == 0) {
        s_t[warp * 4 * D + d] = v0;
        s_t[warp * 4 * D + D + d] = v1;
        s_t[warp * 4 * D + 2 * D + d] = v2;
        s_t[warp * 4 * D + 3 * D + d] = v3;
      }
    }
  }
  if (GRAD)
    for (int k = threadIdx.x; k < I * F; k += blockDim.x) s_item[k] = 0.0f;
  __syncthreads();
  if (GRAD) {
    for (int t = 0; t < NQ; ++t) {
      if (team == t) {
#pragma unroll
        for (int k = 0; k < NGB; ++k) {
          const int g = tt + kF2TeamThreads * k;
          if (g < n_groups) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float lo, hi;
#pragma unroll
              for (int d = 0; d < DA; ++d) {
                unpack2(accA[k][d][h], lo, hi);
                s_item[(size_t)(4 * g + 2 * h) * F + d] += lo;
                s_item[(size_t)(4 * g + 2 * h + 1) * F + d] += hi;
              }
              unpack2(accB[k][h], lo, hi);
              s_item[(size_t)(4 * g + 2 * h) * F + DA] += lo;
              s_item[(size_t)(4 * g + 2 * h + 1) * F + DA] += hi;
            }
          }
        }
      }
      __syncthreads();
    }
    float* dst = p.part_item + (size_t)blockIdx.x * I * F;
    for (int k = threadIdx.x; k < I * F; k += blockDim.x) dst[k] = s_item[k];
  }
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kF2Warps; ++w) {
      a += s_d[w * 2];
      b += s_d[w * 2 + 1];
    }
    p.part_scalar[(size_t)blockIdx.x * 2] = a;
    p.part_scalar[(size_t)blockIdx.x * 2 + 1] = b;
  }
  if (GRAD && threadIdx.x < 4 * D) {
    float v = 0.0f;
    for (int w = 0; w < kF2Warps; ++w) v += s_t[w * 4 * D + threadIdx.x];
    p.part_table[(size_t)blockIdx.x * 4 * D + threadIdx.x] = v;
  }
}

template <int MODEL, int D>
cudaError_t launch_fused2_md(const FusedParams& p, int grid, size_t smem, bool grad, cudaStream_t st);

template <int MODEL, int D, int LPP, int NG>
static cudaError_t launch_fused2_cfg(const FusedParams& p, int grid, size_t smem, bool grad, cudaStream_t st) {
  // the opt-in to > 48 KB dynamic shared memory is per device and per function: set it on every
  // launch (a host-side attribute write, ~1 us) instead of caching it in process-wide state
  cudaError_t e;
  if (grad) {
    auto k = fused2_kernel<MODEL, D, LPP, NG, true>;
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<grid, kF2Threads, smem, st>>>(p);
  } else {
    auto k = fused2_kernel<MODEL, D, LPP, NG, false>;
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<grid, kF2Threads, smem, st>>>(p);
  }
  return cudaGetLastError();
}

#define VIBO_FUSED2_INSTANTIATE(MODEL, D)                                                             \
  template <>                                                                                         \
  cudaError_t launch_fused2_md<MODEL, D>(const FusedParams& p, int grid, size_t smem, bool grad,      \
                                         cudaStream_t st) {                                           \
    int lpp, ng;                                                                                      \
    fused_pick(p.I, &lpp, &ng);                                                                       \
    if (lpp == 8 && ng == 4) return launch_fused2_cfg<MODEL, D, 8, 4>(p, grid, smem, grad, st);       \
    if (lpp == 8 && ng == 8) return launch_fused2_cfg<MODEL, D, 8, 8>(p, grid, smem, grad, st);       \
    if (lpp == 16 && ng == 4) return launch_fused2_cfg<MODEL, D, 16, 4>(p, grid, smem, grad, st);     \
    if (lpp == 16 && ng == 8) return launch_fused2_cfg<MODEL, D, 16, 8>(p, grid, smem, grad, st);     \
    if (lpp == 32 && ng == 4) return launch_fused2_cfg<MODEL, D, 32, 4>(p, grid, smem, grad, st);     \
    return launch_fused2_cfg<MODEL, D, 32, 8>(p, grid, smem, grad, st);                               \
  }

}  // namespace vibo
