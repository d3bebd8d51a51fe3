// Item-owner single-pass ELBO forward + backward for the unconditional 1PL / 2PL encoder
// (sm_100a): the TRAINING kernel of the headline configuration when rows are wide (I >= 384).
//
// Same data movement, plan and partial-result layout as fused2_kernel (per-team rings of 1-D TMA
// bulk copies on mbarriers; every row read from HBM once; same prologue and CTA combine), but the
// work of a team is split over ITEMS instead of rows: each of the team's 128 threads owns one or
// two 4-item groups for the whole kernel, for every row.  That changes what moves through shared
// memory:
//
//   * the owner's item parameters (a', b') live in registers -- fused2 re-reads them from shared
//     memory for every row (8 B per cell for the 2PL);
//   * d ll / d z of a cell is consumed where it is produced: the item-gradient sums
//     sum_i dz_ij (theta_i, 1) are register accumulators of the owner, so there is no write-back
//     of dz over the response tile and no phase B that reads it again (8 B per cell + a pass).
//
// Shared-memory traffic per cell drops from ~32 B (stage 5, counts 5, link 16, phase B 6) to 14 B
// (stage 5, counts 5, link 4), and ~35 % of the instructions go with it -- the two things
// fused2_kernel<GRAD> is bound by.  The price is one team barrier per block of 4 rows, because
// the per-person quantities now need all four warps:
//
//   pass 1   every thread counts its cells of the 4 rows      -> per-warp (n1, nobs) in smem
//   barrier  (the only one; counts and partials are double-buffered by block parity)
//   the owner thread of each row of the PREVIOUS block combines that block's 4 partials:
//            posterior chain rule -> table gradients
//   posterior/draw of row (lane & 3) in every lane (SIMD over the 4 rows, redundantly per warp)
//   pass 2   link + ll + dz for the owner's cells of each row -> acc (registers), per-warp
//            partial sum_j dz_ij a_j of the 4 rows (one transposed 6-shuffle reduction) in smem
#pragma once

#include "vibo_fused2_kernel.cuh"

namespace vibo {

constexpr int kF3MinGroups = 96;   // below this more than a quarter of a team would idle: fused2 serves
// team scratch (bytes): packed counts, one int4 per warp, double-buffered by block parity; constant tiles
// for unused group slots (four 0.5f, four 0.0f, one all-observed mask word); g_theta partials
// [row 4][warp 4][D <= 2], double-buffered; table sums [row owner 4][A0|A1|B0|B1][D]
constexpr int kF3TeamScratch = 768;
constexpr int kF3Cnt = 0, kF3Half = 128, kF3Zero = 144, kF3Ones = 160, kF3Gth = 256, kF3Tab = 512;
// CTA constants of the posterior (floats at kFusedHdr + kF2SumOff, the setup partials' region):
//   [0,4) tau[r][d]  [4,8) mu tau [r][d]  [8] prior tau   (amax / bmax / sums stay in s_max)
constexpr int kF3Tau = 0, kF3Mt = 4, kF3Prior = 8;

// General cell group (mask and/or eps32 clamp possible): 4 cells, natural-log accumulation
// (s1 = sum (x-1) zc, s3 = sum log2(1 + E)), dz returned.  Same arithmetic as f2_group.
template <int MODEL, int D, bool FULL>
__device__ __forceinline__ void f3_group(const float4& x4, uint32_t m4, const float (&bs)[4],
                                         const float (&as)[MODEL == 1 ? 1 : D][4], const float (&th)[D], float t1,
                                         float (&gth)[D], float& s1, float& s3, float (&dzs)[4]) {
  constexpr int DA = MODEL == 1 ? 0 : D;
  const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float zs = bs[c];   // z' = -log2(e) z
    if (MODEL == 1) {
      zs += t1;
    } else {
#pragma unroll
      for (int d = 0; d < DA; ++d) zs = fmaf(th[d], as[d][c], zs);
    }
    float z = zs * -kLn2f;
    float x = xs[c];
    if (!FULL) {
      const bool o = ((m4 >> (8 * c)) & 0xffu) != 0;
      x = o ? x : 0.5f;   // missing cell -> neutral cell (x = 1/2, z = 0): gradient 0, ll = -log 2
      z = o ? z : 0.0f;
    }
    const float zc = fminf(fmaxf(z, -kLogitClamp), kLogitClamp);
    const float w = 1.0f + ex2_approx(zc * -kLog2e);
    s1 = fmaf(x - 1.0f, zc, s1);
    s3 += lg2_approx(w);
    float dz = x - rcp_approx(w);     // x - sigmoid(zc)
    dz = (z == zc) ? dz : 0.0f;       // zero gradient outside the eps32 clamp
    if (MODEL == 1) {
      gth[0] -= dz;
    } else {
#pragma unroll
      for (int d = 0; d < DA; ++d) gth[d] = fmaf(dz, as[d][c], gth[d]);   // log2(e) * sum dz a
    }
    dzs[c] = dz;
  }
}

// Sum four per-lane values over the warp with 6 shuffles: afterwards lane 8 r holds the sum of v[r].
__device__ __forceinline__ float f3_reduce4(const float (&v)[4], int lane) {
  const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
  float k0 = h16 ? v[2] : v[0], s0 = h16 ? v[0] : v[2];
  float k1 = h16 ? v[3] : v[1], s1 = h16 ? v[1] : v[3];
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  float k = h8 ? k1 : k0;
  const float s = h8 ? k0 : k1;
  k += __shfl_xor_sync(0xffffffffu, s, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;   // lane 0: v[0], lane 8: v[1], lane 16: v[2], lane 24: v[3]
}

// Refill with the grid's ragged last chunk (sizes not 16-byte multiples: copied by hand by one warp), out of
// line so that its layout arithmetic stays out of the hot loop.
template <int MODEL, int D>
__device__ __noinline__ void f3_issue_ragged(const FusedParams& p, int64_t c, int team, int s, int lane) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FusedSmem L = fused_smem_layout(p.I, D, MODEL, p.R, p.nstage, kF2Teams, f2_scratch_bytes(p.R, p.nstage, D));
  unsigned char* st = smem + L.stage_off + ((size_t)team * p.nstage + s) * L.stage_bytes;
  fused_issue_chunk<D>(p, L, c, st, reinterpret_cast<uint64_t*>(smem) + team * p.nstage + s, lane);
}
template <int MODEL, int D, int NGB>   // NGB: 4-item groups per thread (1: I <= 512, 2: I <= 1024)
__global__ void __launch_bounds__(kF2Threads, 1) fused3_kernel(const __grid_constant__ FusedParams p) {
  static_assert(MODEL == 1 || MODEL == 2, "item-owner kernel covers 1PL / 2PL");
  static_assert(D <= 2, "setup sums / accumulators are sized for D <= 2");
  constexpr int F = item_width(MODEL, D);
  constexpr int DA = MODEL == 1 ? 0 : D;
  constexpr int DP = DA > 0 ? DA : 1;
  constexpr int TW = kF2TeamWarps, NQ = kF2Teams;
  constexpr float kLn2 = 0.6931471805599453f;

  extern __shared__ __align__(128) unsigned char smem[];
  const int I = p.I, R = p.R, NS = p.nstage;
  const FusedSmem L = fused_smem_layout(I, D, MODEL, R, NS, NQ, f2_scratch_bytes(R, NS, D));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);   // [NQ * NS] (<= 16)
  float* s_max = reinterpret_cast<float*>(smem + kFusedHdr);      // amax[0..7], bmax, sum a'[0..1], sum b'
  double* s_wsum = reinterpret_cast<double*>(smem + kFusedHdr + kF2SumOff);   // [warp][3] setup partials
  float* s_param = reinterpret_cast<float*>(smem + L.params_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / TW, wt = warp % TW, tt = threadIdx.x - team * kF2TeamThreads;
  const int n_groups = I >> 2;
  const int64_t n_chunks = (p.P + R - 1) / R;
  const int64_t chunk0 = (int64_t)blockIdx.x * NQ + team, chunk_step = (int64_t)gridDim.x * NQ;
  uint64_t* t_full = full_bar + team * NS;
  unsigned char* t_stage = smem + L.stage_off + (size_t)team * NS * L.stage_bytes;
  // team scratch (the region fused2 keeps its theta rows in)
  int* t_cnt = reinterpret_cast<int*>(smem + kFusedHdr + kF2ThetaOff + team * kF3TeamScratch);

  // ---- one-time setup (as fused2_kernel) ---------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < NQ * NS; ++s) {
      mbar_init(&full_bar[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 16) s_max[threadIdx.x] = 0.0f;
  fused_draw_noise<D>(p, NQ);
  __syncthreads();
  {
    float amax[DP], bmax = 0.0f;
    double asum[DP], bsum = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
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
#pragma unroll
    for (int d = 0; d < DA; ++d) {
      const double v = warp_sum(asum[d]);
      if (lane == 0) s_wsum[warp * 3 + d] = v;
    }
    {
      const double v = warp_sum(bsum);
      if (lane == 0) s_wsum[warp * 3 + 2] = v;
    }
#pragma unroll
    for (int d = 0; d < DA; ++d) atomicMax(reinterpret_cast<int*>(&s_max[d]), __float_as_int(amax[d]));
    atomicMax(reinterpret_cast<int*>(&s_max[8]), __float_as_int(bmax));
  }
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

  // ---- per-thread state --------------------------------------------------------------------------
  float ll_acc = 0.0f, ll2_acc = 0.0f, term_acc = 0.0f;   // per lane; summed in double at the end
  f2_t accA[NGB][DP][2], accB[NGB][2];                    // item-gradient sums of the owned groups
  f2_t pa[NGB][DP][2], pb[NGB][2];                        // the owned groups' a' (per d) and b'
  bool gv[NGB];
  f2_t used2[NGB];                                        // (1, 1) for an owned group, (0, 0) for an unused slot
  uint32_t goff[NGB];
  int cells_lane = 0;
#pragma unroll
  for (int k = 0; k < NGB; ++k) {
    const int g = tt + kF2TeamThreads * k;
    gv[k] = g < n_groups;
    used2[k] = gv[k] ? pack2(1.0f, 1.0f) : pack2(0.0f, 0.0f);
    goff[k] = (uint32_t)(gv[k] ? g : 0) * 16u;
    cells_lane += gv[k] ? 4 : 0;
    accB[k][0] = accB[k][1] = pack2(0.0f, 0.0f);
    const uint32_t pp = smem_u32(s_param) + goff[k];
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      accA[k][d][0] = accA[k][d][1] = pack2(0.0f, 0.0f);
      if (DA > 0 && gv[k]) lds128_f2(pp + (uint32_t)(d * n_groups) * 16u, pa[k][d][0], pa[k][d][1]);
      else pa[k][d][0] = pa[k][d][1] = pack2(0.0f, 0.0f);
    }
    if (gv[k]) lds128_f2(pp + (uint32_t)(DA * n_groups) * 16u, pb[k][0], pb[k][1]);
    else pb[k][0] = pb[k][1] = pack2(0.0f, 0.0f);   // an unused group computes on neutral cells
  }
  const int cells_warp = __reduce_add_sync(0xffffffffu, cells_lane);
  // four neutral cells (x = 1/2, a' = b' = 0) add log2 16 = 4 to sum log2(1 + E): per block of 4 rows
  const float s3_unused = -4.0f * (float)(4 * NGB - cells_lane);
  if (tt < 4) {   // constant tiles; read after the CTA barrier below
    reinterpret_cast<float*>(t_cnt)[kF3Half / 4 + tt] = 0.5f;
    reinterpret_cast<float*>(t_cnt)[kF3Zero / 4 + tt] = 0.0f;
    t_cnt[kF3Ones / 4 + tt] = 0x01010101;
  }

  // posterior constants live in shared memory (read once per block of rows; registers are scarce)
  float* s_c = reinterpret_cast<float*>(smem + kFusedHdr + kF2SumOff);
  if (threadIdx.x < 2 * D) {
    const int r = threadIdx.x / D, d = threadIdx.x % D;
    const float mu = p.table[r * 2 * D + d], lam = p.table[r * 2 * D + D + d];
    const float tau = 1.0f / (expf(lam) + kPoeEps);
    s_c[kF3Tau + r * 2 + d] = tau;
    s_c[kF3Mt + r * 2 + d] = mu * tau;
  }
  if (threadIdx.x == 0) s_c[kF3Prior] = p.missing_policy == VIBO_MISSING_PRIOR ? 1.0f / (1.0f + kPoeEps) : 0.0f;
  if (tt < 32) reinterpret_cast<float*>(t_cnt)[kF3Tab / 4 + tt] = 0.0f;   // table sums
  __syncthreads();

  // one opaque base per region (a plain cvta result would be re-derived at every use); every address below is
  // base + a non-negative offset (compute-sanitizer mis-tracks an mbarrier wait with a negative immediate)
  uint32_t stage0 = smem_u32(t_stage), cnt0 = smem_u32(t_cnt), sm0 = smem_u32(smem);
  asm volatile("mov.u32 %0, %0;" : "+r"(stage0));
  asm volatile("mov.u32 %0, %0;" : "+r"(cnt0));
  asm volatile("mov.u32 %0, %0;" : "+r"(sm0));
  const uint32_t cst0 = sm0 + (uint32_t)kFusedHdr;                    // s_max
  const uint32_t bar0 = sm0 + (uint32_t)(team * NS) * 8u;             // this team's 'full' barriers
  const uint32_t tab0 = cnt0 + kF3Tab, half0 = cnt0 + kF3Half, zero0 = cnt0 + kF3Zero, ones0 = cnt0 + kF3Ones;
  const uint32_t sc0 = cst0 + (uint32_t)kF2SumOff;   // s_c; s_max itself is at cst0
  const uint32_t row_x = (uint32_t)I * 4u, row_m = (uint32_t)I;
  auto ldsf = [](uint32_t a) { return __uint_as_float(lds32(a)); };
  const int rr = lane & 3;                       // the row of a block this lane computes the posterior of
  const bool owner_lane = wt == rr && lane == rr;   // ... and, in one thread per row, accounts for it
  int s = 0;
  uint32_t phase = 0;
  const int n_it = chunk0 < n_chunks ? (int)((n_chunks - chunk0 + chunk_step - 1) / chunk_step) : 0;
  const int last_rows = (int)(p.P - (n_chunks - 1) * R);
  const bool owns_last = n_it > 0 && chunk0 + (int64_t)(n_it - 1) * chunk_step == n_chunks - 1;
  const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
  const f2_t one2 = pack2(1.0f, 1.0f);

  // row state of the previous block, kept by every lane for its row rr until the owner has applied the
  // posterior chain rule (after the next barrier)
  float amu[D], invS[D], sd[D], th[D], epsv[D];
  float n0f = 0.0f, n1f = 0.0f;
  bool pending = false;   // owner thread: a row of the previous block awaits its backward
  // a stage is released by the team barrier that follows its last block (every warp has left it by then):
  // the refill is issued right after that barrier, one bulk copy per warp
  bool refill = false;
  int refill_s = 0, refill_it = 0;
  // warp 0 copies the responses, warp 1 the mask, warp 2 the noise: each keeps the global address of its
  // next copy and advances it by one round of the grid's teams per refill
  const uint32_t cp_row = wt == 0 ? (uint32_t)I * 4u : (wt == 1 ? (uint32_t)I : (uint32_t)D * 4u);   // bytes per row
  const uint32_t cp_bytes = (uint32_t)R * cp_row;
  const uint32_t cp_off = wt == 0 ? 0u : (wt == 1 ? (uint32_t)p.mask_off : (uint32_t)p.eps_off);
  const char* cp_src = (wt == 0 ? reinterpret_cast<const char*>(p.resp)
                                : (wt == 1 ? reinterpret_cast<const char*>(p.mask) : reinterpret_cast<const char*>(p.eps))) +
                       (chunk0 + (int64_t)NS * chunk_step) * R * (int64_t)cp_row;
  const bool ragged_last = owns_last && last_rows != R;   // the grid's last chunk is partial and this team's
  uint32_t bp = 0;        // block parity: which copy of the counts / partials this block writes
  auto row_backward = [&](uint32_t gth_prev) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float gvv = 0.0f;
#pragma unroll
      for (int w = 0; w < TW; ++w) gvv += ldsf(gth_prev + (uint32_t)((rr * 4 + w) * D + (MODEL == 1 ? 0 : d)) * 4u);
      if (MODEL != 1) gvv *= kLn2;   // the item discriminations in registers carry log2(e)
      float g_mu, g_lv;
      if (p.form == VIBO_ELBO_KL) {
        g_mu = fmaf(p.beta, amu[d], gvv);
        g_lv = 0.5f * gvv * epsv[d] * sd[d] + 0.5f * p.beta * (invS[d] - 1.0f);
      } else {
        gvv += th[d];
        g_mu = gvv;
        g_lv = 0.5f * gvv * epsv[d] * sd[d] - 0.5f;
      }
      const float GN = g_mu * invS[d];
      const float GS = -(g_mu * amu[d] + g_lv) * invS[d];
      // table sums of this owner, [A0 | A1 | B0 | B1][D] in team scratch (touched by this thread only)
      const uint32_t tb = tab0 + (uint32_t)(rr * 4 * D + d) * 4u;
      const float add[4] = {n0f * GN, n1f * GN, n0f * GS, n1f * GS};
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float v = ldsf(tb + q4 * D * 4) + add[q4];
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(tb + q4 * D * 4), "f"(v) : "memory");
      }
    }
  };

  for (int it = 0; it < n_it; ++it) {
    mbar_wait_addr(bar0 + (uint32_t)s * 8, phase);
    // the grid's ragged last chunk is copied by hand by one warp (ordinary stores, then an mbarrier arrive):
    // put a team barrier behind the wait as well, once per kernel -- compute-sanitizer racecheck credits
    // mbarrier waits for bulk copies only, not for ordinary shared-memory stores
    if (ragged_last && it == n_it - 1) team_barrier(team);
    const uint32_t sb = stage0 + (uint32_t)s * stage_bytes;
    const int rows = (owns_last && it == n_it - 1) ? last_rows : R;

    for (int rbase = 0; rbase < rows; rbase += 4, bp ^= 1u) {
      const int nr = rows - rbase < 4 ? rows - rbase : 4;   // team-uniform
      const uint32_t xb = sb + (uint32_t)rbase * row_x;
      const uint32_t mb = sb + (uint32_t)p.mask_off + (uint32_t)rbase * row_m;
      const uint32_t eb = sb + (uint32_t)p.eps_off + (uint32_t)rbase * D * 4u;
      const uint32_t cntb = cnt0 + kF3Cnt + bp * 64u, gthb = cnt0 + kF3Gth + bp * 128u;
      // an unused group slot reads constant tiles with row stride 0 (no branches in the hot blocks)
      uint32_t xa[NGB], xs[NGB];
#pragma unroll
      for (int k = 0; k < NGB; ++k) {
        xa[k] = gv[k] ? xb + goff[k] : half0;
        xs[k] = gv[k] ? row_x : 0u;
      }

      // ---- pass 1: this warp's share of the 4 rows' counts, two 16-bit fields per word ----------
      //      (n1 rows 0|1, n1 rows 2|3, nobs rows 0|1, nobs rows 2|3); a field is at most 256 per warp
      int c_n1a = 0, c_n1b = 0, c_noa = cells_warp * 0x10001, c_nob = cells_warp * 0x10001;
      bool general_count = nr < 4;
      if (!general_count) {
        f2_t n1p[4];
        uint32_t mand = 0x01010101u;
#pragma unroll
        for (int k = 0; k < NGB; ++k) {
          const uint32_t xz = gv[k] ? xa[k] : zero0;   // zeros for an unused slot
          const uint32_t ma = gv[k] ? mb + (goff[k] >> 2) : ones0, ms = gv[k] ? row_m : 0u;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            f2_t x01, x23;
            lds128_f2(xz + (uint32_t)r * xs[k], x01, x23);
            n1p[r] = k == 0 ? add2(x01, x23) : add2(n1p[r], add2(x01, x23));
            mand &= lds32(ma + (uint32_t)r * ms);
          }
        }
        general_count = !__all_sync(0xffffffffu, mand == 0x01010101u);
        if (!general_count) {
          float sr[4];   // 0/1 responses: exact small integers
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float lo, hi;
            unpack2(n1p[r], lo, hi);
            sr[r] = lo + hi;
          }
          c_n1a = __reduce_add_sync(0xffffffffu, (int)fmaf(sr[1], 65536.0f, sr[0]));
          c_n1b = __reduce_add_sync(0xffffffffu, (int)fmaf(sr[3], 65536.0f, sr[2]));
        }
      }
      if (general_count) {   // missing cells somewhere in this warp's share, or a ragged last block
        c_n1a = c_n1b = c_noa = c_nob = 0;
#pragma unroll 1
        for (int r = 0; r < nr; ++r) {
          int n1 = 0, no = 0;
#pragma unroll
          for (int k = 0; k < NGB; ++k)
            if (gv[k]) {
              const float4 x = lds128(xb + (uint32_t)r * row_x + goff[k]);
              const uint32_t m = lds32(mb + (uint32_t)r * row_m + (goff[k] >> 2));
              const bool o0 = (m & 0xffu) != 0, o1 = (m & 0xff00u) != 0, o2 = (m & 0xff0000u) != 0,
                         o3 = (m & 0xff000000u) != 0;
              no += (int)o0 + (int)o1 + (int)o2 + (int)o3;
              n1 += (int)(o0 && x.x > 0.5f) + (int)(o1 && x.y > 0.5f) + (int)(o2 && x.z > 0.5f) +
                    (int)(o3 && x.w > 0.5f);
            }
          n1 = __reduce_add_sync(0xffffffffu, n1) << ((r & 1) * 16);
          no = __reduce_add_sync(0xffffffffu, no) << ((r & 1) * 16);
          if (r < 2) {
            c_n1a += n1;
            c_noa += no;
          } else {
            c_n1b += n1;
            c_nob += no;
          }
        }
      }
      if (lane == 0)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(cntb + (uint32_t)wt * 16u), "r"(c_n1a),
                     "r"(c_n1b), "r"(c_noa), "r"(c_nob) : "memory");
      team_barrier(team);
      if (refill) {   // team-uniform
        refill = false;
        const uint32_t dst = stage0 + (uint32_t)refill_s * stage_bytes, bar = bar0 + (uint32_t)refill_s * 8u;
        if (!(ragged_last && refill_it + NS == n_it - 1)) {   // a whole chunk: every copy a 16-byte multiple (fused_plan)
          if (lane == 0 && wt < 3) {
            // the phase cannot complete before warp 0's arrival, whatever the order of the three copies
            if (wt == 0)
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                           "r"((uint32_t)R * ((uint32_t)I * 5u + (uint32_t)D * 4u)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + cp_off),
                         "l"(cp_src), "r"(cp_bytes), "r"(bar) : "memory");
          }
          cp_src += (size_t)cp_bytes * (size_t)chunk_step;
        } else if (wt == 3) {
          f3_issue_ragged<MODEL, D>(p, chunk0 + (int64_t)(refill_it + NS) * chunk_step, team, refill_s, lane);
        }
      }

      // ---- backward of the previous block's rows (their 4 partials are complete now) -------------
      if (pending) row_backward(cnt0 + kF3Gth + (bp ^ 1u) * 128u);

      // ---- posterior and draw of row rr (every lane; lanes with equal rr agree bit for bit) -------
      float t1;
      int mode;   // 0 fast (all observed, clamp unreachable), 1 all observed, 2 general
      {
        int n1w = 0, now = 0;
#pragma unroll
        for (int w = 0; w < TW; ++w) {
          const float4 c = lds128(cntb + (uint32_t)w * 16u);
          n1w += __float_as_int(rr < 2 ? c.x : c.y);
          now += __float_as_int(rr < 2 ? c.z : c.w);
        }
        const int n1 = (n1w >> ((rr & 1) * 16)) & 0xffff, nobs = (now >> ((rr & 1) * 16)) & 0xffff;
        n1f = (float)n1;
        const float nobsf = (float)nobs;
        n0f = nobsf - n1f;
        const float nmiss = (float)I - nobsf;
        float tsum = 0.0f, term = 0.0f, bound = ldsf(cst0 + 8 * 4);
        const float prior_tau = ldsf(sc0 + kF3Prior * 4);
        const uint32_t ep = eb + (uint32_t)(rr < nr ? rr : 0) * D * 4u;
        const bool owns_row = owner_lane && rr < nr;
        pending = owns_row;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float S = fmaf(n0f, ldsf(sc0 + (kF3Tau + d) * 4), fmaf(n1f, ldsf(sc0 + (kF3Tau + 2 + d) * 4), nmiss * prior_tau));
          const float N = fmaf(n0f, ldsf(sc0 + (kF3Mt + d) * 4), n1f * ldsf(sc0 + (kF3Mt + 2 + d) * 4));
          invS[d] = rcp_approx(S);                 // = exp(logvar)
          amu[d] = N * invS[d];
          const float alv = -kLn2 * lg2_approx(S); // log(1 / S)
          sd[d] = rsqrt_approx(S);                 // exp(logvar / 2)
          epsv[d] = __uint_as_float(lds32(ep + d * 4));
          th[d] = fmaf(epsv[d], sd[d], amu[d]);
          tsum += th[d];
          bound = MODEL == 1 ? bound + fabsf(th[d]) : fmaf(fabsf(th[d]), ldsf(cst0 + d * 4), bound);
          if (p.form == VIBO_ELBO_KL) {
            term += -0.5f * (1.0f + alv - amu[d] * amu[d] - invS[d]);
          } else {
            term += -0.5f * th[d] * th[d] + 0.5f * epsv[d] * epsv[d] + 0.5f * alv;
          }
          if (owns_row && p.out_mu != nullptr) {
            const int64_t row = (chunk0 + (int64_t)it * chunk_step) * R + rbase + rr;
            p.out_mu[row * D + d] = amu[d];
            p.out_lv[row * D + d] = alv;
            p.out_theta[row * D + d] = th[d];
          }
        }
        t1 = -kLog2e * tsum;   // 1PL: z' = b' + t1
        // a row can reach the eps32 clamp only if its logit bound exceeds it (NaN bounds go exact)
        const bool full_obs = nobs == I, exact = !(bound <= kLogitClamp);
        mode = (full_obs && !exact) ? 0 : (full_obs ? 1 : 2);
        if (owns_row) term_acc += term;
      }
      const bool all_fast = nr == 4 && __all_sync(0xffffffffu, mode == 0);   // team-uniform
      if (all_fast && owner_lane) {
        // log2 units: sum_j z'_j of the row in closed form
        float zsum = ldsf(cst0 + 11 * 4);
        if (MODEL == 1) {
          zsum = fmaf((float)I, t1, zsum);
        } else {
#pragma unroll
          for (int d = 0; d < DA; ++d) zsum = fmaf(th[d], ldsf(cst0 + (9 + d) * 4), zsum);
        }
        ll2_acc += zsum;
      }

      // ---- pass 2: link, log-likelihood, dz; item sums stay in registers ------------------------
      if (all_fast) {
        // the common block: 8 independent (row, group) units, no branches.  An unused group computes
        // on neutral cells (x = 1/2 from the constant tile, a' = b' = 0: dz = 0, log2(1 + E) = 1).
        f2_t th2[4][D], t2[4], g2[4][D];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const float v = __shfl_sync(0xffffffffu, th[d], r);
            th2[r][d] = pack2(v, v);
            g2[r][d] = pack2(0.0f, 0.0f);
          }
          const float v = MODEL == 1 ? __shfl_sync(0xffffffffu, t1, r) : 0.0f;
          t2[r] = pack2(v, v);
        }
        f2_t s1p = pack2(0.0f, 0.0f);
        float s3f = s3_unused;
#pragma unroll
        for (int k = 0; k < NGB; ++k) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            f2_t x01, x23, z01 = pb[k][0], z23 = pb[k][1];
            lds128_f2(xa[k] + (uint32_t)r * xs[k], x01, x23);
            if (MODEL == 1) {
              z01 = fma2(t2[r], used2[k], z01);   // an unused group stays at z' = 0
              z23 = fma2(t2[r], used2[k], z23);
            } else {
#pragma unroll
              for (int d = 0; d < DA; ++d) {
                z01 = fma2(th2[r][d], pa[k][d][0], z01);
                z23 = fma2(th2[r][d], pa[k][d][1], z23);
              }
            }
            float z0, z1, z2, z3;
            unpack2(z01, z0, z1);
            unpack2(z23, z2, z3);
            const f2_t w01 = add2(pack2(ex2_approx(z0), ex2_approx(z1)), one2);
            const f2_t w23 = add2(pack2(ex2_approx(z2), ex2_approx(z3)), one2);
            const f2_t c = mul2(w01, w23);   // (w0 w2, w1 w3)
            float cx, cy;
            unpack2(c, cx, cy);
            const float prod = cx * cy;
            s3f += lg2_approx(prod);
            s1p = fma2(x01, z01, s1p);
            s1p = fma2(x23, z23, s1p);
            const float rp = rcp_approx(prod);
            const f2_t cs = pack2(cy * rp, cx * rp);          // (1/(w0 w2), 1/(w1 w3))
            const f2_t dz01 = sub2(x01, mul2(w23, cs));      // x - 1/w = x - sigmoid(z)
            const f2_t dz23 = sub2(x23, mul2(w01, cs));
            if (MODEL == 1) {
              g2[r][0] = add2(g2[r][0], add2(dz01, dz23));
            } else {
#pragma unroll
              for (int d = 0; d < DA; ++d) {
                g2[r][d] = fma2(dz01, pa[k][d][0], g2[r][d]);
                g2[r][d] = fma2(dz23, pa[k][d][1], g2[r][d]);
                accA[k][d][0] = fma2(dz01, th2[r][d], accA[k][d][0]);
                accA[k][d][1] = fma2(dz23, th2[r][d], accA[k][d][1]);
              }
            }
            accB[k][0] = add2(accB[k][0], dz01);
            accB[k][1] = add2(accB[k][1], dz23);
          }
        }
        {
          float lo, hi;
          unpack2(s1p, lo, hi);
          ll2_acc -= (lo + hi) + s3f;   // log2 units: - sum_j x_j z'_j - sum_j log2(1 + E_j)
        }
        // per-warp partials of sum_j dz_ij a_j for the 4 rows: lane 8 r holds row r
#pragma unroll
        for (int d = 0; d < (MODEL == 1 ? 1 : DA); ++d) {
          float gp[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float lo, hi;
            unpack2(g2[r][d], lo, hi);
            gp[r] = MODEL == 1 ? -(lo + hi) : lo + hi;
          }
          const float v = f3_reduce4(gp, lane);
          if ((lane & 7) == 0)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(gthb + (uint32_t)(((lane >> 3) * 4 + wt) * D + d) * 4u), "f"(v) : "memory");
        }
      } else {
        // a block with missing cells, rows that may reach the clamp, or fewer than 4 rows: row by row
#pragma unroll 1
        for (int r = 0; r < nr; ++r) {
          float thr[D], gth[D];
#pragma unroll
          for (int d = 0; d < D; ++d) {
            thr[d] = __shfl_sync(0xffffffffu, th[d], r);
            gth[d] = 0.0f;
          }
          const float t1r = __shfl_sync(0xffffffffu, t1, r);
          const int mr = __shfl_sync(0xffffffffu, mode, r);
          float s1 = 0.0f, s3 = 0.0f;
          int nmiss_lane = 0;
#pragma unroll
          for (int k = 0; k < NGB; ++k)
            if (gv[k]) {
              const float4 x4 = lds128(xb + (uint32_t)r * row_x + goff[k]);
              float bs[4], as[DP][4], dzs[4];
              unpack2(pb[k][0], bs[0], bs[1]);
              unpack2(pb[k][1], bs[2], bs[3]);
#pragma unroll
              for (int d = 0; d < DP; ++d) {
                unpack2(pa[k][d][0], as[d][0], as[d][1]);
                unpack2(pa[k][d][1], as[d][2], as[d][3]);
              }
              if (mr == 2) {
                const uint32_t m4 = lds32(mb + (uint32_t)r * row_m + (goff[k] >> 2));
                nmiss_lane += 4 - __popc(m4 & 0x01010101u);
                f3_group<MODEL, D, false>(x4, m4, bs, as, thr, t1r, gth, s1, s3, dzs);
              } else {
                f3_group<MODEL, D, true>(x4, 0x01010101u, bs, as, thr, t1r, gth, s1, s3, dzs);
              }
              const f2_t dz01 = pack2(dzs[0], dzs[1]), dz23 = pack2(dzs[2], dzs[3]);
#pragma unroll
              for (int d = 0; d < DA; ++d) {
                const f2_t th2 = pack2(thr[d], thr[d]);
                accA[k][d][0] = fma2(dz01, th2, accA[k][d][0]);
                accA[k][d][1] = fma2(dz23, th2, accA[k][d][1]);
              }
              accB[k][0] = add2(accB[k][0], dz01);
              accB[k][1] = add2(accB[k][1], dz23);
            }
          s3 -= (float)nmiss_lane;   // each neutral cell added log2(2) = 1
          ll_acc += s1 - kLn2 * s3;
#pragma unroll
          for (int d = 0; d < (MODEL == 1 ? 1 : DA); ++d) {
            const float v = warp_sum(gth[d]);
            if (lane == 0)
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(gthb + (uint32_t)((r * 4 + wt) * D + d) * 4u), "f"(v) : "memory");
          }
        }
      }

      // the stage's last block of rows has been read: refill it after the next team barrier
      if (rbase + 4 >= rows && it + NS < n_it) {
        refill = true;
        refill_s = s;
        refill_it = it;
      }
    }
    if (++s == NS) {
      s = 0;
      phase ^= 1u;
    }
  }
  if (n_it > 0) {   // the last block's rows
    team_barrier(team);
    if (pending) row_backward(cnt0 + kF3Gth + (bp ^ 1u) * 128u);
  }

  // ---- CTA-level combine (deterministic order; as fused2_kernel) ---------------------------------
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
      // the row owners' table sums (lane r of warp r of the team), fetched before the stages are reused
      const bool own = owner_lane;
      const float v0 = warp_sum(own ? ldsf(tab0 + (uint32_t)(rr * 4 * D + d) * 4u) : 0.0f),
                  v1 = warp_sum(own ? ldsf(tab0 + (uint32_t)(rr * 4 * D + D + d) * 4u) : 0.0f),
                  v2 = warp_sum(own ? ldsf(tab0 + (uint32_t)(rr * 4 * D + 2 * D + d) * 4u) : 0.0f),
                  v3 = warp_sum(own ? ldsf(tab0 + (uint32_t)(rr * 4 * D + 3 * D + d) * 4u) : 0.0f);
      if (lane == 0) {
        s_t[warp * 4 * D + d] = v0;
        s_t[warp * 4 * D + D + d] = v1;
        s_t[warp * 4 * D + 2 * D + d] = v2;
        s_t[warp * 4 * D + 3 * D + d] = v3;
      }
    }
  }
  for (int k = threadIdx.x; k < I * F; k += blockDim.x) s_item[k] = 0.0f;
  __syncthreads();
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
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kF2Warps; ++w) {
      a += s_d[w * 2];
      b += s_d[w * 2 + 1];
    }
    p.part_scalar[(size_t)blockIdx.x * 2] = a;
    p.part_scalar[(size_t)blockIdx.x * 2 + 1] = b;
  }
  if (threadIdx.x < 4 * D) {
    float v = 0.0f;
    for (int w = 0; w < kF2Warps; ++w) v += s_t[w * 4 * D + threadIdx.x];
    p.part_table[(size_t)blockIdx.x * 4 * D + threadIdx.x] = v;
  }
}

template <int MODEL, int D>
cudaError_t launch_fused3_md(const FusedParams& p, int grid, size_t smem, cudaStream_t st);

template <int MODEL, int D, int NGB>
static cudaError_t launch_fused3_cfg(const FusedParams& p, int grid, size_t smem, cudaStream_t st) {
  auto k = fused3_kernel<MODEL, D, NGB>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<grid, kF2Threads, smem, st>>>(p);
  return cudaGetLastError();
}

#define VIBO_FUSED3_INSTANTIATE(MODEL, D)                                                              \
  template <>                                                                                          \
  cudaError_t launch_fused3_md<MODEL, D>(const FusedParams& p, int grid, size_t smem, cudaStream_t st) { \
    if ((p.I >> 2) <= kF2TeamThreads) return launch_fused3_cfg<MODEL, D, 1>(p, grid, smem, st);        \
    return launch_fused3_cfg<MODEL, D, 2>(p, grid, smem, st);                                          \
  }

}  // namespace vibo
