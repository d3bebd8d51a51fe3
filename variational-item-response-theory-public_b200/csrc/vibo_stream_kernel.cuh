// "Slab-stream" kernels: the general (decomposed) passes of the ELBO path for
// every variant the single-pass kernels do not cover -- conditional posterior,
// ability_dim up to 8, 3PL, any number of items (no alignment requirement on
// I), missing data -- at streaming speed.
//
//   encode_stream_kernel      product-of-experts sums   (models.py:596-629, utils.py:105-113);
//                             unconditional: two counts per person (also vibo_person_counts)
//   link_stream_kernel        link + Bernoulli log-lik  (models.py:729-766, utils.py:46-49)
//                             + d/d ability, d/d item_feat
//   encode_bwd_stream_kernel  per-(response value, item) sums of the expert-table gradient
//   encode_mma_kernel         conditional product-of-experts sums on the tensor cores
//                             (bf16 x 3 split, mma.sync m16n8k16)
//   encode_bwd_mma_kernel     conditional expert-table gradient sums on the tensor cores
//                             (tf32 x 2 split, mma.sync m16n8k8)
//
// Shared skeleton.  A persistent CTA streams chunks of R rows (response f32 +
// mask u8 + the per-person arrays the pass needs) through a ring of NS
// shared-memory stages with 1-D TMA bulk copies completing on mbarriers; R is
// a multiple of 16 / gcd(I, 16), which makes every block of a chunk a 16-byte
// multiple at a 16-byte aligned address whatever I is.  Warp w owns the item
// slab [w*32*M, (w+1)*32*M): lane l owns items w*32*M + m*32 + l (m < M) for
// EVERY row, so item parameters, expert-table entries and per-item gradient
// sums live in registers for the whole kernel (no atomics, deterministic) and
// shared memory holds nothing but row data.  Per-person sums over items are
// reduced across the 32 lanes NR rows at a time with a transpose-reduce (about
// 4 instructions per value instead of 10 for a butterfly per value), then
// across the slabs through shared memory in a fixed order.
#pragma once

#include <type_traits>

#include "vibo_fused2_kernel.cuh"

namespace vibo {

constexpr int kStreamMaxStages = 8;
constexpr int kStreamMaxParr = 4;

struct StreamParams {
  int64_t P;
  int I;
  int R;        // rows per stage
  int NS;       // ring depth
  int n_parr;   // per-person (P, D) float arrays staged with the rows
  int D;
  uint32_t mask_off, parr_off, stage_bytes;  // byte offsets inside a stage
  uint32_t red_off, info_off, stage_off;      // byte offsets inside dynamic shared memory
  const float* resp;
  const uint8_t* mask;
  const float* parr[kStreamMaxParr];
};

// Host-side launch plan (vibo_stream.cu).
struct StreamPlan {
  bool ok = false;
  int M = 0, NW = 0, R = 0, NS = 0, grid = 0;
  size_t smem = 0;
  uint32_t mask_off = 0, parr_off = 0, stage_bytes = 0, red_off = 0, info_off = 0, stage_off = 0;
};
cudaError_t stream_link_run1(const StreamPlan& pl, const StreamParams& p, int D, const float* item_feat,
                             double* part_ll, float* g_ability, float* part_g, bool grad, cudaStream_t st);
cudaError_t stream_link_run2(const StreamPlan& pl, const StreamParams& p, int D, const float* item_feat,
                             double* part_ll, float* g_ability, float* part_g, bool grad, cudaStream_t st);
cudaError_t stream_link_run3(const StreamPlan& pl, const StreamParams& p, int D, const float* item_feat,
                             double* part_ll, float* g_ability, float* part_g, bool grad, cudaStream_t st);

cudaError_t stream_encode_mma_run(int D, int KS, bool even, int grid, size_t smem, const StreamParams& p,
                                  int missing_policy, const float* table, float* mu, float* lv, float* S,
                                  cudaStream_t st);
cudaError_t stream_encode_bwd_mma_run(int D, int MT, int grid, size_t smem, const StreamParams& p, float* part,
                                      cudaStream_t st);

// ---- staging -----------------------------------------------------------------
// Thread 0: arm `bar` and issue the bulk copies of the full chunk c.
__device__ __forceinline__ void stream_issue(const StreamParams& p, int64_t c, unsigned char* st, uint64_t* bar) {
  const int64_t row0 = c * p.R;
  const uint32_t b_resp = (uint32_t)p.R * p.I * 4, b_mask = (uint32_t)p.R * p.I, b_parr = (uint32_t)p.R * p.D * 4;
  mbar_expect_tx(bar, b_resp + b_mask + (uint32_t)p.n_parr * b_parr);
  bulk_g2s(st, p.resp + row0 * p.I, b_resp, bar);
  bulk_g2s(st + p.mask_off, p.mask + row0 * p.I, b_mask, bar);
  for (int a = 0; a < p.n_parr; ++a)
    bulk_g2s(st + p.parr_off + (size_t)a * b_parr, p.parr[a] + row0 * p.D, b_parr, bar);
}
// Whole CTA: copy the ragged last chunk (rows < R; its sizes need not be
// 16-byte multiples) by hand.
__device__ __forceinline__ void stream_copy_ragged(const StreamParams& p, int64_t c, unsigned char* st, int rows) {
  const int64_t row0 = c * p.R;
  const float* gr = p.resp + row0 * p.I;
  float* sr = reinterpret_cast<float*>(st);
  for (int k = threadIdx.x; k < rows * p.I; k += blockDim.x) sr[k] = gr[k];
  const uint8_t* gm = p.mask + row0 * p.I;
  uint8_t* sm = st + p.mask_off;
  for (int k = threadIdx.x; k < rows * p.I; k += blockDim.x) sm[k] = gm[k];
  for (int a = 0; a < p.n_parr; ++a) {
    const float* ga = p.parr[a] + row0 * p.D;
    float* sa = reinterpret_cast<float*>(st + p.parr_off + (size_t)a * p.R * p.D * 4);
    for (int k = threadIdx.x; k < rows * p.D; k += blockDim.x) sa[k] = ga[k];
  }
}

// ---- transpose-reduce ----------------------------------------------------------
// Each lane holds NR values (one per row).  Returns, in every lane, the sum
// over the 32 lanes of the value of row stream_row<NR>(lane).
template <int NR>
__device__ __forceinline__ int stream_row(int lane) {
  return NR == 8 ? ((lane >> 2) & 7) : ((lane >> 3) & 3);
}
template <int NR>
__device__ __forceinline__ float transpose_reduce(const float (&v)[NR], int lane) {
  static_assert(NR == 8 || NR == 4, "NR must be 4 or 8");
  float w[NR / 2];
  const bool b4 = (lane & 16) != 0;
#pragma unroll
  for (int i = 0; i < NR / 2; ++i) {
    const float send = b4 ? v[i] : v[i + NR / 2], keep = b4 ? v[i + NR / 2] : v[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float u[NR / 4];
  const bool b3 = (lane & 8) != 0;
#pragma unroll
  for (int i = 0; i < NR / 4; ++i) {
    const float send = b3 ? w[i] : w[i + NR / 4], keep = b3 ? w[i + NR / 4] : w[i];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float t;
  if (NR == 8) {
    const bool b2 = (lane & 4) != 0;
    const float send = b2 ? u[0] : u[NR == 8 ? 1 : 0], keep = b2 ? u[NR == 8 ? 1 : 0] : u[0];
    t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  } else {
    t = u[0] + __shfl_xor_sync(0xffffffffu, u[0], 4);
  }
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  return t;
}

// ---- shared prologue -----------------------------------------------------------
struct StreamCtx {
  uint64_t* bar;
  unsigned char* stages;
  float* red;     // [2][R][NW][Q]
  float* info;    // [R][..] per-row scratch
};
__device__ __forceinline__ StreamCtx stream_setup(const StreamParams& p, unsigned char* smem) {
  StreamCtx c;
  c.bar = reinterpret_cast<uint64_t*>(smem);
  c.red = reinterpret_cast<float*>(smem + p.red_off);
  c.info = reinterpret_cast<float*>(smem + p.info_off);
  c.stages = smem + p.stage_off;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.NS; ++s) mbar_init(&c.bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int64_t n_full = p.P / p.R;  // chunks [0, n_full) are complete
    for (int s = 0; s < p.NS; ++s) {
      const int64_t ch = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (ch < n_full) stream_issue(p, ch, c.stages + (size_t)s * p.stage_bytes, &c.bar[s]);
    }
  }
  __syncthreads();
  return c;
}

// Fast logistic cell (1PL / 2PL): ll = (x-1) zc - log(1 + exp(-zc)), zc = clamp(z),
// d ll / d z = x - sigmoid(zc), zero outside the eps32 clamp.
__device__ __forceinline__ void cell_logistic_fast(float z, float x, float& ll, float& dz) {
  const float zc = fminf(fmaxf(z, -kLogitClamp), kLogitClamp);
  const float w = 1.0f + ex2_approx(zc * -kLog2e);
  ll = fmaf(x - 1.0f, zc, -kLn2f * lg2_approx(w));
  const float g = x - rcp_approx(w);
  dz = (z == zc) ? g : 0.0f;
}
// Fast 3PL cell: p = g + (1-g) sigmoid(z) clamped to [eps32, 1-eps32] (models.py:753-766).
__device__ __forceinline__ void cell_3pl_fast(float z, float g, bool x1, float& ll, float& dz, float& dgam) {
  const float e = ex2_approx(fabsf(z) * -kLog2e);
  const float r = rcp_approx(1.0f + e);
  const float er = e * r;
  const float s = z >= 0.0f ? r : er;    // sigmoid(z)
  const float sn = z >= 0.0f ? er : r;   // sigmoid(-z)
  const float p = fmaf(1.0f - g, s, g);
  const float q = (1.0f - g) * sn;       // 1 - p with full relative precision
  const bool inside = (p >= kEps32) && (q >= kEps32);
  const float u = x1 ? p : q;
  const float uc = fminf(fmaxf(u, kEps32), 1.0f - kEps32);
  ll = kLn2f * lg2_approx(uc);
  float dp = rcp_approx(uc);
  dp = x1 ? dp : -dp;
  dp = inside ? dp : 0.0f;
  const float t = dp * sn * (1.0f - g);
  dz = t * s;
  dgam = t * g;
}

// ---------------------------------------------------------------------------
// encode: S_i = sum_j tau_ij, N_i = sum_j mu_ij tau_ij  ->  mu, logvar, S
// ---------------------------------------------------------------------------
// Unconditional encoder (COND = false): the posterior depends on the row only
// through the counts (n1, n_missing), so only those two sums are reduced.
template <int D, int M, int NR, bool COND>
__global__ void __launch_bounds__(512) encode_stream_kernel(const __grid_constant__ StreamParams p,
                                                            int missing_policy, const float* __restrict__ table,
                                                            float* __restrict__ out_mu, float* __restrict__ out_lv,
                                                            float* __restrict__ out_S,
                                                            float* __restrict__ out_counts) {
  // out_counts (unconditional only): also write (n1, n_observed) per person -- the sufficient statistics the
  // backward pass needs (encode_bwd_counts_kernel) -- or, with out_mu == nullptr, only those
  extern __shared__ __align__(128) unsigned char smem[];
  const StreamCtx cx = stream_setup(p, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  const int I = p.I, R = p.R, NS = p.NS;
  constexpr int Q = COND ? 2 * D : 2;
  constexpr int DT = COND ? D : 1;   // width of the lane-owned expert registers

  // lane-owned expert entries: tau0, (tau1 - tau0), mu0 tau0, (mu1 tau1 - mu0 tau0)
  int joff[M];
  bool valid[M];
  float t0[M][DT], dt[M][DT], n0[M][DT], dn[M][DT];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int j = (warp * M + m) * 32 + lane;
    valid[m] = j < I;
    joff[m] = min(j, I - 1);
    const int jt = joff[m], It = I;
#pragma unroll
    for (int d = 0; d < (COND ? D : 0); ++d) {
      float tv[2], nv[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float mu = table[((size_t)r * It + jt) * 2 * D + d];
        const float lam = table[((size_t)r * It + jt) * 2 * D + D + d];
        tv[r] = 1.0f / (expf(lam) + kPoeEps);
        nv[r] = mu * tv[r];
      }
      t0[m][d] = valid[m] ? tv[0] : 0.0f;
      dt[m][d] = valid[m] ? tv[1] - tv[0] : 0.0f;
      n0[m][d] = valid[m] ? nv[0] : 0.0f;
      dn[m][d] = valid[m] ? nv[1] - nv[0] : 0.0f;
    }
  }
  float baseS[DT], baseN[DT];   // this lane's share of sum_j tau0_j, sum_j mu0_j tau0_j
#pragma unroll
  for (int d = 0; d < (COND ? D : 0); ++d) {
    baseS[d] = baseN[d] = 0.0f;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      baseS[d] += t0[m][d];
      baseN[d] += n0[m][d];
    }
  }
  const float prior_tau = (missing_policy == VIBO_MISSING_PRIOR) ? 1.0f / (1.0f + kPoeEps) : 0.0f;

  const int64_t n_chunks = (p.P + R - 1) / R, n_full = p.P / R;
  int s = 0, buf = 0;
  uint32_t phase = 0;
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    unsigned char* st = cx.stages + (size_t)s * p.stage_bytes;
    const int rows = (int)((p.P - c * R < R) ? p.P - c * R : R);
    if (c < n_full) {
      mbar_wait(&cx.bar[s], phase);
    } else {
      stream_copy_ragged(p, c, st, rows);
      __syncthreads();
    }
    const float* sx = reinterpret_cast<const float*>(st);
    const uint8_t* sm = st + p.mask_off;
    float* red = cx.red + (size_t)buf * R * NW * Q;
    for (int r0 = 0; r0 < rows; r0 += NR) {
      float part[Q][NR];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const int r = r0 + rr;
        float S[DT], N[DT], c1 = 0.0f, cm = 0.0f;
#pragma unroll
        for (int d = 0; d < DT; ++d) S[d] = N[d] = 0.0f;
        if (r < rows) {
          float x[M];
          bool o[M];
          bool all_obs = true;
#pragma unroll
          for (int m = 0; m < M; ++m) {
            x[m] = sx[r * I + joff[m]];
            o[m] = sm[r * I + joff[m]] != 0;
            all_obs = all_obs && (o[m] || !valid[m]);
          }
          if (!COND) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
              c1 += (valid[m] && o[m] && x[m] > 0.5f) ? 1.0f : 0.0f;
              cm += (valid[m] && !o[m]) ? 1.0f : 0.0f;
            }
          } else if (__all_sync(0xffffffffu, all_obs)) {
#pragma unroll
            for (int d = 0; d < DT; ++d) {
              S[d] = baseS[d];
              N[d] = baseN[d];
#pragma unroll
              for (int m = 0; m < M; ++m) {
                S[d] = fmaf(x[m], dt[m][d], S[d]);
                N[d] = fmaf(x[m], dn[m][d], N[d]);
              }
            }
          } else {
            float nmiss = 0.0f;
#pragma unroll
            for (int m = 0; m < M; ++m) {
              const float wo = o[m] ? 1.0f : 0.0f, wx = o[m] ? x[m] : 0.0f;
              nmiss += (valid[m] && !o[m]) ? 1.0f : 0.0f;
#pragma unroll
              for (int d = 0; d < DT; ++d) {
                S[d] = fmaf(wo, t0[m][d], fmaf(wx, dt[m][d], S[d]));
                N[d] = fmaf(wo, n0[m][d], fmaf(wx, dn[m][d], N[d]));
              }
            }
#pragma unroll
            for (int d = 0; d < DT; ++d) S[d] = fmaf(nmiss, prior_tau, S[d]);
          }
        }
        if (COND) {
#pragma unroll
          for (int d = 0; d < DT; ++d) {
            part[d][rr] = S[d];
            part[(COND ? D : 0) + d][rr] = N[d];
          }
        } else {
          part[0][rr] = c1;
          part[1][rr] = cm;
        }
      }
      const int row = r0 + stream_row<NR>(lane);
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        const float t = transpose_reduce<NR>(part[k], lane);
        if ((lane & (NR == 8 ? 3 : 7)) == 0 && row < rows) red[((size_t)row * NW + warp) * Q + k] = t;
      }
    }
    __syncthreads();   // every warp is done with stage s; red[buf] is complete
    if (threadIdx.x == 0) {
      const int64_t cn = c + (int64_t)NS * gridDim.x;
      if (cn < n_full) stream_issue(p, cn, st, &cx.bar[s]);
    }
    for (int t = threadIdx.x; t < rows * D; t += blockDim.x) {
      const int r = t / D, d = t % D;
      float sv = 0.0f, nv = 0.0f;
      if (COND) {
        for (int w = 0; w < NW; ++w) {
          sv += red[((size_t)r * NW + w) * Q + d];
          nv += red[((size_t)r * NW + w) * Q + D + d];
        }
      } else {
        float n1 = 0.0f, nm = 0.0f;
        for (int w = 0; w < NW; ++w) {
          n1 += red[((size_t)r * NW + w) * Q];
          nm += red[((size_t)r * NW + w) * Q + 1];
        }
        if (out_counts != nullptr) {
          if (d == 0) {
            out_counts[(c * R + r) * 2] = n1;
            out_counts[(c * R + r) * 2 + 1] = (float)I - nm;
          }
          if (out_mu == nullptr) continue;   // counts only (vibo_person_counts)
        }
        const float nz = (float)I - nm - n1;
        const float mu0 = table[d], mu1 = table[2 * D + d];
        const float ta0 = 1.0f / (expf(table[D + d]) + kPoeEps), ta1 = 1.0f / (expf(table[3 * D + d]) + kPoeEps);
        sv = fmaf(nz, ta0, fmaf(n1, ta1, nm * prior_tau));
        nv = fmaf(nz, mu0 * ta0, n1 * (mu1 * ta1));
      }
      const int64_t row = c * R + r;
      out_mu[row * D + d] = nv / sv;
      out_lv[row * D + d] = logf(1.0f / sv);
      if (out_S) out_S[row * D + d] = sv;
    }
    buf ^= 1;
    if (++s == NS) {
      s = 0;
      phase ^= 1u;
    }
  }
}

// ---------------------------------------------------------------------------
// link + log-likelihood, d LL / d ability, d LL / d item_feat
// ---------------------------------------------------------------------------
// staged per-person array 0: ability (P, D).  part_gitem: [grid][I*F] with
// d LL/d a = -sum dz theta, d LL/d b = sum dz, d LL/d gamma = sum dgam.
//
// One cell of the link.  3PL: p = g + (1-g) sigmoid(z), clamped to [eps32, 1-eps32]
// (models.py:753-766; the clamp is torch.distributions' clamp_probs).  With
// E = exp(-max(z,-80)), r = 1/(1+E): sigmoid(z) = r, sigmoid(-z) = E r (full relative
// precision on both sides), u = x ? p : 1-p, ll = log clamp(u), and
//   t0 = d ll/d u * sigmoid(-z) (signed, zero outside the clamp):
//   d ll/d z = t0 sigmoid(z) (1-g),   d ll/d gamma = t0 g (1-g)   (g (1-g) is applied per item at the end).
// wv: 1 for a lane-owned item that exists, 0 for padding (j >= I): d ll/d z is scaled by
// it so that padding lanes drop out of the per-person sums without selects.
template <int MODEL>
__device__ __forceinline__ void link_cell(float z, float x, float g, float omg, float wv, float& ll, float& dz,
                                          float& t0) {
  if (MODEL == 3) {
    const float zc = fmaxf(z, -80.0f);
    const float e = ex2_approx(zc * -kLog2e);
    const float r = rcp_approx(1.0f + e);
    const float sn = e * r;
    const float pp = fmaf(omg, r, g), q = omg * sn;
    const bool x1 = x > 0.5f;
    const float u = x1 ? pp : q;
    const float uc = fminf(fmaxf(u, kEps32), 1.0f - kEps32);
    ll = kLn2f * lg2_approx(uc);
    float dp = rcp_approx(uc) * fmaf(x, 2.0f, -1.0f);   // +1/p for x = 1, -1/(1-p) for x = 0
    dp = (uc == u) ? dp : 0.0f;
    t0 = dp * sn;
    dz = t0 * r * (omg * wv);   // omg * wv is loop-invariant per item
  } else {
    cell_logistic_fast(z, x, ll, dz);
    dz *= wv;
    t0 = 0.0f;
  }
}

// Two 3PL cells (one lane's item pair) in packed f32x2 arithmetic, with TWO MUFU operations per cell instead of
// four.  With E = exp(-zc), w = 1 + E, v = 1 + g E:  p = g + (1-g)/(1+E) = v / w, so ONE reciprocal
// R = 1 / (w v) yields both sigmoid(z) = R v and 1 / v = R w (so sigmoid(-z) / p = E / v); for x = 0 the factor
// d ll/d u * sigmoid(-z) = -sigmoid(-z) / ((1-g) sigmoid(-z)) is the per-item constant -1/(1-g) (nio2).
// zc = max(z, -40) keeps w v finite; below -40 sigmoid(z) < 4.3e-18 moves p = g + (1-g) sigmoid(z) by less
// than 4e-11 of the eps32 floor the clamp holds it above, so value and gradient are unchanged.
// Returns the clamped probabilities of the observed responses (u2: the caller multiplies them over four
// persons before one lg2), and with GRAD t02 = d ll/d u * sigmoid(-z) (zero outside the clamp) and
// dz2 = t0 sigmoid(z) (1-g), as link_cell (no validity weight: the caller zeroes a for padding lanes).
template <bool GRAD>
__device__ __forceinline__ void link_pair_3pl(f2_t z2, float x0, float x1, f2_t g2, f2_t omg2, f2_t nio2, f2_t& u2,
                                              f2_t& dz2, f2_t& t02) {
  float z0, z1;
  unpack2(z2, z0, z1);
  const f2_t one2 = pack2(1.0f, 1.0f);
  float s0, s1;
  unpack2(mul2(pack2(fmaxf(z0, -40.0f), fmaxf(z1, -40.0f)), pack2(-kLog2e, -kLog2e)), s0, s1);
  const f2_t e2 = pack2(ex2_approx(s0), ex2_approx(s1));
  const f2_t w2 = add2(e2, one2), v2 = fma2(g2, e2, one2);
  float wv0, wv1;
  unpack2(mul2(w2, v2), wv0, wv1);
  const f2_t R2 = pack2(rcp_approx(wv0), rcp_approx(wv1));
  const f2_t r2 = mul2(R2, v2);       // sigmoid(z)
  const f2_t sn2 = mul2(e2, r2);      // sigmoid(-z)
  float p0, p1, q0, q1;
  unpack2(fma2(omg2, r2, g2), p0, p1);
  unpack2(mul2(omg2, sn2), q0, q1);   // 1 - p with full relative precision
  const bool b0 = x0 > 0.5f, b1 = x1 > 0.5f;
  const float u0 = b0 ? p0 : q0, u1 = b1 ? p1 : q1;
  const float c0 = fminf(fmaxf(u0, kEps32), 1.0f - kEps32), c1 = fminf(fmaxf(u1, kEps32), 1.0f - kEps32);
  u2 = pack2(c0, c1);
  if (GRAD) {
    float t0, t1, n0, n1;
    unpack2(mul2(e2, mul2(w2, R2)), t0, t1);   // x = 1: sigmoid(-z) / p = (E / w) / (v / w) = E / v = E w R
    unpack2(nio2, n0, n1);
    t0 = b0 ? t0 : n0;
    t1 = b1 ? t1 : n1;
    t0 = (c0 == u0) ? t0 : 0.0f;
    t1 = (c1 == u1) ? t1 : 0.0f;
    t02 = pack2(t0, t1);
    dz2 = mul2(mul2(t02, r2), omg2);
  } else {
    t02 = pack2(0.0f, 0.0f);
    dz2 = t02;
  }
}

template <int MODEL, int D, int M, int NR, bool GRAD>
__global__ void __launch_bounds__(512) link_stream_kernel(const __grid_constant__ StreamParams p,
                                                          const float* __restrict__ item_feat,
                                                          double* __restrict__ part_ll,
                                                          float* __restrict__ g_ability,
                                                          float* __restrict__ part_gitem) {
  constexpr int F = item_width(MODEL, D);
  constexpr int DA = MODEL == 1 ? 1 : D;   // width of the discrimination registers
  constexpr int Q = D;
  constexpr int MP = (M + 1) / 2;          // the lane's items are handled as f32x2 pairs
  extern __shared__ __align__(128) unsigned char smem[];
  const StreamCtx cx = stream_setup(p, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  const int I = p.I, R = p.R, NS = p.NS;
  int* sflag = reinterpret_cast<int*>(smem + 64);   // [2]: "stage has a missing cell", by chunk parity

  int joff[2 * MP];
  bool valid[2 * MP];
  f2_t a2[MP][DA], b2[MP], acc2[MP][GRAD ? F : 1];   // pairs over the lane's items (2k, 2k+1)
  float gs[2 * MP], omg[2 * MP], wv[2 * MP];
  f2_t g2k[MP], omg2k[MP], nio2k[MP], wl2k[MP];   // 3PL only
#pragma unroll
  for (int k = 0; k < MP; ++k) {
    float av[2][DA], bv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = 2 * k + h;
      const int j = (warp * M + m) * 32 + lane;
      valid[m] = m < M && j < I;
      joff[m] = min(j, I - 1);
      if (MODEL == 1) {
        bv[h] = item_feat[joff[m]];
        av[h][0] = 0.0f;
        gs[m] = 0.0f;
      } else {
        // 3PL: a padding lane carries a = 0, so its d ll/d z drops out of the per-person sums without a weight
#pragma unroll
        for (int d = 0; d < D; ++d)
          av[h][d] = (MODEL == 3 && !valid[m]) ? 0.0f : item_feat[(size_t)joff[m] * F + d];
        bv[h] = item_feat[(size_t)joff[m] * F + D];
        gs[m] = MODEL == 3 ? 1.0f / (1.0f + expf(-item_feat[(size_t)joff[m] * F + D + 1])) : 0.0f;
      }
      omg[m] = 1.0f - gs[m];
      wv[m] = valid[m] ? 1.0f : 0.0f;
    }
    b2[k] = pack2(bv[0], bv[1]);
#pragma unroll
    for (int d = 0; d < DA; ++d) a2[k][d] = pack2(-av[0][d], -av[1][d]);   // the NEGATED discriminations
#pragma unroll
    for (int f = 0; f < (GRAD ? F : 1); ++f) acc2[k][f] = pack2(0.0f, 0.0f);
    if (MODEL == 3) {   // per-pair constants of link_pair_3pl
      g2k[k] = pack2(gs[2 * k], gs[2 * k + 1]);
      omg2k[k] = pack2(omg[2 * k], omg[2 * k + 1]);
      nio2k[k] = pack2(-1.0f / omg[2 * k], -1.0f / omg[2 * k + 1]);
      wl2k[k] = pack2(wv[2 * k] * kLn2f, wv[2 * k + 1] * kLn2f);
    }
  }
  float ll_lane = 0.0f;   // flushed into a double once per stage
  double ll_acc = 0.0;

  const int64_t n_chunks = (p.P + R - 1) / R, n_full = p.P / R;
  // 16 mask bytes per thread per load: does the stage hold a missing cell?
  auto scan_missing = [&](const unsigned char* st_, int rows_) {
    bool miss = false;
    const uint4* m4 = reinterpret_cast<const uint4*>(st_ + p.mask_off);
    const int n16 = (rows_ * I) >> 4;
    for (int k = threadIdx.x; k < n16; k += blockDim.x) {
      const uint4 w = m4[k];
      const uint32_t zz = ((w.x - 0x01010101u) & ~w.x) | ((w.y - 0x01010101u) & ~w.y) |
                          ((w.z - 0x01010101u) & ~w.z) | ((w.w - 0x01010101u) & ~w.w);
      miss = miss || (zz & 0x80808080u) != 0;
    }
    return miss;
  };
  if (threadIdx.x < 2) sflag[threadIdx.x] = 0;
  __syncthreads();
  if ((int64_t)blockIdx.x < n_full) {
    mbar_wait(&cx.bar[0], 0);
    if (scan_missing(cx.stages, R)) sflag[0] = 1;
  } else if (threadIdx.x == 0) {
    sflag[0] = 1;   // ragged first chunk: take the masked path
  }
  __syncthreads();

  int s = 0, buf = 0, par = 0;
  uint32_t phase = 0;
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    unsigned char* st = cx.stages + (size_t)s * p.stage_bytes;
    const int rows = (int)((p.P - c * R < R) ? p.P - c * R : R);
    if (c < n_full) {
      mbar_wait(&cx.bar[s], phase);
    } else {
      stream_copy_ragged(p, c, st, rows);
      __syncthreads();
    }
    const bool has_missing = sflag[par] != 0;
    const float* sx = reinterpret_cast<const float*>(st);
    const uint8_t* sm = st + p.mask_off;
    const float* sth = reinterpret_cast<const float*>(st + p.parr_off);
    float* red = cx.red + (size_t)buf * R * NW * Q;
    // One tile of NR rows.  MASKED: consult the mask / item validity / row count;
    // otherwise every cell of the tile is a valid observed cell (no selects).
    auto tile = [&](int r0, auto omask_tag, auto ragged_tag) {
      constexpr bool OMASK = decltype(omask_tag)::value;     // consult the mask bytes
      constexpr bool RAGGED = decltype(ragged_tag)::value;   // the tile may run past `rows`
      float part[GRAD ? Q : 1][NR];
      const float* xr = sx + r0 * I;
      const uint8_t* mr = sm + r0 * I;
      const float* tr = sth + r0 * D;
      // 3PL: product of the clamped probabilities of up to four rows per item, ONE lg2 per item and four rows
      // (four factors >= eps32 stay far above the smallest normal float)
      f2_t prod2[MP];
#pragma unroll
      for (int k = 0; k < MP; ++k) prod2[k] = pack2(1.0f, 1.0f);
#pragma unroll
      for (int rr = 0; rr < NR; ++rr, xr += I, mr += I, tr += D) {
        f2_t gth2[D];   // the lane's two halves of d LL / d theta; added at the end of the row
#pragma unroll
        for (int d = 0; d < D; ++d) gth2[d] = pack2(0.0f, 0.0f);
        if (!RAGGED || r0 + rr < rows) {
          f2_t th2[D];
          float tsum = 0.0f;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const float thd = tr[d];
            th2[d] = pack2(thd, thd);
            tsum += thd;
          }
#pragma unroll
          for (int k = 0; k < MP; ++k) {
            f2_t z2 = b2[k];
            if (MODEL == 1) {
              z2 = add2(z2, pack2(tsum, tsum));
            } else {
#pragma unroll
              for (int d = 0; d < D; ++d) z2 = fma2(th2[d], a2[k][d], z2);   // z = b - theta . a
            }
            f2_t dz2, t02;
            if (MODEL == 3) {
              f2_t u2;
              link_pair_3pl<GRAD>(z2, xr[joff[2 * k]], xr[joff[2 * k + 1]], g2k[k], omg2k[k], nio2k[k], u2, dz2,
                                  t02);
              if (OMASK) {
                const bool o0 = mr[joff[2 * k]] != 0, o1 = mr[joff[2 * k + 1]] != 0;
                float a0, a1, d0, d1, e0, e1;
                unpack2(u2, a0, a1);
                u2 = pack2(o0 ? a0 : 1.0f, o1 ? a1 : 1.0f);
                if (GRAD) {
                  unpack2(dz2, d0, d1);
                  unpack2(t02, e0, e1);
                  dz2 = pack2(o0 ? d0 : 0.0f, o1 ? d1 : 0.0f);
                  t02 = pack2(o0 ? e0 : 0.0f, o1 ? e1 : 0.0f);
                }
              }
              prod2[k] = mul2(prod2[k], u2);
            } else {
              float z[2], dz[2];
              unpack2(z2, z[0], z[1]);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int m = 2 * k + h;
                float ll1, t0u;
                link_cell<MODEL>(z[h], xr[joff[m]], gs[m], omg[m], wv[m], ll1, dz[h], t0u);
                if (OMASK) {
                  const bool o = mr[joff[m]] != 0;
                  ll1 = o ? ll1 : 0.0f;
                  dz[h] = o ? dz[h] : 0.0f;
                }
                ll_lane = fmaf(wv[m], ll1, ll_lane);
              }
              dz2 = pack2(dz[0], dz[1]);
              t02 = pack2(0.0f, 0.0f);
            }
            if (GRAD) {
              if (MODEL == 1) {
                gth2[0] = add2(gth2[0], dz2);
                acc2[k][0] = add2(acc2[k][0], dz2);
              } else {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                  gth2[d] = fma2(dz2, a2[k][d], gth2[d]);        // d LL/d theta = -sum dz a
                  acc2[k][d] = fma2(dz2, th2[d], acc2[k][d]);    // +sum dz theta: negated when it is stored
                }
                acc2[k][D] = add2(acc2[k][D], dz2);
                if (MODEL == 3) acc2[k][D + 1] = add2(acc2[k][D + 1], t02);
              }
            }
          }
        }
        if (MODEL == 3 && ((rr & 3) == 3 || rr == NR - 1)) {
#pragma unroll
          for (int k = 0; k < MP; ++k) {
            float u0, u1, w0, w1;
            unpack2(prod2[k], u0, u1);
            unpack2(wl2k[k], w0, w1);
            ll_lane = fmaf(w0, lg2_approx(u0), fmaf(w1, lg2_approx(u1), ll_lane));
            prod2[k] = pack2(1.0f, 1.0f);
          }
        }
        if (GRAD) {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            float lo, hi;
            unpack2(gth2[MODEL == 1 ? 0 : d], lo, hi);
            part[d][rr] = lo + hi;
          }
        }
      }
      if (GRAD) {
        const int row = r0 + stream_row<NR>(lane);
#pragma unroll
        for (int k = 0; k < Q; ++k) {
          const float t = transpose_reduce<NR>(part[k], lane);
          if ((lane & (NR == 8 ? 3 : 7)) == 0 && (!RAGGED || row < rows))
            red[((size_t)row * NW + warp) * Q + k] = t;
        }
      }
    };
    int r0 = 0;
    if (has_missing) {
      for (; r0 + NR <= rows; r0 += NR) tile(r0, std::true_type{}, std::false_type{});
      if (r0 < rows) tile(r0, std::true_type{}, std::true_type{});
    } else {
      for (; r0 + NR <= rows; r0 += NR) tile(r0, std::false_type{}, std::false_type{});
      if (r0 < rows) tile(r0, std::false_type{}, std::true_type{});
    }
    ll_acc += (double)ll_lane;
    ll_lane = 0.0f;
    // look ahead: does the next chunk of this CTA hold a missing cell?  (its copy was
    // issued NS chunks ago, so this wait is normally free)
    {
      const int64_t cnext = c + gridDim.x;
      if (threadIdx.x == 0) sflag[par ^ 1] = (cnext < n_chunks && cnext >= n_full) ? 1 : 0;
      __syncwarp();
      if (cnext < n_full) {
        const int sn2 = s + 1 == NS ? 0 : s + 1;
        mbar_wait(&cx.bar[sn2], sn2 == 0 ? phase ^ 1u : phase);
      }
    }
    __syncthreads();   // every warp is done with stage s; red[buf] is complete; sflag[par^1] is reset
    {
      const int64_t cnext = c + gridDim.x;
      if (cnext < n_full) {
        const int sn2 = s + 1 == NS ? 0 : s + 1;
        if (scan_missing(cx.stages + (size_t)sn2 * p.stage_bytes, R)) sflag[par ^ 1] = 1;
      }
    }
    if (threadIdx.x == 0) {
      const int64_t cn = c + (int64_t)NS * gridDim.x;
      if (cn < n_full) stream_issue(p, cn, st, &cx.bar[s]);
    }
    if (GRAD) {
      for (int t = threadIdx.x; t < rows * D; t += blockDim.x) {
        const int r = t / D, d = t % D;
        float v = 0.0f;
        for (int w = 0; w < NW; ++w) v += red[((size_t)r * NW + w) * Q + d];
        g_ability[(c * R + r) * D + d] = v;
      }
    }
    __syncthreads();   // sflag[par^1] is final
    buf ^= 1;
    par ^= 1;
    if (++s == NS) {
      s = 0;
      phase ^= 1u;
    }
  }
  if (GRAD) {
    float* dst = part_gitem + (size_t)blockIdx.x * I * F;
#pragma unroll
    for (int k = 0; k < MP; ++k) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        float v[2];
        unpack2(acc2[k][f], v[0], v[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = 2 * k + h;
          // d ll/d a = -sum dz theta (the registers hold +sum); 3PL guess logit: d ll/d gamma = g (1 - g) sum t0
          const float sc = (MODEL == 3 && f == D + 1) ? gs[m] * omg[m] : ((MODEL != 1 && f < D) ? -1.0f : 1.0f);
          if (valid[m]) dst[(size_t)joff[m] * F + f] = v[h] * sc;
        }
      }
    }
  }
  // deterministic CTA sum of the log-likelihood
  __syncthreads();
  double* s_part = reinterpret_cast<double*>(cx.red);
  const double v = warp_sum(ll_acc);
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < NW; ++w) t += s_part[w];
    part_ll[blockIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------
// encode backward: A^r_j = sum_{i: o_ij, x_ij = r} GN_i,  B^r_j likewise with GS_i
//   GN = g_mu / S,  GS = -(g_mu mu + g_lv) / S      (SURVEY Appendix A "PoE")
// staged per-person arrays: 0 ability_mu, 1 S, 2 g_mu, 3 g_lv.
// part: cond  -> [grid][2][I][2D] (A | B);  uncond -> [grid][2][1][2D] (summed over items)
// ---------------------------------------------------------------------------
template <int D, int M>
__global__ void __launch_bounds__(512) encode_bwd_stream_kernel(const __grid_constant__ StreamParams p, int cond,
                                                                float* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char smem[];
  const StreamCtx cx = stream_setup(p, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  const int I = p.I, R = p.R, NS = p.NS;

  int joff[M];
  bool valid[M];
  float A[M][2][D], B[M][2][D];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int j = (warp * M + m) * 32 + lane;
    valid[m] = j < I;
    joff[m] = min(j, I - 1);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int d = 0; d < D; ++d) A[m][r][d] = B[m][r][d] = 0.0f;
  }
  const int64_t n_chunks = (p.P + R - 1) / R, n_full = p.P / R;
  int s = 0;
  uint32_t phase = 0;
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    unsigned char* st = cx.stages + (size_t)s * p.stage_bytes;
    const int rows = (int)((p.P - c * R < R) ? p.P - c * R : R);
    if (c < n_full) {
      mbar_wait(&cx.bar[s], phase);
    } else {
      stream_copy_ragged(p, c, st, rows);
      __syncthreads();
    }
    const float* sx = reinterpret_cast<const float*>(st);
    const uint8_t* sm = st + p.mask_off;
    const float* s_mu = reinterpret_cast<const float*>(st + p.parr_off);
    const float* s_S = s_mu + (size_t)R * D;
    const float* s_gm = s_S + (size_t)R * D;
    const float* s_gl = s_gm + (size_t)R * D;
    float* info = cx.info;   // [R][2D]: GN | GS
    for (int t = threadIdx.x; t < rows * D; t += blockDim.x) {
      const int r = t / D, d = t % D;
      const float sv = s_S[t], gm = s_gm[t];
      info[r * 2 * D + d] = gm / sv;
      info[r * 2 * D + D + d] = -(gm * s_mu[t] + s_gl[t]) / sv;
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
      float GN[D], GS[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        GN[d] = info[r * 2 * D + d];
        GS[d] = info[r * 2 * D + D + d];
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float x = sx[r * I + joff[m]];
        const bool o = valid[m] && sm[r * I + joff[m]] != 0;
        const float w1 = (o && x > 0.5f) ? 1.0f : 0.0f, w0 = (o && !(x > 0.5f)) ? 1.0f : 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          A[m][1][d] = fmaf(w1, GN[d], A[m][1][d]);
          B[m][1][d] = fmaf(w1, GS[d], B[m][1][d]);
          A[m][0][d] = fmaf(w0, GN[d], A[m][0][d]);
          B[m][0][d] = fmaf(w0, GS[d], B[m][0][d]);
        }
      }
    }
    __syncthreads();   // every warp is done with stage s and with info
    if (threadIdx.x == 0) {
      const int64_t cn = c + (int64_t)NS * gridDim.x;
      if (cn < n_full) stream_issue(p, cn, st, &cx.bar[s]);
    }
    if (++s == NS) {
      s = 0;
      phase ^= 1u;
    }
  }
  if (cond) {
    float* dst = part + (size_t)blockIdx.x * 2 * I * 2 * D;
#pragma unroll
    for (int m = 0; m < M; ++m)
      if (valid[m]) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int d = 0; d < D; ++d) {
            dst[(r * I + joff[m]) * 2 * D + d] = A[m][r][d];
            dst[(r * I + joff[m]) * 2 * D + D + d] = B[m][r][d];
          }
      }
  } else {
    // unconditional table: one entry per response value -> sum over items (fixed order)
    float* s_w = cx.red;   // [NW][4D]
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float va = 0.0f, vb = 0.0f;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          va += A[m][r][d];
          vb += B[m][r][d];
        }
        va = warp_sum(va);
        vb = warp_sum(vb);
        if (lane == 0) {
          s_w[warp * 4 * D + r * 2 * D + d] = va;
          s_w[warp * 4 * D + r * 2 * D + D + d] = vb;
        }
      }
    __syncthreads();
    if ((int)threadIdx.x < 4 * D) {
      float v = 0.0f;
      for (int w = 0; w < NW; ++w) v += s_w[w * 4 * D + threadIdx.x];
      part[(size_t)blockIdx.x * 4 * D + threadIdx.x] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// Conditional encode on the tensor cores.
//
// With a conditional posterior the product-of-experts sums are a matrix product
// with the 0/1 response matrix:  S_i[d] = sum_j tau0_jd + sum_j (o x)_ij (tau1 - tau0)_jd
// (N likewise), i.e. (rows x items) . (items x 2D).  The response indicator is
// exact in bf16; the table differences are split into three bf16 terms
// (hi + mid + lo = 24 significant bits, summed back in fp32), so the product
// carries fp32 accuracy:  columns = [S hi|mid|lo, N hi|mid|lo] = 6 D.
// mma.sync m16n8k16 (bf16 in, fp32 accumulate): this shape -- K = items up to
// 1024, N = 6 D <= 48, HBM-bound -- has no use for a tcgen05/TMEM pipeline.
//
// Work split: the 8 warps of a CTA split the items (K) of a 16-row tile:
// warp w owns k-steps [w*KS, (w+1)*KS) of 16 items each and keeps their B
// fragments in registers for the whole kernel; the partial 16 x 6D tiles are
// summed across warps through shared memory in a fixed order.  Missing cells
// (sparse) are corrected on the CUDA cores: S -= tau0_j, N -= mu0_j tau0_j,
// plus the prior expert's precision under VIBO_MISSING_PRIOR.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float bf16_round(float v) {
  uint32_t u = __float_as_uint(v);
  u += 0x7fffu + ((u >> 16) & 1u);   // round to nearest even on the upper 16 bits (finite inputs)
  return __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {   // both already bf16-representable
  return (__float_as_uint(lo) >> 16) | (__float_as_uint(hi) & 0xffff0000u);
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// expert entries of item j, dimension d: tau0, tau1 - tau0, mu0 tau0, mu1 tau1 - mu0 tau0
__device__ __forceinline__ void expert_entry(const float* __restrict__ table, int I, int D, int j, int d, float& t0,
                                             float& dt, float& n0, float& dn) {
  const float mu0 = table[((size_t)j) * 2 * D + d], lam0 = table[((size_t)j) * 2 * D + D + d];
  const float mu1 = table[((size_t)I + j) * 2 * D + d], lam1 = table[((size_t)I + j) * 2 * D + D + d];
  t0 = 1.0f / (expf(lam0) + kPoeEps);
  const float t1 = 1.0f / (expf(lam1) + kPoeEps);
  dt = t1 - t0;
  n0 = mu0 * t0;
  dn = mu1 * t1 - n0;
}

// A fragment of an edge tile: element e (bit0 column within the pair, bit1 row half g / g+8,
// bit2 column half +8) is the response where the cell exists and is observed, else 0.
struct MmaAFrag {
  uint32_t a0, a1, a2, a3, miss;
};
static __device__ __noinline__ MmaAFrag mma_a_guarded(const float* sx, const uint8_t* sm, int I, int rows, int ra, int rb,
                                               int j0) {
  float w[8];
  uint32_t miss = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int r = (e & 2) ? rb : ra;
    const int j = j0 + (e & 1) + ((e & 4) ? 8 : 0);
    const bool in = r < rows && j < I;
    float x = 0.0f;
    bool o = true;
    if (in) {
      x = sx[r * I + j];
      o = sm[r * I + j] != 0;
    }
    w[e] = (in && o) ? x : 0.0f;
    if (in && !o) miss |= 1u << e;
  }
  MmaAFrag f;
  f.a0 = pack_bf16x2(bf16_round(w[0]), bf16_round(w[1]));   // row g,   cols 2t, 2t+1
  f.a1 = pack_bf16x2(bf16_round(w[2]), bf16_round(w[3]));   // row g+8, cols 2t, 2t+1
  f.a2 = pack_bf16x2(bf16_round(w[4]), bf16_round(w[5]));   // row g,   cols 2t+8, 2t+9
  f.a3 = pack_bf16x2(bf16_round(w[6]), bf16_round(w[7]));   // row g+8, cols 2t+8, 2t+9
  f.miss = miss;
  return f;
}

// smem (after the barrier block): red [NW][16][NT*8] | corr [NW][16][2D+1] | base [2D] |
// tab [I][2D] (tau0 | mu0 tau0, for the missing-cell correction) | stages
constexpr int kMmaWarps = 8;
template <int D, int KS, bool EVEN>
__global__ void __launch_bounds__(kMmaWarps * 32) encode_mma_kernel(const __grid_constant__ StreamParams p, int missing_policy,
                                                         const float* __restrict__ table,
                                                         float* __restrict__ out_mu, float* __restrict__ out_lv,
                                                         float* __restrict__ out_S) {
  constexpr int NT = (6 * D + 7) / 8, NC = NT * 8, QC = 2 * D + 1;
  extern __shared__ __align__(128) unsigned char smem[];
  const StreamCtx cx = stream_setup(p, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int I = p.I, R = p.R, NS = p.NS;
  float* red = cx.red;                       // [NW][16][NC]
  float* corr = cx.info;                     // [NW][16][QC]
  float* base = corr + (size_t)NW * 16 * QC;   // [2D]
  float* tab = base + 2 * D;                   // [I][2D]
  int* sflag = reinterpret_cast<int*>(smem + 64);   // behind the (<= 8) mbarriers

  // ---- B fragments of this warp's k-steps (registers, whole kernel) ----------
  // b0: rows k = 2t, 2t+1 of the k-step, column n = g of the n-tile; b1: rows 2t+8, 2t+9.
  uint32_t bf[KS][NT][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int j0 = 16 * (warp * KS + ks) + 2 * t;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int col = 8 * nt + g;
      const int kind = col / (3 * D), sp = (col % (3 * D)) / D, d = col % D;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + (e & 1) + (e >> 1) * 8;
        float val = 0.0f;
        if (j < I && col < 6 * D) {
          float t0, dt, n0, dn;
          expert_entry(table, I, D, j, d, t0, dt, n0, dn);
          const float full = kind == 0 ? dt : dn;
          const float hi = bf16_round(full), mid = bf16_round(full - hi), lo = bf16_round(full - hi - mid);
          val = sp == 0 ? hi : (sp == 1 ? mid : lo);
        }
        v[e] = val;
      }
      bf[ks][nt][0] = pack_bf16x2(v[0], v[1]);
      bf[ks][nt][1] = pack_bf16x2(v[2], v[3]);
    }
  }
  for (int k = threadIdx.x; k < I * D; k += blockDim.x) {
    const int j = k / D, d = k % D;
    float t0, dt, n0, dn;
    expert_entry(table, I, D, j, d, t0, dt, n0, dn);
    tab[j * 2 * D + d] = t0;
    tab[j * 2 * D + D + d] = n0;
  }
  __syncthreads();
  // base sums over all items, fixed order (thread k < 2D)
  if ((int)threadIdx.x < 2 * D) {
    double acc = 0.0;   // the missing-cell corrections are subtracted from this sum: keep it tight
    for (int j = 0; j < I; ++j) acc += (double)tab[j * 2 * D + threadIdx.x];
    base[threadIdx.x] = (float)acc;
  }
  const float prior_tau = (missing_policy == VIBO_MISSING_PRIOR) ? 1.0f / (1.0f + kPoeEps) : 0.0f;
  __syncthreads();

  const int64_t n_chunks = (p.P + R - 1) / R, n_full = p.P / R;
  int s = 0;
  uint32_t phase = 0;
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    unsigned char* st = cx.stages + (size_t)s * p.stage_bytes;
    const int rows = (int)((p.P - c * R < R) ? p.P - c * R : R);
    if (threadIdx.x == 0) *sflag = 0;
    if (c < n_full) {
      mbar_wait(&cx.bar[s], phase);
    } else {
      stream_copy_ragged(p, c, st, rows);
    }
    __syncthreads();
    const float* sx = reinterpret_cast<const float*>(st);
    const uint8_t* sm = st + p.mask_off;
    // any zero byte in the stage's mask block?  (16 bytes per load)
    {
      bool miss = false;
      const uint4* m4 = reinterpret_cast<const uint4*>(sm);
      const int n16 = (rows * I) >> 4;
      for (int k = threadIdx.x; k < n16; k += blockDim.x) {
        const uint4 w = m4[k];
        const uint32_t z = ((w.x - 0x01010101u) & ~w.x) | ((w.y - 0x01010101u) & ~w.y) |
                           ((w.z - 0x01010101u) & ~w.z) | ((w.w - 0x01010101u) & ~w.w);
        miss = miss || (z & 0x80808080u) != 0;
      }
      for (int k = (n16 << 4) + threadIdx.x; k < rows * I; k += blockDim.x) miss = miss || sm[k] == 0;
      if (miss) *sflag = 1;
    }
    __syncthreads();
    const bool has_missing = *sflag != 0;
    for (int r0 = 0; r0 < rows; r0 += 16) {
      float C[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) C[nt][0] = C[nt][1] = C[nt][2] = C[nt][3] = 0.0f;
      float cs[2][D], cn[2][D], cnt[2] = {0.0f, 0.0f};   // missing-cell corrections of rows g, g+8
#pragma unroll
      for (int d = 0; d < D; ++d) cs[0][d] = cs[1][d] = cn[0][d] = cn[1][d] = 0.0f;
      const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int jbase = 16 * (warp * KS + ks);
        if (jbase < I) {   // warp-uniform
          uint32_t a[4];
          bool miss[8];
          bool any_miss = false;
          if (EVEN && r0 + 16 <= rows && jbase + 16 <= I) {
            // interior tile: unguarded 64-bit loads; a 0/1 float's upper half IS its bf16
            const float* xa = sx + ra * I + jbase + 2 * t;
            const float2 x0 = *reinterpret_cast<const float2*>(xa);
            const float2 x1 = *reinterpret_cast<const float2*>(xa + 8 * I);
            const float2 x2 = *reinterpret_cast<const float2*>(xa + 8);
            const float2 x3 = *reinterpret_cast<const float2*>(xa + 8 * I + 8);
            a[0] = __byte_perm(__float_as_uint(x0.x), __float_as_uint(x0.y), 0x7632);
            a[1] = __byte_perm(__float_as_uint(x1.x), __float_as_uint(x1.y), 0x7632);
            a[2] = __byte_perm(__float_as_uint(x2.x), __float_as_uint(x2.y), 0x7632);
            a[3] = __byte_perm(__float_as_uint(x3.x), __float_as_uint(x3.y), 0x7632);
            uint32_t mm[4] = {0x0101u, 0x0101u, 0x0101u, 0x0101u};
            if (has_missing) {
              const uint8_t* ma = sm + ra * I + jbase + 2 * t;
              mm[0] = *reinterpret_cast<const uint16_t*>(ma);
              mm[1] = *reinterpret_cast<const uint16_t*>(ma + 8 * I);
              mm[2] = *reinterpret_cast<const uint16_t*>(ma + 8);
              mm[3] = *reinterpret_cast<const uint16_t*>(ma + 8 * I + 8);
            }
            // all eight mask bytes non-zero?
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (!has_missing) {
                miss[((k & 1) ? 2 : 0) + ((k & 2) ? 4 : 0)] = miss[((k & 1) ? 2 : 0) + ((k & 2) ? 4 : 0) + 1] = false;
                continue;
              }
              const bool lo_miss = (mm[k] & 0xffu) == 0, hi_miss = (mm[k] & 0xff00u) == 0;
              // e index: bit0 column within the pair, bit1 row half, bit2 column half
              const int e0 = ((k & 1) ? 2 : 0) + ((k & 2) ? 4 : 0);
              miss[e0] = lo_miss;
              miss[e0 + 1] = hi_miss;
              any_miss = any_miss || lo_miss || hi_miss;
              if (lo_miss) a[k] &= 0xffff0000u;   // a missing cell holds -1: drop it from the product
              if (hi_miss) a[k] &= 0x0000ffffu;
            }
          } else {
            // edge tile (ragged rows / items, odd I): guarded scalar loads, kept out of line
            const MmaAFrag fr = mma_a_guarded(sx, sm, I, rows, ra, rb, jbase + 2 * t);
            a[0] = fr.a0; a[1] = fr.a1; a[2] = fr.a2; a[3] = fr.a3;
#pragma unroll
            for (int e = 0; e < 8; ++e) miss[e] = (fr.miss >> e) & 1u;
            any_miss = fr.miss != 0;
          }
          if (any_miss) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (miss[e]) {
                const int j = jbase + 2 * t + (e & 1) + ((e & 4) ? 8 : 0);
                const int h = (e & 2) ? 1 : 0;
                cnt[h] += 1.0f;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                  cs[h][d] += tab[j * 2 * D + d];
                  cn[h][d] += tab[j * 2 * D + D + d];
                }
              }
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(C[nt], a, bf[ks][nt]);
        }
      }
      // partial tile -> shared memory;  C: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
      float* rw = red + (size_t)warp * 16 * NC;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        *reinterpret_cast<float2*>(&rw[g * NC + 8 * nt + 2 * t]) = make_float2(C[nt][0], C[nt][1]);
        *reinterpret_cast<float2*>(&rw[(g + 8) * NC + 8 * nt + 2 * t]) = make_float2(C[nt][2], C[nt][3]);
      }
      float* cw = corr + (size_t)warp * 16 * QC;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!has_missing) break;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float v = cs[h][d], u = cn[h][d];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          u += __shfl_xor_sync(0xffffffffu, u, 1);
          u += __shfl_xor_sync(0xffffffffu, u, 2);
          if (t == 0) {
            cw[(g + 8 * h) * QC + d] = v;
            cw[(g + 8 * h) * QC + D + d] = u;
          }
        }
        float k = cnt[h];
        k += __shfl_xor_sync(0xffffffffu, k, 1);
        k += __shfl_xor_sync(0xffffffffu, k, 2);
        if (t == 0) cw[(g + 8 * h) * QC + 2 * D] = k;
      }
      __syncthreads();
      const int tile_rows = min(16, rows - r0);
      for (int q = threadIdx.x; q < tile_rows * D; q += blockDim.x) {
        const int r = q / D, d = q % D;
        float sv = 0.0f, nv = 0.0f, cS = 0.0f, cN = 0.0f, cm = 0.0f;
        for (int w2 = 0; w2 < NW; ++w2) {
          const float* rr = red + ((size_t)w2 * 16 + r) * NC;
          sv += (rr[d] + rr[D + d]) + rr[2 * D + d];
          nv += (rr[3 * D + d] + rr[4 * D + d]) + rr[5 * D + d];
          if (has_missing) {
            const float* cc = corr + ((size_t)w2 * 16 + r) * QC;
            cS += cc[d];
            cN += cc[D + d];
            cm += cc[2 * D];
          }
        }
        sv = (base[d] - cS) + sv + cm * prior_tau;
        nv = (base[D + d] - cN) + nv;
        if (has_missing && cm >= (float)I) {
          // no observed cell: base - correction is a rounding residue, not zero.  Exact values
          // instead: prior experts only, or (--drop-missing) an empty product -> 0 / 0 = NaN,
          // as in the reference (models.py:614-621).
          sv = cm * prior_tau;
          nv = 0.0f;
        }
        const int64_t row = c * R + r0 + r;
        out_mu[row * D + d] = nv / sv;
        out_lv[row * D + d] = logf(1.0f / sv);
        if (out_S) out_S[row * D + d] = sv;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const int64_t cn2 = c + (int64_t)NS * gridDim.x;
      if (cn2 < n_full) stream_issue(p, cn2, st, &cx.bar[s]);
    }
    if (++s == NS) {
      s = 0;
      phase ^= 1u;
    }
  }
}

// ---------------------------------------------------------------------------
// Encode backward on the tensor cores (conditional table).
//
//   A^r_j[d] = sum_i [o_ij, x_ij = r] GN_i[d],   B^r_j[d] likewise with GS_i
// is (items x rows) . (rows x 2D) with a 0/1 left operand.  TF32 mma.sync
// m16n8k8 holds ONE element per fragment register, so the transposed operand
// W^T[item][row] is read straight from the row-major stage (no transposition
// pass); the right operand [GN | GS] is split into two TF32 terms (22
// significant bits) -> columns = [hi(2D) | lo(2D)].  Two accumulator sets: the
// x = 1 indicator and the observed indicator (A^0 = A^obs - A^1); for a stage
// without missing cells the observed product does not depend on the item and
// collapses to a column sum of G kept by 4D threads.
// Warp w owns items [w*16*MT, (w+1)*16*MT) and keeps their accumulators in
// registers for the whole kernel (deterministic).
// smem: red = G tile [R][NC] | info: flags, column sums | stages
// ---------------------------------------------------------------------------
__device__ __forceinline__ float tf32_round(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32_1688(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int D, int MT>
__global__ void __launch_bounds__(512) encode_bwd_mma_kernel(const __grid_constant__ StreamParams p,
                                                             float* __restrict__ part) {
  constexpr int NT = (4 * D + 7) / 8, NC = NT * 8;
  extern __shared__ __align__(128) unsigned char smem[];
  const StreamCtx cx = stream_setup(p, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int I = p.I, R = p.R, NS = p.NS;
  float* Gs0 = cx.red;                                  // [2][R][NC] tf32 bit patterns (double-buffered)
  int* wflag = reinterpret_cast<int*>(cx.info);         // [2][16]: per-warp "saw a missing cell"
  const int NWB = blockDim.x >> 5;
  float T1[MT][NT][4], To[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int k = 0; k < 4; ++k) T1[mt][nt][k] = To[mt][nt][k] = 0.0f;
  float gfull = 0.0f;   // thread c < NC: sum over fully observed stages of G[.][c]
  const int j0 = warp * 16 * MT;

  const int64_t n_chunks = (p.P + R - 1) / R, n_full = p.P / R;
  // Stage prep (G tile as two TF32 terms + missing-cell flags) for the stage holding chunk c_;
  // runs one stage AHEAD of the MMAs, so each stage costs a single CTA-wide barrier.
  auto prep = [&](int64_t c_, int s_, uint32_t ph_, int buf_) {
    unsigned char* st_ = cx.stages + (size_t)s_ * p.stage_bytes;
    const int rows_ = (int)((p.P - c_ * R < R) ? p.P - c_ * R : R);
    if (c_ < n_full) {
      mbar_wait(&cx.bar[s_], ph_);
    } else {
      stream_copy_ragged(p, c_, st_, rows_);
      __syncthreads();
    }
    const uint8_t* sm_ = st_ + p.mask_off;
    const float* s_mu = reinterpret_cast<const float*>(st_ + p.parr_off);
    const float* s_S = s_mu + (size_t)R * D;
    const float* s_gm = s_S + (size_t)R * D;
    const float* s_gl = s_gm + (size_t)R * D;
    float* G = Gs0 + (size_t)buf_ * R * NC;
    for (int q = threadIdx.x; q < R * NC; q += blockDim.x) {
      const int r = q / NC, col = q % NC;
      float v = 0.0f;
      if (r < rows_ && col < 4 * D) {
        const int sp = col / (2 * D), k = col % (2 * D), d = k % D;
        const float sv = s_S[r * D + d], gm = s_gm[r * D + d];
        const float full = k < D ? gm / sv : -(gm * s_mu[r * D + d] + s_gl[r * D + d]) / sv;
        const float hi = tf32_round(full);
        v = sp == 0 ? hi : tf32_round(full - hi);
      }
      G[q] = v;
    }
    // any zero byte in the stage's mask block?  (16 bytes per load; every warp reports)
    bool miss = false;
    const uint4* m4 = reinterpret_cast<const uint4*>(sm_);
    const int n16 = (rows_ * I) >> 4;
    for (int k = threadIdx.x; k < n16; k += blockDim.x) {
      const uint4 w = m4[k];
      const uint32_t z = ((w.x - 0x01010101u) & ~w.x) | ((w.y - 0x01010101u) & ~w.y) |
                         ((w.z - 0x01010101u) & ~w.z) | ((w.w - 0x01010101u) & ~w.w);
      miss = miss || (z & 0x80808080u) != 0;
    }
    for (int k = (n16 << 4) + threadIdx.x; k < rows_ * I; k += blockDim.x) miss = miss || sm_[k] == 0;
    miss = __any_sync(0xffffffffu, miss);
    if (lane == 0) wflag[buf_ * 16 + warp] = miss ? 1 : 0;
  };
  prep(blockIdx.x, 0, 0, 0);
  __syncthreads();

  int s = 0, buf = 0;
  uint32_t phase = 0;
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    unsigned char* st = cx.stages + (size_t)s * p.stage_bytes;
    const int rows = (int)((p.P - c * R < R) ? p.P - c * R : R);
    const float* sx = reinterpret_cast<const float*>(st);
    const uint8_t* sm = st + p.mask_off;
    const float* Gs = Gs0 + (size_t)buf * R * NC;
    bool has_missing = false;
    for (int w2 = 0; w2 < NWB; ++w2) has_missing = has_missing || wflag[buf * 16 + w2] != 0;
    if (!has_missing && (int)threadIdx.x < NC) {
      float acc = 0.0f;
      for (int r = 0; r < rows; ++r) acc += Gs[r * NC + threadIdx.x];
      gfull += acc;
    }
    if (j0 < I) {
      for (int k0 = 0; k0 < rows; k0 += 8) {
        uint32_t bfr[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          bfr[nt][0] = __float_as_uint(Gs[(k0 + t) * NC + 8 * nt + g]);       // rows >= `rows` hold zeros
          bfr[nt][1] = __float_as_uint(Gs[(k0 + t + 4) * NC + 8 * nt + g]);
        }
        if (!has_missing && k0 + 8 <= rows && j0 + 16 * MT <= I) {
          // interior, fully observed: a 0/1 response IS its TF32 operand
          const float* xb = sx + (k0 + t) * I + j0 + g;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint32_t a1[4];
            a1[0] = __float_as_uint(xb[16 * mt]);
            a1[1] = __float_as_uint(xb[16 * mt + 8]);
            a1[2] = __float_as_uint(xb[4 * I + 16 * mt]);
            a1[3] = __float_as_uint(xb[4 * I + 16 * mt + 8]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_tf32_1688(T1[mt][nt], a1, bfr[nt]);
          }
          continue;
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          // a0: (item g, row t)  a1: (item g+8, row t)  a2: (item g, row t+4)  a3: (item g+8, row t+4)
          uint32_t a1[4], ao[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = j0 + 16 * mt + g + ((e & 1) ? 8 : 0);
            const int r = k0 + t + ((e & 2) ? 4 : 0);
            const bool in = j < I && r < rows;
            const int jj = in ? j : 0, rr = in ? r : 0;
            const float x = sx[rr * I + jj];
            const bool o = in && (has_missing ? sm[rr * I + jj] != 0 : true);
            a1[e] = (o && x > 0.5f) ? 0x3f800000u : 0u;
            ao[e] = o ? 0x3f800000u : 0u;
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_tf32_1688(T1[mt][nt], a1, bfr[nt]);
          if (has_missing) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_tf32_1688(To[mt][nt], ao, bfr[nt]);
          }
        }
      }
    }
    // prepare the NEXT stage of this CTA while slower warps finish their MMAs
    const int64_t cnext = c + gridDim.x;
    const int sn2 = s + 1 == NS ? 0 : s + 1;
    if (cnext < n_chunks) prep(cnext, sn2, sn2 == 0 ? phase ^ 1u : phase, buf ^ 1);
    __syncthreads();   // every warp is done with stage s and Gs[buf]; Gs[buf^1] / flags are ready
    if (threadIdx.x == 0) {
      const int64_t cn = c + (int64_t)NS * gridDim.x;
      if (cn < n_full) stream_issue(p, cn, st, &cx.bar[s]);
    }
    buf ^= 1;
    if (++s == NS) {
      s = 0;
      phase ^= 1u;
    }
  }
  // ---- epilogue: per-item sums -> global partials -----------------------------
  // stage memory is free: tile [I_pad][NC], one accumulator set at a time; the two
  // TF32 terms are added and A^0 = A^obs - A^1.
  __syncthreads();
  float* tile = reinterpret_cast<float*>(cx.stages);
  float* gf = Gs0;   // [NC]
  if ((int)threadIdx.x < NC) gf[threadIdx.x] = gfull;
  float* dst = part + (size_t)blockIdx.x * 2 * I * 2 * D;
#pragma unroll
  for (int set = 0; set < 2; ++set) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        // c0 (item g, col 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
        const int ja = j0 + 16 * mt + g, jb = ja + 8, col = 8 * nt + 2 * t;
        const float (&T)[4] = set == 0 ? T1[mt][nt] : To[mt][nt];
        tile[(size_t)ja * NC + col] = T[0];
        tile[(size_t)ja * NC + col + 1] = T[1];
        tile[(size_t)jb * NC + col] = T[2];
        tile[(size_t)jb * NC + col + 1] = T[3];
      }
    __syncthreads();
    for (int q = threadIdx.x; q < I * 2 * D; q += blockDim.x) {
      const int j = q / (2 * D), k = q % (2 * D);
      const float v = tile[(size_t)j * NC + k] + tile[(size_t)j * NC + 2 * D + k];
      if (set == 0) {
        dst[((size_t)1 * I + j) * 2 * D + k] = v;
      } else {
        const float vo = v + (gf[k] + gf[2 * D + k]);
        dst[((size_t)0 * I + j) * 2 * D + k] = vo - dst[((size_t)1 * I + j) * 2 * D + k];
      }
    }
    __syncthreads();
  }
}

// Defines stream_link_run<MODEL_> in its own translation unit (compile time).
#define VIBO_STREAM_LINK_INSTANTIATE(MODEL_, NAME_)                                                        \
  template <int D, int M>                                                                                  \
  static cudaError_t NAME_##_dm(const StreamPlan& pl, const StreamParams& p, const float* item_feat,       \
                                double* part_ll, float* g_ability, float* part_g, bool grad,               \
                                cudaStream_t st) {                                                         \
    constexpr int NR = D <= 4 ? 8 : 4;                                                                     \
    cudaError_t e;                                                                                         \
    if (grad) {                                                                                            \
      auto k = link_stream_kernel<MODEL_, D, M, NR, true>;                                                 \
      e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);              \
      if (e != cudaSuccess) return e;                                                                      \
      k<<<pl.grid, pl.NW * 32, pl.smem, st>>>(p, item_feat, part_ll, g_ability, part_g);                   \
    } else {                                                                                               \
      auto k = link_stream_kernel<MODEL_, D, M, NR, false>;                                                \
      e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);              \
      if (e != cudaSuccess) return e;                                                                      \
      k<<<pl.grid, pl.NW * 32, pl.smem, st>>>(p, item_feat, part_ll, nullptr, nullptr);                    \
    }                                                                                                      \
    return cudaGetLastError();                                                                             \
  }                                                                                                        \
  template <int D>                                                                                         \
  static cudaError_t NAME_##_d(const StreamPlan& pl, const StreamParams& p, const float* item_feat,        \
                               double* part_ll, float* g_ability, float* part_g, bool grad,                \
                               cudaStream_t st) {                                                          \
    switch (pl.M) {                                                                                        \
      case 1: return NAME_##_dm<D, 1>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);             \
      case 2: return NAME_##_dm<D, 2>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);             \
      case 4: return NAME_##_dm<D, 4>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);             \
      default: return cudaErrorInvalidValue;                                                               \
    }                                                                                                      \
  }                                                                                                        \
  cudaError_t NAME_(const StreamPlan& pl, const StreamParams& p, int D, const float* item_feat,            \
                    double* part_ll, float* g_ability, float* part_g, bool grad, cudaStream_t st) {        \
    switch (D) {                                                                                           \
      case 1: return NAME_##_d<1>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);                 \
      case 2: return NAME_##_d<2>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);                 \
      case 3: return NAME_##_d<3>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);                 \
      case 4: return NAME_##_d<4>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);                 \
      case 5: return NAME_##_d<5>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);                 \
      case 6: return NAME_##_d<6>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);                 \
      case 7: return NAME_##_d<7>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);                 \
      case 8: return NAME_##_d<8>(pl, p, item_feat, part_ll, g_ability, part_g, grad, st);                 \
      default: return cudaErrorInvalidValue;                                                               \
    }                                                                                                      \
  }

}  // namespace vibo
