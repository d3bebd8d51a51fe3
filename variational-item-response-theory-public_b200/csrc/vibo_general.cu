// General (decomposed) kernels of the VIBO ELBO path: encode, encode-backward,
// link + log-likelihood, decode, Bernoulli log-likelihood on a materialised
// response_mu.  They cover every variant (1/2/3PL, D <= 8, conditional or not,
// missing data with prior experts or dropped) and back the module API pieces
// (encode / decode / elbo / flows).  The single-pass fused kernel for the
// headline configuration lives in vibo_fused.cu.
//
// Layout ("row-slab"): a CTA has NS warps; warp w owns the item slab
// [w*32*M, (w+1)*32*M) -- lane l owns items w*32*M + m*32 + l, m < M -- for
// EVERY row the CTA processes, so per-item quantities (item parameters, expert
// table entries, per-item gradient accumulators) live in registers for the
// whole kernel and cross-person sums need no atomics.  Per-row quantities are
// combined across the NS slabs through shared memory in a fixed order
// (deterministic).  CTAs are persistent over row tiles; per-CTA partials go to
// the workspace and a small second kernel reduces them in a fixed order.
#include <cstdlib>

#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

constexpr int kRowsPerTile = 8;

__device__ __forceinline__ void block_sum_to(double v, double* dst) {
  __shared__ double s_part[32];
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
    *dst = t;
  }
}

// ---------------------------------------------------------------------------
// encode: product-of-experts sums (models.py:596-629, utils.py:105-113)
// ---------------------------------------------------------------------------
template <int D, int M>
__global__ void __launch_bounds__(512) encode_kernel(int64_t P, int I, int cond, int missing_policy,
                                                       const float* __restrict__ resp,
                                                       const uint8_t* __restrict__ mask,
                                                       const float* __restrict__ table,
                                                       float* __restrict__ out_mu,
                                                       float* __restrict__ out_lv,
                                                       float* __restrict__ out_S) {
  extern __shared__ float red[];  // [kRowsPerTile][NS][2D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NS = blockDim.x >> 5;
  int j[M];
  float tau[M][2][D], mt[M][2][D];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    j[m] = (warp * M + m) * 32 + lane;
    const int jt = cond ? min(j[m], I - 1) : 0;
    const int It = cond ? I : 1;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float mu = table[((size_t)r * It + jt) * 2 * D + d];
        const float lam = table[((size_t)r * It + jt) * 2 * D + D + d];
        const float t = 1.0f / (expf(lam) + kPoeEps);
        tau[m][r][d] = t;
        mt[m][r][d] = mu * t;
      }
  }
  const float prior_tau = (missing_policy == 0) ? 1.0f / (1.0f + kPoeEps) : 0.0f;
  const int64_t n_tiles = (P + kRowsPerTile - 1) / kRowsPerTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 2
    for (int rt = 0; rt < kRowsPerTile; ++rt) {
      const int64_t row = tile * kRowsPerTile + rt;
      if (row >= P) break;
      float S[D], N[D];
#pragma unroll
      for (int d = 0; d < D; ++d) S[d] = N[d] = 0.0f;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        if (j[m] < I) {
          const float x = resp[row * I + j[m]];
          const bool o = mask[row * I + j[m]] != 0;
          const bool x1 = x > 0.5f;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            S[d] += o ? (x1 ? tau[m][1][d] : tau[m][0][d]) : prior_tau;
            N[d] += o ? (x1 ? mt[m][1][d] : mt[m][0][d]) : 0.0f;
          }
        }
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float s = warp_sum(S[d]), n = warp_sum(N[d]);
        if (lane == 0) {
          red[(rt * NS + warp) * 2 * D + d] = s;
          red[(rt * NS + warp) * 2 * D + D + d] = n;
        }
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kRowsPerTile * D; t += blockDim.x) {
      const int rt = t / D, d = t % D;
      const int64_t row = tile * kRowsPerTile + rt;
      if (row < P) {
        float s = 0.0f, n = 0.0f;
        for (int w = 0; w < NS; ++w) {
          s += red[(rt * NS + w) * 2 * D + d];
          n += red[(rt * NS + w) * 2 * D + D + d];
        }
        out_mu[row * D + d] = n / s;
        out_lv[row * D + d] = logf(1.0f / s);
        if (out_S) out_S[row * D + d] = s;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// encode backward: scatter per-person (GN, GS) into per-(r, item) sums
// ---------------------------------------------------------------------------
template <int D, int M>
__global__ void __launch_bounds__(512) encode_bwd_kernel(int64_t P, int I, const float* __restrict__ resp,
                                                           const uint8_t* __restrict__ mask,
                                                           const float* __restrict__ amu,
                                                           const float* __restrict__ Ssum,
                                                           const float* __restrict__ g_mu,
                                                           const float* __restrict__ g_lv,
                                                           float* __restrict__ part /*[grid][2][I][2D] (A|B)*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int j[M];
  float A[M][2][D], B[M][2][D];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    j[m] = (warp * M + m) * 32 + lane;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int d = 0; d < D; ++d) A[m][r][d] = B[m][r][d] = 0.0f;
  }
  const int64_t n_tiles = (P + kRowsPerTile - 1) / kRowsPerTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 2
    for (int rt = 0; rt < kRowsPerTile; ++rt) {
      const int64_t row = tile * kRowsPerTile + rt;
      if (row >= P) break;
      float GN[D], GS[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float s = Ssum[row * D + d];
        const float gm = g_mu[row * D + d];
        GN[d] = gm / s;
        GS[d] = -(gm * amu[row * D + d] + g_lv[row * D + d]) / s;
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        if (j[m] < I) {
          const float x = resp[row * I + j[m]];
          const bool o = mask[row * I + j[m]] != 0;
          const bool x1 = x > 0.5f;
          if (o) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              if (x1) {
                A[m][1][d] += GN[d];
                B[m][1][d] += GS[d];
              } else {
                A[m][0][d] += GN[d];
                B[m][0][d] += GS[d];
              }
            }
          }
        }
      }
    }
  }
  float* dst = part + (size_t)blockIdx.x * 2 * I * 2 * D;
#pragma unroll
  for (int m = 0; m < M; ++m)
    if (j[m] < I) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int d = 0; d < D; ++d) {
          dst[((size_t)r * I + j[m]) * 2 * D + d] = A[m][r][d];
          dst[((size_t)r * I + j[m]) * 2 * D + D + d] = B[m][r][d];
        }
    }
}

// Reduce the per-CTA (A, B) partials in a fixed order and apply the expert
// chain rule: d/d mu = tau A ; d/d lam = (mu A + B) (-exp(lam) tau^2).
// Unconditional tables additionally sum over items.
constexpr int kSumSlices = 32;   // partial slices per block of the fixed-order reducers (short dependent chains)
__global__ void __launch_bounds__(32 * kSumSlices) encode_bwd_finalize_kernel(int I, int D, int cond, int nparts,
                                                                              const float* __restrict__ part,
                                                                              const float* __restrict__ table,
                                                                              float* __restrict__ g_table) {
  // block = 32 outputs x 32 slices of the partials, combined through shared memory in a fixed order
  __shared__ double sa[kSumSlices][33], sb[kSumSlices][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int It = cond ? I : 1;
  const int n = 2 * It * D;
  const int t = blockIdx.x * 32 + lane;
  double a = 0.0, b = 0.0;
  int d = 0, jt = 0, r = 0;
  if (t < n) {
    d = t % D;
    jt = (t / D) % It;
    r = t / (D * It);
    const int j0 = cond ? jt : 0, j1 = cond ? jt + 1 : I;
    for (int p = slice; p < nparts; p += kSumSlices) {
      const float* src = part + (size_t)p * 2 * I * 2 * D;
      for (int j = j0; j < j1; ++j) {
        a += src[((size_t)r * I + j) * 2 * D + d];
        b += src[((size_t)r * I + j) * 2 * D + D + d];
      }
    }
  }
  sa[slice][lane] = a;
  sb[slice][lane] = b;
  __syncthreads();
  if (slice == 0 && t < n) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int q = 0; q < kSumSlices; ++q) {
      ta += sa[q][lane];
      tb += sb[q][lane];
    }
    const float mu = table[((size_t)r * It + jt) * 2 * D + d];
    const float lam = table[((size_t)r * It + jt) * 2 * D + D + d];
    const float el = expf(lam);
    const float tau = 1.0f / (el + kPoeEps);
    g_table[((size_t)r * It + jt) * 2 * D + d] = tau * (float)ta;
    g_table[((size_t)r * It + jt) * 2 * D + D + d] = (mu * (float)ta + (float)tb) * (-el * tau * tau);
  }
}

// Unconditional encode for NARROW rows (I <= 256), one warp per row with direct coalesced loads.  The
// unconditional posterior needs a row only through (observed ones, missing cells); for rows of a few hundred
// bytes the slab-stream kernel's per-chunk chain (bulk copy -> wait -> reduce -> barrier) caps it near
// 2.3 TB/s, while plain loads with four rows in flight per warp and 64 warps per SM are bound by DRAM latency
// alone.  Same posterior arithmetic as encode_stream_kernel<.., COND = false>; no alignment requirement.
// out_counts (P, 2) = (observed ones, observed cells) and / or the posterior (out_mu == nullptr: counts only).
template <int D>
__global__ void __launch_bounds__(256) encode_rows_uncond_kernel(int64_t P, int I, int missing_policy,
                                                                 const float* __restrict__ resp,
                                                                 const uint8_t* __restrict__ mask,
                                                                 const float* __restrict__ table,
                                                                 float* __restrict__ out_mu,
                                                                 float* __restrict__ out_lv,
                                                                 float* __restrict__ out_S,
                                                                 float* __restrict__ out_counts) {
  constexpr int NRW = 4;   // rows per warp and trip
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float prior_tau = (missing_policy == VIBO_MISSING_PRIOR) ? 1.0f / (1.0f + kPoeEps) : 0.0f;
  for (int64_t row0 = gwarp * NRW; row0 < P; row0 += nwarps * NRW) {
    const float* xr[NRW];
    const uint8_t* mr[NRW];
#pragma unroll
    for (int q = 0; q < NRW; ++q) {
      const int64_t row = row0 + q < P ? row0 + q : P - 1;   // clamped: the result of a row past the end is dropped
      xr[q] = resp + row * I;
      mr[q] = mask + row * I;
    }
    uint32_t cnt[NRW];   // ones | missing << 16  (I <= 2048)
#pragma unroll
    for (int q = 0; q < NRW; ++q) cnt[q] = 0;
#pragma unroll 2
    for (int j = lane; j < I; j += 32) {
      float x[NRW];
      uint8_t o[NRW];
#pragma unroll
      for (int q = 0; q < NRW; ++q) {
        x[q] = xr[q][j];
        o[q] = mr[q][j];
      }
#pragma unroll
      for (int q = 0; q < NRW; ++q) cnt[q] += o[q] != 0 ? (x[q] > 0.5f ? 1u : 0u) : 65536u;
    }
#pragma unroll
    for (int q = 0; q < NRW; ++q) cnt[q] = __reduce_add_sync(0xffffffffu, cnt[q]);
    // lane q finishes row row0 + q
    uint32_t c = cnt[0];
#pragma unroll
    for (int q = 1; q < NRW; ++q) c = lane == q ? cnt[q] : c;
    const int64_t row = row0 + lane;
    if (lane < NRW && row < P) {
      const float n1 = (float)(c & 0xffffu), nm = (float)(c >> 16);
      if (out_counts != nullptr) {
        out_counts[row * 2] = n1;
        out_counts[row * 2 + 1] = (float)I - nm;
      }
      if (out_mu != nullptr) {
        const float nz = (float)I - nm - n1;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float mu0 = table[d], mu1 = table[2 * D + d];
          const float ta0 = 1.0f / (expf(table[D + d]) + kPoeEps), ta1 = 1.0f / (expf(table[3 * D + d]) + kPoeEps);
          const float sv = fmaf(nz, ta0, fmaf(n1, ta1, nm * prior_tau));
          const float nv = fmaf(nz, mu0 * ta0, n1 * (mu1 * ta1));
          out_mu[row * D + d] = nv / sv;
          out_lv[row * D + d] = logf(1.0f / sv);
          if (out_S) out_S[row * D + d] = sv;
        }
      }
    }
  }
}

// Unconditional encode backward from the forward pass's per-person counts.  The unconditional table has ONE
// entry per response value, so A^r = sum_i n^r_i GN_i and B^r = sum_i n^r_i GS_i (n^1 = observed ones,
// n^0 = observed - ones; missing cells carry no table entry under either policy) need 16 + 16 D bytes per
// person instead of a second pass over the 5 I bytes of its row.  part: [grid][2][1][2D], the layout
// encode_bwd_finalize_kernel reads for the unconditional table.  Fixed summation order.
template <int D>
__global__ void __launch_bounds__(256) encode_bwd_counts_kernel(int64_t P, const float* __restrict__ counts,
                                                                const float* __restrict__ amu,
                                                                const float* __restrict__ S,
                                                                const float* __restrict__ g_mu,
                                                                const float* __restrict__ g_lv,
                                                                float* __restrict__ part) {
  float acc[4 * D];   // [r][A | B][d]
#pragma unroll
  for (int q = 0; q < 4 * D; ++q) acc[q] = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 c = reinterpret_cast<const float2*>(counts)[i];
    const float n1 = c.x, n0 = c.y - c.x;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float sv = S[i * D + d], gm = g_mu[i * D + d];
      const float gn = gm / sv, gs = -(gm * amu[i * D + d] + g_lv[i * D + d]) / sv;
      acc[d] = fmaf(n0, gn, acc[d]);
      acc[D + d] = fmaf(n0, gs, acc[D + d]);
      acc[2 * D + d] = fmaf(n1, gn, acc[2 * D + d]);
      acc[3 * D + d] = fmaf(n1, gs, acc[3 * D + d]);
    }
  }
  __shared__ float s_w[8][4 * D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 4 * D; ++q) {
    const float v = warp_sum(acc[q]);
    if (lane == 0) s_w[warp][q] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < 4 * D) {
    float v = 0.0f;
    for (int w = 0; w < 8; ++w) v += s_w[w][threadIdx.x];
    part[(size_t)blockIdx.x * 4 * D + threadIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------
// link + log-likelihood (models.py:729-766, utils.py:46-49, models.py:399)
// ---------------------------------------------------------------------------
template <int D, int MODEL, int M, bool GRAD>
__global__ void __launch_bounds__(512) link_kernel(int64_t P, int I, const float* __restrict__ resp,
                                                     const uint8_t* __restrict__ mask,
                                                     const float* __restrict__ ability,
                                                     const float* __restrict__ item_feat,
                                                     double* __restrict__ part_ll,
                                                     float* __restrict__ g_ability,
                                                     float* __restrict__ part_gitem /*[grid][I*F]*/) {
  constexpr int F = item_width(MODEL, D);
  constexpr int DA = MODEL == 1 ? 1 : D;  // discrimination width
  extern __shared__ float red[];          // [kRowsPerTile][NS][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NS = blockDim.x >> 5;
  int j[M];
  float a[M][DA], b[M], gs[M], acc[M][F];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    j[m] = (warp * M + m) * 32 + lane;
    const int jj = min(j[m], I - 1);
    if constexpr (MODEL == 1) {
      b[m] = item_feat[jj];
      a[m][0] = 0.0f;
      gs[m] = 0.0f;
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) a[m][d] = item_feat[(size_t)jj * F + d];
      b[m] = item_feat[(size_t)jj * F + D];
      gs[m] = MODEL == 3 ? 1.0f / (1.0f + expf(-item_feat[(size_t)jj * F + D + 1])) : 0.0f;
    }
#pragma unroll
    for (int f = 0; f < F; ++f) acc[m][f] = 0.0f;
  }
  double ll_acc = 0.0;
  const int64_t n_tiles = (P + kRowsPerTile - 1) / kRowsPerTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 2
    for (int rt = 0; rt < kRowsPerTile; ++rt) {
      const int64_t row = tile * kRowsPerTile + rt;
      if (row >= P) break;
      float th[D], gth[D];
      float tsum = 0.0f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        th[d] = ability[row * D + d];
        tsum += th[d];
        gth[d] = 0.0f;
      }
      float ll_row = 0.0f;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        if (j[m] < I) {
          const float x = resp[row * I + j[m]];
          const bool o = mask[row * I + j[m]] != 0;
          if (o) {
            float z = b[m];
            if constexpr (MODEL == 1) {
              z += tsum;
            } else {
#pragma unroll
              for (int d = 0; d < D; ++d) z = fmaf(-th[d], a[m][d], z);
            }
            CellGrad c;
            if constexpr (MODEL == 3) c = cell_3pl<true>(z, gs[m], x > 0.5f);
            else c = cell_logistic<true>(z, x > 0.5f);
            ll_row += c.ll;
            if (GRAD) {
              if constexpr (MODEL == 1) {
                gth[0] += c.dz;
                acc[m][0] += c.dz;
              } else {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                  gth[d] = fmaf(-c.dz, a[m][d], gth[d]);
                  acc[m][d] = fmaf(-c.dz, th[d], acc[m][d]);
                }
                acc[m][D] += c.dz;
                if constexpr (MODEL == 3) acc[m][D + 1] += c.dgam;
              }
            }
          }
        }
      }
      ll_acc += (double)ll_row;
      if (GRAD) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float v = warp_sum(MODEL == 1 ? gth[0] : gth[d]);
          if (lane == 0) red[(rt * NS + warp) * D + d] = v;
        }
      }
    }
    if (GRAD) {
      __syncthreads();
      for (int t = threadIdx.x; t < kRowsPerTile * D; t += blockDim.x) {
        const int rt = t / D, d = t % D;
        const int64_t row = tile * kRowsPerTile + rt;
        if (row < P) {
          float v = 0.0f;
          for (int w = 0; w < NS; ++w) v += red[(rt * NS + w) * D + d];
          g_ability[row * D + d] = v;
        }
      }
      __syncthreads();
    }
  }
  if (GRAD) {
    float* dst = part_gitem + (size_t)blockIdx.x * I * F;
#pragma unroll
    for (int m = 0; m < M; ++m)
      if (j[m] < I) {
#pragma unroll
        for (int f = 0; f < F; ++f) dst[(size_t)j[m] * F + f] = acc[m][f];
      }
  }
  block_sum_to(ll_acc, part_ll + blockIdx.x);
}

// out[k] = scale * sum_p part[p][k], fixed order (deterministic).  Block = 32 outputs x 32
// slices: each thread sums every 32nd partial, the slices are combined through shared memory.
__global__ void __launch_bounds__(32 * kSumSlices) sum_partials_f32_kernel(const float* __restrict__ part, int nparts,
                                                                           int n, float scale, float* __restrict__ out) {
  __shared__ double sh[kSumSlices][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (k < n)
    for (int p = slice; p < nparts; p += kSumSlices) s += part[(size_t)p * n + k];
  sh[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && k < n) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < kSumSlices; ++q) t += sh[q][lane];
    out[k] = scale * (float)t;
  }
}

// out[c] = sum_p part[p * stride + c] for c < ncols: one block per column, strided partial
// sums then a shared-memory tree in a fixed order.
__global__ void __launch_bounds__(128) sum_partials_f64_kernel(const double* __restrict__ part, int nparts,
                                                               int stride, int ncols, double* __restrict__ out) {
  __shared__ double sh[128];
  const int c = blockIdx.x;
  double s = 0.0;
  if (c < ncols)
    for (int p = threadIdx.x; p < nparts; p += 128) s += part[(size_t)p * stride + c];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 64; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0 && c < ncols) out[c] = sh[0];
}

// ---------------------------------------------------------------------------
// decode: response_mu (P, I)  (models.py:729-766)
// ---------------------------------------------------------------------------
template <int MODEL>
__global__ void decode_kernel(int64_t P, int I, int D, const float* __restrict__ ability,
                              const float* __restrict__ item_feat, float* __restrict__ out) {
  const int F = item_width(MODEL, D);
  const int64_t n = P * (int64_t)I;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = c / I;
    const int j = (int)(c - row * I);
    float z;
    if (MODEL == 1) {
      z = item_feat[j];
      for (int d = 0; d < D; ++d) z += ability[row * D + d];
    } else {
      z = item_feat[(size_t)j * F + D];
      for (int d = 0; d < D; ++d) z = fmaf(-ability[row * D + d], item_feat[(size_t)j * F + d], z);
    }
    const float s = 1.0f / (1.0f + expf(-z));
    if (MODEL == 3) {
      const float g = 1.0f / (1.0f + expf(-item_feat[(size_t)j * F + D + 1]));
      out[c] = fmaf(1.0f - g, s, g);
    } else {
      out[c] = s;
    }
  }
}

// ---------------------------------------------------------------------------
// masked Bernoulli log-likelihood of a materialised response_mu (utils.py:46-49)
// ---------------------------------------------------------------------------
__global__ void bernoulli_ll_kernel(int64_t n, const float* __restrict__ resp,
                                    const uint8_t* __restrict__ mask, const float* __restrict__ prob,
                                    double* __restrict__ part_ll, float* __restrict__ g_prob) {
  double acc = 0.0;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n;
       c += (int64_t)gridDim.x * blockDim.x) {
    const bool o = mask[c] != 0;
    const bool x1 = resp[c] > 0.5f;
    const float p = prob[c];
    const float pc = fminf(fmaxf(p, kEps32), 1.0f - kEps32);
    const bool inside = (p >= kEps32) && (p <= 1.0f - kEps32);
    const float ll = x1 ? logf(pc) : log1pf(-pc);
    if (o) acc += (double)ll;
    if (g_prob) g_prob[c] = (o && inside) ? (x1 ? 1.0f / pc : -1.0f / (1.0f - pc)) : 0.0f;
  }
  block_sum_to(acc, part_ll + blockIdx.x);
}

// ---------------------------------------------------------------------------
// (n1, n_observed) per person: fallback of stream_counts for unaligned rows / I > 2048
// (one warp per row, fixed summation order)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) person_counts_kernel(int64_t P, int I, const float* __restrict__ resp,
                                                            const uint8_t* __restrict__ mask,
                                                            float* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < P; row += nwarps) {
    int n1 = 0, no = 0;
    for (int j = lane; j < I; j += 32) {
      const bool o = mask[row * I + j] != 0;
      no += o ? 1 : 0;
      n1 += (o && resp[row * I + j] > 0.5f) ? 1 : 0;
    }
    n1 = __reduce_add_sync(0xffffffffu, n1);
    no = __reduce_add_sync(0xffffffffu, no);
    if (lane == 0) {
      counts[row * 2] = (float)n1;
      counts[row * 2 + 1] = (float)no;
    }
  }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static unsigned long long g_launches = 0;
void note_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

static int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

static inline int slab_m(int D) { return D <= 4 ? 4 : 2; }

int general_max_items(int D) { return 32 * slab_m(D) * 16; }  // NS <= 16 warps per CTA

int general_grid(int64_t P, int I, int D) {
  const int ns = (I + 32 * slab_m(D) - 1) / (32 * slab_m(D));
  const int ctas_per_sm = ns >= 16 ? 1 : (ns >= 8 ? 2 : (ns >= 4 ? 4 : 8));
  const int64_t tiles = (P + kRowsPerTile - 1) / kRowsPerTile;
  int64_t g = (int64_t)sm_count() * ctas_per_sm;
  if (g > tiles) g = tiles;
  if (g < 1) g = 1;
  return (int)g;
}

template <int D>
static cudaError_t launch_encode_d(const vibo_desc& d, const float* resp, const uint8_t* mask,
                                   const float* table, float* mu, float* lv, float* S,
                                   cudaStream_t st) {
  constexpr int M = D <= 4 ? 4 : 2;
  const int ns = (d.num_item + 32 * M - 1) / (32 * M);
  const int grid = general_grid(d.num_person, d.num_item, D);
  const size_t smem = (size_t)kRowsPerTile * ns * 2 * D * sizeof(float);
  encode_kernel<D, M><<<grid, ns * 32, smem, st>>>(d.num_person, d.num_item, d.conditional,
                                                   d.missing_policy, resp, mask, table, mu, lv, S);
  note_launch();
  return cudaGetLastError();
}

template <int D>
static cudaError_t launch_encode_bwd_d(const vibo_desc& d, const float* resp, const uint8_t* mask,
                                       const float* table, const float* amu, const float* S,
                                       const float* g_mu, const float* g_lv, float* g_table,
                                       float* part, cudaStream_t st) {
  constexpr int M = D <= 4 ? 4 : 2;
  const int ns = (d.num_item + 32 * M - 1) / (32 * M);
  const int grid = general_grid(d.num_person, d.num_item, D);
  encode_bwd_kernel<D, M><<<grid, ns * 32, 0, st>>>(d.num_person, d.num_item, resp, mask, amu, S,
                                                    g_mu, g_lv, part);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int n = 2 * (d.conditional ? d.num_item : 1) * D;
  encode_bwd_finalize_kernel<<<(n + 31) / 32, 32 * kSumSlices, 0, st>>>(d.num_item, D, d.conditional, grid,
                                                              part, table, g_table);
  note_launch(2);
  return cudaGetLastError();
}

template <int D, int MODEL>
static cudaError_t launch_link_dm(const vibo_desc& d, const float* resp, const uint8_t* mask,
                                  const float* ability, const float* item_feat, double* out_ll,
                                  float* g_ability, float* g_item, double* part_ll, float* part_g,
                                  cudaStream_t st) {
  constexpr int M = D <= 4 ? 4 : 2;
  constexpr int F = item_width(MODEL, D);
  const int ns = (d.num_item + 32 * M - 1) / (32 * M);
  const int grid = general_grid(d.num_person, d.num_item, D);
  const size_t smem = (size_t)kRowsPerTile * ns * D * sizeof(float);
  if (g_item != nullptr) {
    link_kernel<D, MODEL, M, true><<<grid, ns * 32, smem, st>>>(
        d.num_person, d.num_item, resp, mask, ability, item_feat, part_ll, g_ability, part_g);
  } else {
    link_kernel<D, MODEL, M, false><<<grid, ns * 32, smem, st>>>(
        d.num_person, d.num_item, resp, mask, ability, item_feat, part_ll, nullptr, nullptr);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sum_partials_f64_kernel<<<1, 128, 0, st>>>(part_ll, grid, 1, 1, out_ll);
  note_launch(2);
  if (g_item != nullptr) {
    const int n = d.num_item * F;
    sum_partials_f32_kernel<<<(n + 31) / 32, 32 * kSumSlices, 0, st>>>(part_g, grid, n, 1.0f, g_item);
    note_launch();
  }
  return cudaGetLastError();
}

#define VIBO_SWITCH_D(D_, ...)                      \
  switch (D_) {                                     \
    case 1: { constexpr int kD = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int kD = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int kD = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int kD = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int kD = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int kD = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int kD = 7; __VA_ARGS__; } break; \
    case 8: { constexpr int kD = 8; __VA_ARGS__; } break; \
    default: return cudaErrorInvalidValue;          \
  }

cudaError_t launch_encode(const vibo_desc& d, const float* resp, const uint8_t* mask,
                          const float* table, float* mu, float* lv, float* S, cudaStream_t st) {
  cudaError_t e = launch_encode_rows_uncond(d, resp, mask, table, mu, lv, S, nullptr, st);
  if (e == cudaErrorNotSupported) {
    (void)cudaGetLastError();
    e = stream_encode(d, resp, mask, table, mu, lv, S, nullptr, st);
  }
  if (e != cudaErrorNotSupported) {
    if (e == cudaSuccess) note_launch();
    return e;
  }
  (void)cudaGetLastError();
  e = cudaSuccess;
  VIBO_SWITCH_D(d.ability_dim, e = launch_encode_d<kD>(d, resp, mask, table, mu, lv, S, st));
  return e;
}

cudaError_t launch_encode_bwd(const vibo_desc& d, const float* resp, const uint8_t* mask,
                              const float* table, const float* amu, const float* S,
                              const float* g_mu, const float* g_lv, float* g_table, float* part,
                              cudaStream_t st) {
  int grid = 0;
  cudaError_t e = stream_encode_bwd(d, resp, mask, amu, S, g_mu, g_lv, part, &grid, st);
  if (e != cudaErrorNotSupported) {
    if (e != cudaSuccess) return e;
    const int n = 2 * (d.conditional ? d.num_item : 1) * d.ability_dim;
    // the stream kernel already summed over items for the unconditional table
    encode_bwd_finalize_kernel<<<(n + 31) / 32, 32 * kSumSlices, 0, st>>>(d.conditional ? d.num_item : 1, d.ability_dim,
                                                                d.conditional, grid, part, table, g_table);
    note_launch(2);
    return cudaGetLastError();
  }
  (void)cudaGetLastError();
  e = cudaSuccess;
  VIBO_SWITCH_D(d.ability_dim,
                e = launch_encode_bwd_d<kD>(d, resp, mask, table, amu, S, g_mu, g_lv, g_table, part, st));
  return e;
}

// cudaErrorNotSupported unless the posterior is unconditional and the rows are narrow.
cudaError_t launch_encode_rows_uncond(const vibo_desc& d, const float* resp, const uint8_t* mask,
                                      const float* table, float* mu, float* lv, float* S, float* counts,
                                      cudaStream_t st) {
  const char* off = getenv("VIBO_DISABLE_ROWWARP");
  if (off != nullptr && off[0] == '1') return cudaErrorNotSupported;
  const char* off2 = getenv("VIBO_DISABLE_STREAM");   // the path-coverage tests' switch to the legacy kernels
  if (off2 != nullptr && off2[0] == '1') return cudaErrorNotSupported;
  if (d.conditional || d.num_item > 256 || d.num_person < 1) return cudaErrorNotSupported;
  if (mu == nullptr && counts == nullptr) return cudaErrorNotSupported;
  int64_t blocks = (d.num_person + 31) / 32;   // 8 warps x 4 rows per block and trip
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaError_t e = cudaSuccess;
  VIBO_SWITCH_D(mu != nullptr ? d.ability_dim : 1,
                (encode_rows_uncond_kernel<kD><<<(int)blocks, 256, 0, st>>>(d.num_person, d.num_item, d.missing_policy,
                                                                         resp, mask, table, mu, lv, S, counts)));
  (void)e;
  return cudaGetLastError();
}

cudaError_t launch_encode_counts(const vibo_desc& d, const float* resp, const uint8_t* mask,
                                 const float* table, float* mu, float* lv, float* S, float* counts,
                                 cudaStream_t st) {
  const char* off = getenv("VIBO_DISABLE_COUNTS_BWD");
  if (off != nullptr && off[0] == '1') return cudaErrorNotSupported;
  cudaError_t e = launch_encode_rows_uncond(d, resp, mask, table, mu, lv, S, counts, st);
  if (e == cudaErrorNotSupported) {
    (void)cudaGetLastError();
    e = stream_encode_counts(d, resp, mask, table, mu, lv, S, counts, st);
  }
  if (e == cudaSuccess) note_launch();
  return e;
}

cudaError_t launch_encode_bwd_counts(const vibo_desc& d, const float* counts, const float* table,
                                     const float* amu, const float* S, const float* g_mu, const float* g_lv,
                                     float* g_table, float* part, cudaStream_t st) {
  if (d.conditional) return cudaErrorNotSupported;
  int64_t blocks = (d.num_person + 255) / 256;
  const int64_t cap = (int64_t)sm_count();   // few partials: the finalize walks them in a dependent chain
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int grid = (int)blocks;
  cudaError_t e = cudaSuccess;
  VIBO_SWITCH_D(d.ability_dim,
                (encode_bwd_counts_kernel<kD><<<grid, 256, 0, st>>>(d.num_person, counts, amu, S, g_mu, g_lv, part)));
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int n = 2 * d.ability_dim;
  encode_bwd_finalize_kernel<<<(n + 31) / 32, 32 * kSumSlices, 0, st>>>(1, d.ability_dim, 0, grid, part, table, g_table);
  note_launch(2);
  return cudaGetLastError();
}

cudaError_t launch_link(const vibo_desc& d, const float* resp, const uint8_t* mask,
                        const float* ability, const float* item_feat, double* out_ll,
                        float* g_ability, float* g_item, double* part_ll, float* part_g,
                        cudaStream_t st) {
  int grid = 0;
  cudaError_t e = stream_link(d, resp, mask, ability, item_feat, part_ll, g_ability, part_g, g_item != nullptr,
                              &grid, st);
  if (e != cudaErrorNotSupported) {
    if (e != cudaSuccess) return e;
    sum_partials_f64_kernel<<<1, 128, 0, st>>>(part_ll, grid, 1, 1, out_ll);
    note_launch(2);
    if (g_item != nullptr) {
      const int n = d.num_item * item_width_host(d.irt_model, d.ability_dim);
      sum_partials_f32_kernel<<<(n + 31) / 32, 32 * kSumSlices, 0, st>>>(part_g, grid, n, 1.0f, g_item);
      note_launch();
    }
    return cudaGetLastError();
  }
  (void)cudaGetLastError();
  e = cudaSuccess;
  switch (d.irt_model) {
    case 1:
      VIBO_SWITCH_D(d.ability_dim, e = (launch_link_dm<kD, 1>(d, resp, mask, ability, item_feat, out_ll,
                                                            g_ability, g_item, part_ll, part_g, st)));
      break;
    case 2:
      VIBO_SWITCH_D(d.ability_dim, e = (launch_link_dm<kD, 2>(d, resp, mask, ability, item_feat, out_ll,
                                                            g_ability, g_item, part_ll, part_g, st)));
      break;
    case 3:
      VIBO_SWITCH_D(d.ability_dim, e = (launch_link_dm<kD, 3>(d, resp, mask, ability, item_feat, out_ll,
                                                            g_ability, g_item, part_ll, part_g, st)));
      break;
    default:
      return cudaErrorInvalidValue;
  }
  return e;
}

cudaError_t launch_person_counts(const vibo_desc& d, const float* resp, const uint8_t* mask, float* counts,
                                 cudaStream_t st) {
  cudaError_t e = launch_encode_rows_uncond(d, resp, mask, nullptr, nullptr, nullptr, nullptr, counts, st);
  if (e == cudaErrorNotSupported) {
    (void)cudaGetLastError();
    e = stream_counts(d, resp, mask, counts, st);
  }
  if (e != cudaErrorNotSupported) {
    if (e == cudaSuccess) note_launch();
    return e;
  }
  (void)cudaGetLastError();
  int64_t blocks = (d.num_person + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  person_counts_kernel<<<(int)blocks, 256, 0, st>>>(d.num_person, d.num_item, resp, mask, counts);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_decode(const vibo_desc& d, const float* ability, const float* item_feat,
                          float* out, cudaStream_t st) {
  const int64_t n = d.num_person * (int64_t)d.num_item;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  switch (d.irt_model) {
    case 1: decode_kernel<1><<<(int)blocks, 256, 0, st>>>(d.num_person, d.num_item, d.ability_dim, ability, item_feat, out); break;
    case 2: decode_kernel<2><<<(int)blocks, 256, 0, st>>>(d.num_person, d.num_item, d.ability_dim, ability, item_feat, out); break;
    case 3: decode_kernel<3><<<(int)blocks, 256, 0, st>>>(d.num_person, d.num_item, d.ability_dim, ability, item_feat, out); break;
    default: return cudaErrorInvalidValue;
  }
  note_launch();
  return cudaGetLastError();
}

int bernoulli_grid(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

cudaError_t launch_bernoulli_ll(const vibo_desc& d, const float* resp, const uint8_t* mask,
                                const float* prob, double* out_ll, float* g_prob, double* part_ll,
                                cudaStream_t st) {
  const int64_t n = d.num_person * (int64_t)d.num_item;
  const int grid = bernoulli_grid(n);
  bernoulli_ll_kernel<<<grid, 256, 0, st>>>(n, resp, mask, prob, part_ll, g_prob);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sum_partials_f64_kernel<<<1, 128, 0, st>>>(part_ll, grid, 1, 1, out_ll);
  note_launch(2);
  return cudaGetLastError();
}

}  // namespace vibo
