// Sample-loop kernels: the S-sample Monte-Carlo estimators of the reference's evaluation
// closures with the loop over samples INSIDE the kernel, the row tile resident in shared memory.
//
//   vibo_log_marginal     models.py:445-504 (IWAE bound): S importance samples, each a fresh item
//                         draw d_s ~ q(d) and ability draws theta_is ~ q(theta_i | x_i); weight
//                         log w_s = sum_i [ LL_i(theta_is, d_s) + log p(theta_is) - log q(theta_is) ]
//                                   + log p(d_s) - log q(d_s)      (elbo(use_kl_divergence=False), :432-441)
//                         result logsumexp_s(log w_s) - log S.  For the UNCONDITIONAL posterior the
//                         ability posterior does not depend on the sample, so a row is read once
//                         and re-scored S times; the reference (and round 1) re-ran the whole
//                         forward pass -- a full read of the matrix -- per sample.
//   vibo_predictive_mean  vibo.py:349-390 + :515: mean over S posterior draws of
//                         decode(theta_s, d_s) (the reference stacks (S, P, I, 1) on the host and
//                         averages later).
//
// Work layout: a CTA owns row tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... and the sample
// chunk blockIdx.y.  Per (tile, sample): all threads build the sample's item parameters in shared
// memory (Philox or injected noise), then warp w scores rows w, w + 8, ... with its lanes striding
// over the items.  Per-sample sums are combined in a fixed order (deterministic).
#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

namespace {

constexpr int kSampleThreads = 256;
constexpr int kSampleWarps = kSampleThreads / 32;
constexpr uint64_t kItemStream = 1ull << 62;

struct SampleParams {
  int64_t P;
  int I, D, F, model, S, R, missing_policy, s_chunk, mode;  // mode 0: log-marginal, 1: predictive mean
  int64_t person_offset;
  const float* resp;
  const uint8_t* mask;
  const float* table;     // (2, 1, 2D)
  const float* item_mu;   // (I, F)
  const float* item_lv;
  const float* amu;       // (P, D)  predictive mode: the given ability posterior
  const float* alv;
  const float* eps_item;     // (S, I, F) or null
  const float* eps_ability;  // (S, P, D) or null
  uint64_t seed;
  const uint64_t* seed_dev;
  double* part;        // [S][gridDim.x]
  double* item_term;   // [S]
  float* out_mean;     // (P, I)
};

// Philox normals with a stream id in the fourth counter word (0 is the training stream of
// philox_normal4): sample s uses stream s + 1.
__device__ __forceinline__ void philox_normal4s(uint64_t seed, uint64_t index, uint32_t block, uint32_t stream,
                                                float out[4]) {
  uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), block, stream};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;
    const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.0f * __logf(u1));
    const float ang = 3.14159265358979f * fmaf(2.0f, u2, -1.0f);
    out[2 * h] = -rad * __cosf(ang);
    out[2 * h + 1] = -rad * __sinf(ang);
  }
}

__host__ __device__ inline size_t sample_align(size_t v) { return (v + 15) / 16 * 16; }

struct SampleSmem {
  size_t item_off, code_off, stat_off, acc_off, total;
};
__host__ __device__ inline SampleSmem sample_layout(int I, int D, int F, int R, int S, int mode) {
  SampleSmem L;
  size_t off = 64 * sizeof(double);                       // per-warp partials + scratch
  L.item_off = off; off += sample_align((size_t)F * I * 4);                 // item sample, SoA [F][I]
  L.code_off = off; off += sample_align(mode == 0 ? (size_t)R * I : 0);     // 0 / 1 / 2 (missing)
  L.stat_off = off; off += sample_align((size_t)R * D * 3 * 4);             // amu | sd | logvar per row
  L.acc_off = off;                                                          // log-marginal: double[S]
  off += mode == 0 ? sample_align((size_t)S * 8) : sample_align((size_t)R * I * 4);  // predictive: float[R][I]
  L.total = off;
  return L;
}

template <int MODEL>
__global__ void __launch_bounds__(kSampleThreads) sample_loop_kernel(const __grid_constant__ SampleParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int I = p.I, D = p.D, F = p.F, R = p.R, S = p.S;
  const SampleSmem L = sample_layout(I, D, F, R, S, p.mode);
  double* s_warp = reinterpret_cast<double*>(smem);                     // [kSampleWarps] (+ scratch)
  float* s_item = reinterpret_cast<float*>(smem + L.item_off);          // [F][I]: a_0.. a_{D-1} | b | guess
  uint8_t* s_code = smem + L.code_off;                                  // [R][I]
  float* s_stat = reinterpret_cast<float*>(smem + L.stat_off);          // [R][3][D]
  double* s_acc = reinterpret_cast<double*>(smem + L.acc_off);          // [S]
  float* s_out = reinterpret_cast<float*>(smem + L.acc_off);            // [R][I] (predictive)
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint64_t key = p.seed_dev != nullptr ? p.seed_dev[0] + p.seed_dev[1] : p.seed;
  const int s_lo = blockIdx.y * p.s_chunk, s_hi = min(S, s_lo + p.s_chunk);
  const int64_t n_tiles = (p.P + R - 1) / R;
  const int DA = MODEL == 1 ? 0 : D;
  if (p.mode == 0)
    for (int s = t; s < S; s += kSampleThreads) s_acc[s] = 0.0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * R;
    const int rows = (int)min((int64_t)R, p.P - row0);
    __syncthreads();  // previous tile fully consumed
    // ---- tile setup: codes + per-row posterior (log-marginal) or the given posterior (predictive)
    if (p.mode == 0) {
      for (int k = t; k < rows * I; k += kSampleThreads) {
        const int64_t g = row0 * I + k;
        s_code[k] = p.mask[g] ? (p.resp[g] > 0.5f ? 1 : 0) : 2;
      }
      __syncthreads();
      for (int r = warp; r < rows; r += kSampleWarps) {
        int n0 = 0, n1 = 0;
        for (int j = lane; j < I; j += 32) {
          const int c = s_code[r * I + j];
          n0 += c == 0;
          n1 += c == 1;
        }
        n0 = __reduce_add_sync(0xffffffffu, n0);
        n1 = __reduce_add_sync(0xffffffffu, n1);
        const float nmiss = (float)(I - n0 - n1);
        const float prior_tau = p.missing_policy == VIBO_MISSING_PRIOR ? 1.0f / (1.0f + kPoeEps) : 0.0f;
        if (lane < D) {
          const int d = lane;
          const float t0 = 1.0f / (expf(p.table[D + d]) + kPoeEps), t1 = 1.0f / (expf(p.table[2 * D + D + d]) + kPoeEps);
          const float Ssum = fmaf((float)n0, t0, fmaf((float)n1, t1, nmiss * prior_tau));
          const float Nsum = fmaf((float)n0, p.table[d] * t0, (float)n1 * p.table[2 * D + d] * t1);
          s_stat[(r * 3 + 0) * D + d] = Nsum / Ssum;
          s_stat[(r * 3 + 1) * D + d] = rsqrtf(Ssum);
          s_stat[(r * 3 + 2) * D + d] = -logf(Ssum);
        }
      }
    } else {
      for (int k = t; k < rows * D; k += kSampleThreads) {
        const int r = k / D, d = k % D;
        const float lv = p.alv[(row0 + r) * D + d];
        s_stat[(r * 3 + 0) * D + d] = p.amu[(row0 + r) * D + d];
        s_stat[(r * 3 + 1) * D + d] = expf(0.5f * lv);
        s_stat[(r * 3 + 2) * D + d] = lv;
      }
      for (int k = t; k < rows * I; k += kSampleThreads) s_out[k] = 0.0f;
    }

    for (int s = s_lo; s < s_hi; ++s) {
      __syncthreads();  // stats ready / previous sample's item parameters consumed
      // ---- this sample's item parameters d_s = mu + exp(lv / 2) eps  (models.py:359-361)
      double it_acc = 0.0;
      for (int k4 = t * 4; k4 < I * F; k4 += kSampleThreads * 4) {
        float nrm[4];
        if (p.eps_item == nullptr) philox_normal4s(key, kItemStream + (uint64_t)(k4 >> 2), 0u, (uint32_t)s + 1u, nrm);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k4 + u;
          if (k >= I * F) break;
          const float e = p.eps_item != nullptr ? p.eps_item[(size_t)s * I * F + k] : nrm[u];
          const float m = p.item_mu[k], l = p.item_lv[k];
          const float dv = fmaf(e, expf(0.5f * l), m);
          const int j = k / F, f = k % F;
          s_item[f * I + j] = (MODEL == 3 && f == D + 1) ? 1.0f / (1.0f + expf(-dv)) : dv;
          it_acc += (double)(-0.5f * dv * dv + 0.5f * e * e + 0.5f * l);  // log p(d) - log q(d)
        }
      }
      if (p.mode == 0 && tile == blockIdx.x && blockIdx.x == 0) {
        // the item term of sample s is counted once: by the CTA that owns tile 0
        it_acc = warp_sum(it_acc);
        if (lane == 0) s_warp[16 + warp] = it_acc;
      }
      __syncthreads();
      if (p.mode == 0 && tile == blockIdx.x && blockIdx.x == 0 && t == 0) {
        double a = 0.0;
        for (int w = 0; w < kSampleWarps; ++w) a += s_warp[16 + w];
        p.item_term[s] = a;
      }
      // ---- rows of the tile against this sample
      double acc = 0.0;
      for (int r = warp; r < rows; r += kSampleWarps) {
        float th[VIBO_MAX_ABILITY_DIM];
        float tsum = 0.0f, pterm = 0.0f;
        const int64_t person = p.person_offset + row0 + r;
        float nrm[4];
        for (int d = 0; d < D; ++d) {
          float e;
          if (p.eps_ability != nullptr) {
            e = p.eps_ability[((size_t)s * p.P + (row0 + r)) * D + d];
          } else {
            if ((d & 3) == 0) philox_normal4s(key, (uint64_t)person, (uint32_t)(d >> 2), (uint32_t)s + 1u, nrm);
            e = nrm[d & 3];
          }
          const float m = s_stat[(r * 3 + 0) * D + d], sd = s_stat[(r * 3 + 1) * D + d];
          th[d] = fmaf(e, sd, m);
          tsum += th[d];
          pterm += -0.5f * th[d] * th[d] + 0.5f * e * e + 0.5f * s_stat[(r * 3 + 2) * D + d];
        }
        float ll = 0.0f;
        for (int j = lane; j < I; j += 32) {
          float z = s_item[DA * I + j];
          if (MODEL == 1) {
            z += tsum;
          } else {
            for (int d = 0; d < D; ++d) z = fmaf(-th[d], s_item[d * I + j], z);
          }
          if (p.mode == 0) {
            const int c = s_code[r * I + j];
            if (c != 2) {
              const CellGrad cg = MODEL == 3 ? cell_3pl<false>(z, s_item[(D + 1) * I + j], c == 1)
                                             : cell_logistic<false>(z, c == 1);
              ll += cg.ll;
            }
          } else {
            const float sg = 1.0f / (1.0f + __expf(-z));
            const float g = MODEL == 3 ? s_item[(D + 1) * I + j] : 0.0f;
            s_out[r * I + j] += MODEL == 3 ? fmaf(1.0f - g, sg, g) : sg;
          }
        }
        acc += (double)ll;
        if (lane == 0) acc += (double)pterm;
      }
      if (p.mode == 0) {
        acc = warp_sum(acc);
        if (lane == 0) s_warp[warp] = acc;
        __syncthreads();
        if (t == 0) {
          double a = 0.0;
          for (int w = 0; w < kSampleWarps; ++w) a += s_warp[w];
          s_acc[s] += a;
        }
      }
    }
    if (p.mode == 1) {
      __syncthreads();
      const float inv = 1.0f / (float)(s_hi - s_lo);
      for (int k = t; k < rows * I; k += kSampleThreads) p.out_mean[row0 * I + k] = s_out[k] * inv;
    }
  }
  if (p.mode == 0) {
    __syncthreads();
    for (int s = s_lo + t; s < s_hi; s += kSampleThreads) p.part[(size_t)s * gridDim.x + blockIdx.x] = s_acc[s];
  }
}

// log w_s = item_term[s] + sum over CTAs (fixed order); logp = logsumexp_s(log w_s) - log S.
__global__ void __launch_bounds__(256) log_marginal_finalize_kernel(int S, int nparts, const double* __restrict__ part,
                                                                    const double* __restrict__ item_term,
                                                                    double* __restrict__ log_w,
                                                                    double* __restrict__ out_logp) {
  __shared__ double sh[256];
  double mx = -1e300;
  for (int s = threadIdx.x; s < S; s += 256) {
    double a = item_term[s];
    for (int q = 0; q < nparts; ++q) a += part[(size_t)s * nparts + q];
    log_w[s] = a;
    mx = a > mx ? a : mx;
  }
  sh[threadIdx.x] = mx;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] = sh[threadIdx.x] > sh[threadIdx.x + w] ? sh[threadIdx.x] : sh[threadIdx.x + w];
    __syncthreads();
  }
  mx = sh[0];
  __syncthreads();
  double e = 0.0;
  for (int s = threadIdx.x; s < S; s += 256) e += exp(log_w[s] - mx);
  sh[threadIdx.x] = e;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_logp = mx + log(sh[0]) - log((double)S);
}

int pick_rows(int I, int D, int F, int S, int mode, int64_t P) {
  int R = mode == 0 ? 32 : 8;
  if ((int64_t)R > P) R = (int)(P < 1 ? 1 : P);
  while (R > 1 && sample_layout(I, D, F, R, S, mode).total > 200 * 1024) R /= 2;
  return R;
}

template <int MODEL>
cudaError_t launch_sample(const SampleParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  auto k = sample_loop_kernel<MODEL>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<grid, kSampleThreads, smem, st>>>(p);
  note_launch();
  return cudaGetLastError();
}

cudaError_t dispatch(const SampleParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  if (p.model == 1) return launch_sample<1>(p, grid, smem, st);
  if (p.model == 2) return launch_sample<2>(p, grid, smem, st);
  return launch_sample<3>(p, grid, smem, st);
}

}  // namespace

size_t log_marginal_workspace_bytes(int S) { return ((size_t)S * (size_t)sm_count() * 2 + (size_t)S) * sizeof(double) + 256; }

cudaError_t launch_log_marginal(const vibo_desc& d, int S, const float* resp, const uint8_t* mask, const float* table,
                                const float* item_mu, const float* item_lv, const float* eps_item,
                                const float* eps_ability, uint64_t seed, const uint64_t* seed_dev, double* log_w,
                                double* out_logp, void* ws, size_t ws_bytes, cudaStream_t st) {
  SampleParams p{};
  p.P = d.num_person; p.I = d.num_item; p.D = d.ability_dim; p.F = item_width_host(d.irt_model, d.ability_dim);
  p.model = d.irt_model; p.S = S; p.missing_policy = d.missing_policy; p.mode = 0; p.person_offset = d.person_offset;
  p.R = pick_rows(p.I, p.D, p.F, S, 0, p.P);
  const SampleSmem L = sample_layout(p.I, p.D, p.F, p.R, S, 0);
  if (L.total > 220 * 1024) return cudaErrorNotSupported;
  const int64_t n_tiles = (p.P + p.R - 1) / p.R;
  const int sms = sm_count();
  int gx = (int)(n_tiles < 2 * sms ? n_tiles : 2 * sms);
  if (gx < 1) gx = 1;
  // few row tiles (the CLI's 16-person batches): spread the samples over the SMs instead
  int gy = gx >= sms ? 1 : (sms + gx - 1) / gx;
  if (gy > S) gy = S;
  p.s_chunk = (S + gy - 1) / gy;
  gy = (S + p.s_chunk - 1) / p.s_chunk;
  if (ws_bytes < ((size_t)S * gx + S) * sizeof(double)) return cudaErrorInvalidValue;
  p.resp = resp; p.mask = mask; p.table = table; p.item_mu = item_mu; p.item_lv = item_lv;
  p.eps_item = eps_item; p.eps_ability = eps_ability; p.seed = seed; p.seed_dev = seed_dev;
  p.part = static_cast<double*>(ws);
  p.item_term = p.part + (size_t)S * gx;
  cudaError_t e = dispatch(p, dim3(gx, gy), L.total, st);
  if (e != cudaSuccess) return e;
  log_marginal_finalize_kernel<<<1, 256, 0, st>>>(S, gx, p.part, p.item_term, log_w, out_logp);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_predictive_mean(const vibo_desc& d, int S, const float* amu, const float* alv, const float* item_mu,
                                   const float* item_lv, uint64_t seed, const uint64_t* seed_dev, float* out_mean,
                                   cudaStream_t st) {
  SampleParams p{};
  p.P = d.num_person; p.I = d.num_item; p.D = d.ability_dim; p.F = item_width_host(d.irt_model, d.ability_dim);
  p.model = d.irt_model; p.S = S; p.mode = 1; p.person_offset = d.person_offset;
  p.R = pick_rows(p.I, p.D, p.F, S, 1, p.P);
  const SampleSmem L = sample_layout(p.I, p.D, p.F, p.R, S, 1);
  if (L.total > 220 * 1024) return cudaErrorNotSupported;
  const int64_t n_tiles = (p.P + p.R - 1) / p.R;
  const int sms = sm_count();
  int gx = (int)(n_tiles < 4 * sms ? n_tiles : 4 * sms);
  if (gx < 1) gx = 1;
  p.s_chunk = S;
  p.amu = amu; p.alv = alv; p.item_mu = item_mu; p.item_lv = item_lv; p.seed = seed; p.seed_dev = seed_dev;
  p.out_mean = out_mean;
  return dispatch(p, dim3(gx, 1), L.total, st);
}

}  // namespace vibo
