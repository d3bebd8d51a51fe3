// Per-person prior / reparameterisation kernels used when vibo_fused_elbo is
// composed from the general kernels (encode -> person_forward -> link ->
// person_backward -> encode_backward).  One thread per person.
//
//   theta = mu + exp(logvar / 2) * eps                 models.py:506-510
//   KL    = -1/2 sum_d (1 + logvar - mu^2 - e^logvar)  utils.py:85-88
//   sample form: log p(theta) - log q(theta)           models.py:433-435, utils.py:59-67
#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

__global__ void person_forward_kernel(int64_t P, int D, int form, int64_t person_offset,
                                      const float* __restrict__ amu, const float* __restrict__ alv,
                                      const float* __restrict__ eps_in, uint64_t seed,
                                      const uint64_t* __restrict__ seed_dev, float* __restrict__ eps_out, float* __restrict__ ability,
                                      double* __restrict__ part_term) {
  double acc = 0.0;
  if (seed_dev != nullptr) seed = seed_dev[0] + seed_dev[1];  // {seed, step} in device memory (graph replays)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    float nrm[4];
    for (int d = 0; d < D; ++d) {
      float e;
      if (eps_in != nullptr) {
        e = eps_in[i * D + d];
      } else {
        if ((d & 3) == 0) philox_normal4(seed, (uint64_t)(person_offset + i), (uint32_t)(d >> 2), nrm);
        e = nrm[d & 3];
      }
      const float m = amu[i * D + d], lv = alv[i * D + d];
      const float sd = expf(0.5f * lv);
      const float th = fmaf(e, sd, m);
      ability[i * D + d] = th;
      if (eps_out != nullptr) eps_out[i * D + d] = e;
      if (form == VIBO_ELBO_KL) {
        acc += (double)(-0.5f * (1.0f + lv - m * m - expf(lv)));
      } else {
        const float log_p = -0.5f * th * th - kHalfLog2Pi;
        const float diff = th - m;
        const float log_q = -(diff * diff) / (2.0f * sd * sd) - logf(sd) - kHalfLog2Pi;
        acc += (double)(log_p - log_q);
      }
    }
  }
  __shared__ double s_part[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
    part_term[blockIdx.x] = t;
  }
}

// out = sum_p part[p]: 256 threads take strided partial sums, then a shared-memory tree in a
// fixed order (deterministic).
__global__ void __launch_bounds__(256) sum_term_kernel(const double* __restrict__ part, int n,
                                                       double* __restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int p = threadIdx.x; p < n; p += 256) s += part[p];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// d loss_k / d (ability_mu, ability_logvar) from d LL / d theta
// (SURVEY.md Appendix A "Backward").
__global__ void person_backward_kernel(int64_t n, int form, float beta, const float* __restrict__ amu,
                                       const float* __restrict__ alv, const float* __restrict__ eps,
                                       const float* __restrict__ ability,
                                       const float* __restrict__ g_ll_ability,
                                       float* __restrict__ g_mu, float* __restrict__ g_lv) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x) {
    const float lv = alv[k];
    const float sd = expf(0.5f * lv);
    float gth = -g_ll_ability[k];  // loss_k carries -LL
    if (form == VIBO_ELBO_KL) {
      g_mu[k] = fmaf(beta, amu[k], gth);
      g_lv[k] = 0.5f * gth * eps[k] * sd + 0.5f * beta * (expf(lv) - 1.0f);
    } else {
      gth += ability[k];
      g_mu[k] = gth;
      g_lv[k] = 0.5f * gth * eps[k] * sd - 0.5f;
    }
  }
}

// Standard normals keyed by (seed, person_offset + row): the same stream the fused kernels draw
// in-kernel, for the composed paths (flows, mean merge) that need eps as a tensor.
__global__ void philox_fill_kernel(int64_t P, int D, int64_t person_offset, uint64_t seed,
                                   const uint64_t* __restrict__ seed_dev, float* __restrict__ eps) {
  if (seed_dev != nullptr) seed = seed_dev[0] + seed_dev[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    float nrm[4];
    for (int d = 0; d < D; ++d) {
      if ((d & 3) == 0) philox_normal4(seed, (uint64_t)(person_offset + i), (uint32_t)(d >> 2), nrm);
      eps[i * D + d] = nrm[d & 3];
    }
  }
}

cudaError_t launch_philox_fill(int64_t P, int D, int64_t person_offset, uint64_t seed,
                               const uint64_t* seed_dev, float* eps, cudaStream_t st) {
  philox_fill_kernel<<<person_grid(P), 256, 0, st>>>(P, D, person_offset, seed, seed_dev, eps);
  note_launch();
  return cudaGetLastError();
}

__global__ void negate_kernel(float* v, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) v[k] = -v[k];
}

__global__ void accumulate_kernel(float* dst, const float* src, int n, double* dst2, const double* src2,
                                  int n2) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) dst[k] += src[k];
  if (blockIdx.x == 0 && (int)threadIdx.x < n2) dst2[threadIdx.x] += src2[threadIdx.x];
}

cudaError_t launch_accumulate(float* dst, const float* src, int n, double* dst2, const double* src2,
                              int n2, cudaStream_t st) {
  int blocks = (n + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 1024) blocks = 1024;
  accumulate_kernel<<<blocks, 256, 0, st>>>(dst, src, n, dst2, src2, n2);
  note_launch();
  return cudaGetLastError();
}

int person_grid(int64_t P) {
  int64_t b = (P + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

cudaError_t launch_person_forward(const vibo_desc& d, const float* amu, const float* alv,
                                  const float* eps_or_null, uint64_t seed, const uint64_t* seed_dev,
                                  float* eps_out, float* ability, double* part_term, double* out_term,
                                  cudaStream_t st) {
  const int grid = person_grid(d.num_person);
  person_forward_kernel<<<grid, 256, 0, st>>>(d.num_person, d.ability_dim, d.elbo_form,
                                              d.person_offset, amu, alv, eps_or_null, seed, seed_dev, eps_out,
                                              ability, part_term);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sum_term_kernel<<<1, 256, 0, st>>>(part_term, grid, out_term);
  note_launch(2);
  return cudaGetLastError();
}

cudaError_t launch_person_backward(const vibo_desc& d, float beta, const float* amu,
                                   const float* alv, const float* eps, const float* ability,
                                   const float* g_ll_ability, float* g_mu, float* g_lv,
                                   cudaStream_t st) {
  const int64_t n = d.num_person * (int64_t)d.ability_dim;
  person_backward_kernel<<<person_grid(n), 256, 0, st>>>(n, d.elbo_form, beta, amu, alv, eps, ability,
                                                         g_ll_ability, g_mu, g_lv);
  note_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Planar flows on the abilities (--n-norm-flows K; reference flows.py:21-41,58-66 and the
// flow form of the ELBO, models.py:406-424), fused per person:
//   theta_0 = mu + eps exp(lv / 2);  for k < K:  a = w_k . z + b_k, h = tanh(a),
//   z <- z + uhat_k h,  ldj_k = log(|1 + (1 - h^2) (w_k . uhat_k)| + 1e-8)
//   term_i = sum_d (-theta_K^2 / 2 + eps^2 / 2 + lv / 2) + sum_k ldj_k
// which is  log N(theta_K; 0, 1) - log N(theta_0; mu, exp lv) + sum_k ldj_k  (the log 2 pi's cancel).
// uhat (the invertibility-corrected u) is formed by the caller from (u, w): a handful of
// parameter-only operations that stay in autograd.
// ---------------------------------------------------------------------------
constexpr int kFlowMaxK = 8, kFlowMaxD = 8;

__device__ __forceinline__ void block_sum_to_p(double v, double* dst) {
  __shared__ double s_part[32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) t += s_part[w2];
    *dst = t;
  }
}

__device__ __forceinline__ float flow_forward(int D, int K, const float* __restrict__ uhat,
                                              const float* __restrict__ w, const float* __restrict__ b,
                                              const float* wu, float* z, float (*zin)[kFlowMaxD], float* hs) {
  float ldj = 0.0f;
  for (int k = 0; k < K; ++k) {
    float a = b[k];
    for (int d = 0; d < D; ++d) {
      if (zin) zin[k][d] = z[d];
      a = fmaf(w[k * D + d], z[d], a);
    }
    const float h = tanhf(a);
    if (hs) hs[k] = h;
    for (int d = 0; d < D; ++d) z[d] = fmaf(uhat[k * D + d], h, z[d]);
    ldj += logf(fabsf(1.0f + (1.0f - h * h) * wu[k]) + 1e-8f);
  }
  return ldj;
}

__global__ void __launch_bounds__(256) flow_person_forward_kernel(int64_t P, int D, int K,
                                                                  const float* __restrict__ amu,
                                                                  const float* __restrict__ alv,
                                                                  const float* __restrict__ eps,
                                                                  const float* __restrict__ uhat,
                                                                  const float* __restrict__ w,
                                                                  const float* __restrict__ b,
                                                                  float* __restrict__ ability0,
                                                                  float* __restrict__ ability_k,
                                                                  double* __restrict__ part_term) {
  float wu[kFlowMaxK];
  for (int k = 0; k < K; ++k) {
    float v = 0.0f;
    for (int d = 0; d < D; ++d) v = fmaf(w[k * D + d], uhat[k * D + d], v);
    wu[k] = v;
  }
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float z[kFlowMaxD], term = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float lv = alv[i * D + d], e = eps[i * D + d];
      z[d] = fmaf(e, expf(0.5f * lv), amu[i * D + d]);
      if (ability0) ability0[i * D + d] = z[d];
      term += 0.5f * e * e + 0.5f * lv;
    }
    term += flow_forward(D, K, uhat, w, b, wu, z, nullptr, nullptr);
    for (int d = 0; d < D; ++d) {
      ability_k[i * D + d] = z[d];
      term -= 0.5f * z[d] * z[d];
    }
    acc += (double)term;
  }
  block_sum_to_p(acc, part_term + blockIdx.x);
}

// g_term: d loss / d (sum_i term_i) (device scalar); g_ability_k: d loss / d theta_K from the link.
// part_g: [grid][K * (2 D + 1)] = (g_uhat | g_w | g_b) partial sums.
__global__ void __launch_bounds__(256) flow_person_backward_kernel(int64_t P, int D, int K,
                                                                   const float* __restrict__ amu,
                                                                   const float* __restrict__ alv,
                                                                   const float* __restrict__ eps,
                                                                   const float* __restrict__ uhat,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ b,
                                                                   const float* __restrict__ g_ability_k,
                                                                   const float* __restrict__ g_term,
                                                                   float* __restrict__ g_mu,
                                                                   float* __restrict__ g_lv,
                                                                   float* __restrict__ part_g) {
  __shared__ float s_red[8];
  const int NP = K * (2 * D + 1);
  float wu[kFlowMaxK];
  for (int k = 0; k < K; ++k) {
    float v = 0.0f;
    for (int d = 0; d < D; ++d) v = fmaf(w[k * D + d], uhat[k * D + d], v);
    wu[k] = v;
  }
  const float c = g_term[0];
  float gu[kFlowMaxK][kFlowMaxD], gw[kFlowMaxK][kFlowMaxD], gb[kFlowMaxK];
  for (int k = 0; k < K; ++k) {
    gb[k] = 0.0f;
    for (int d = 0; d < D; ++d) gu[k][d] = gw[k][d] = 0.0f;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float z[kFlowMaxD], zin[kFlowMaxK][kFlowMaxD], hs[kFlowMaxK], sd[kFlowMaxD], ev[kFlowMaxD];
    for (int d = 0; d < D; ++d) {
      sd[d] = expf(0.5f * alv[i * D + d]);
      ev[d] = eps[i * D + d];
      z[d] = fmaf(ev[d], sd[d], amu[i * D + d]);
    }
    flow_forward(D, K, uhat, w, b, wu, z, zin, hs);
    float G[kFlowMaxD];
    for (int d = 0; d < D; ++d) G[d] = g_ability_k[i * D + d] - c * z[d];   // d term / d theta_K = -theta_K
    for (int k = K - 1; k >= 0; --k) {
      const float h = hs[k], omh = 1.0f - h * h;
      const float q = 1.0f + omh * wu[k];
      const float dl = c * (q >= 0.0f ? 1.0f : -1.0f) / (fabsf(q) + 1e-8f);   // d loss / d s
      float gdot = 0.0f;
      for (int d = 0; d < D; ++d) gdot = fmaf(G[d], uhat[k * D + d], gdot);
      const float ga = gdot * omh + dl * (-2.0f * h * omh) * wu[k];
      for (int d = 0; d < D; ++d) {
        gu[k][d] += G[d] * h + dl * omh * w[k * D + d];
        gw[k][d] += ga * zin[k][d] + dl * omh * uhat[k * D + d];
        G[d] = fmaf(ga, w[k * D + d], G[d]);
      }
      gb[k] += ga;
    }
    for (int d = 0; d < D; ++d) {
      g_mu[i * D + d] = G[d];
      g_lv[i * D + d] = 0.5f * G[d] * ev[d] * sd[d] + 0.5f * c;
    }
  }
  // block sums of the parameter gradients, fixed order
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = 0; n < NP; ++n) {
    const int k = n / (2 * D + 1), r = n % (2 * D + 1);
    float v = r < D ? gu[k][r] : (r < 2 * D ? gw[k][r - D] : gb[k]);
    v = warp_sum(v);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
      for (int q = 0; q < 8; ++q) t += s_red[q];
      part_g[(size_t)blockIdx.x * NP + n] = t;
    }
    __syncthreads();
  }
}

// out[k] = sum_p part[p][k] for the flow parameter gradients (n small): one block per output, strided
// partial sums then a shared-memory tree in a fixed order (deterministic)
__global__ void __launch_bounds__(128) flow_sum_kernel(const float* __restrict__ part, int nparts, int n, int D, int K,
                                                       float* __restrict__ g_uhat, float* __restrict__ g_w,
                                                       float* __restrict__ g_b) {
  __shared__ double sh[128];
  const int t = blockIdx.x;
  double s = 0.0;
  for (int p = threadIdx.x; p < nparts; p += 128) s += part[(size_t)p * n + t];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float v = (float)sh[0];
    const int k = t / (2 * D + 1), r = t % (2 * D + 1);
    if (r < D) g_uhat[k * D + r] = v;
    else if (r < 2 * D) g_w[k * D + (r - D)] = v;
    else g_b[k] = v;
  }
}

int flow_grid(int64_t P) {
  int64_t g = (P + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 4;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

cudaError_t launch_flow_person_forward(int64_t P, int D, int K, const float* amu, const float* alv,
                                       const float* eps, const float* uhat, const float* w, const float* b,
                                       float* ability0, float* ability_k, double* part_term, double* out_term,
                                       cudaStream_t st) {
  if (D > kFlowMaxD || K > kFlowMaxK || K < 1) return cudaErrorInvalidValue;
  const int grid = flow_grid(P);
  flow_person_forward_kernel<<<grid, 256, 0, st>>>(P, D, K, amu, alv, eps, uhat, w, b, ability0, ability_k,
                                                   part_term);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sum_term_kernel<<<1, 256, 0, st>>>(part_term, grid, out_term);
  note_launch(2);
  return cudaGetLastError();
}

cudaError_t launch_flow_person_backward(int64_t P, int D, int K, const float* amu, const float* alv,
                                        const float* eps, const float* uhat, const float* w, const float* b,
                                        const float* g_ability_k, const float* g_term, float* g_mu, float* g_lv,
                                        float* g_uhat, float* g_w, float* g_b, float* part_g, cudaStream_t st) {
  if (D > kFlowMaxD || K > kFlowMaxK || K < 1) return cudaErrorInvalidValue;
  const int grid = flow_grid(P);
  flow_person_backward_kernel<<<grid, 256, 0, st>>>(P, D, K, amu, alv, eps, uhat, w, b, g_ability_k, g_term, g_mu,
                                                    g_lv, part_g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int n = K * (2 * D + 1);
  flow_sum_kernel<<<n, 128, 0, st>>>(part_g, grid, n, D, K, g_uhat, g_w, g_b);
  note_launch(2);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Planar-flow parameter chain (flows.py:26-29): the K flows keep separate (u, w, b) parameters (state_dict
// keys flows.{k}.{u,w,b}); the fused per-row flow kernels take them stacked, with the invertibility correction
//   uhat_k = u_k + c_k w_k,   c_k = (softplus(a_k) - 1 - a_k) / n_k,   a_k = w_k . u_k,   n_k = |w_k|^2.
// One thread per flow gathers its three parameter tensors and forms (uhat, w, b) rows / their gradients:
// two launches instead of ~13 + ~25 elementwise PyTorch launches per flow stack and step.
// ---------------------------------------------------------------------------
struct PlanarPtrs {
  const float* u[kFlowMaxK];
  const float* w[kFlowMaxK];
  const float* b[kFlowMaxK];
};

// torch.nn.functional.softplus (beta 1, threshold 20) and its derivative
__device__ __forceinline__ float softplus_t(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_grad_t(float x) {
  if (x > 20.0f) return 1.0f;
  const float z = expf(x);
  return z / (z + 1.0f);
}

__global__ void planar_params_forward_kernel(int K, int D, PlanarPtrs p, float* __restrict__ uhat,
                                             float* __restrict__ w_out, float* __restrict__ b_out) {
  const int k = threadIdx.x;
  if (k >= K) return;
  float a = 0.0f, n = 0.0f;
  for (int d = 0; d < D; ++d) {
    a = fmaf(p.u[k][d], p.w[k][d], a);
    n = fmaf(p.w[k][d], p.w[k][d], n);
  }
  const float c = (softplus_t(a) - 1.0f - a) / n;
  for (int d = 0; d < D; ++d) {
    uhat[k * D + d] = fmaf(c, p.w[k][d], p.u[k][d]);
    w_out[k * D + d] = p.w[k][d];
  }
  b_out[k] = p.b[k][0];
}

// g_u (K, D), g_w (K, D), g_b (K): gradients of the separate parameters given those of (uhat, w_out, b_out)
__global__ void planar_params_backward_kernel(int K, int D, PlanarPtrs p, const float* __restrict__ g_uhat,
                                              const float* __restrict__ g_w_out, const float* __restrict__ g_b_out,
                                              float* __restrict__ g_u, float* __restrict__ g_w,
                                              float* __restrict__ g_b) {
  const int k = threadIdx.x;
  if (k >= K) return;
  float a = 0.0f, n = 0.0f, gw = 0.0f;   // gw = g_uhat . w
  for (int d = 0; d < D; ++d) {
    a = fmaf(p.u[k][d], p.w[k][d], a);
    n = fmaf(p.w[k][d], p.w[k][d], n);
    gw = fmaf(g_uhat[k * D + d], p.w[k][d], gw);
  }
  const float c = (softplus_t(a) - 1.0f - a) / n;
  const float dc_da = (softplus_grad_t(a) - 1.0f) / n, dc_dn = -c / n;
  for (int d = 0; d < D; ++d) {
    const float ud = p.u[k][d], wd = p.w[k][d], gh = g_uhat[k * D + d];
    g_u[k * D + d] = fmaf(gw * dc_da, wd, gh);
    g_w[k * D + d] = g_w_out[k * D + d] + c * gh + gw * (dc_da * ud + 2.0f * dc_dn * wd);
  }
  g_b[k] = g_b_out[k];
}

cudaError_t launch_planar_params_forward(int K, int D, const float* const* u, const float* const* w,
                                         const float* const* b, float* uhat, float* w_out, float* b_out,
                                         cudaStream_t st) {
  if (D > kFlowMaxD || K > kFlowMaxK || K < 1 || D < 1) return cudaErrorInvalidValue;
  PlanarPtrs p;
  for (int k = 0; k < kFlowMaxK; ++k) {
    p.u[k] = k < K ? u[k] : nullptr;
    p.w[k] = k < K ? w[k] : nullptr;
    p.b[k] = k < K ? b[k] : nullptr;
  }
  planar_params_forward_kernel<<<1, 32, 0, st>>>(K, D, p, uhat, w_out, b_out);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_planar_params_backward(int K, int D, const float* const* u, const float* const* w,
                                          const float* g_uhat, const float* g_w_out, const float* g_b_out,
                                          float* g_u, float* g_w, float* g_b, cudaStream_t st) {
  if (D > kFlowMaxD || K > kFlowMaxK || K < 1 || D < 1) return cudaErrorInvalidValue;
  PlanarPtrs p;
  for (int k = 0; k < kFlowMaxK; ++k) {
    p.u[k] = k < K ? u[k] : nullptr;
    p.w[k] = k < K ? w[k] : nullptr;
    p.b[k] = nullptr;
  }
  planar_params_backward_kernel<<<1, 32, 0, st>>>(K, D, p, g_uhat, g_w_out, g_b_out, g_u, g_w, g_b);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_negate(float* v, int n, cudaStream_t st) {
  negate_kernel<<<(n + 255) / 256, 256, 0, st>>>(v, n);
  note_launch();
  return cudaGetLastError();
}

}  // namespace vibo
