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
                                      float* __restrict__ eps_out, float* __restrict__ ability,
                                      double* __restrict__ part_term) {
  double acc = 0.0;
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

__global__ void sum_term_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int p = 0; p < n; ++p) s += part[p];
    *out = s;
  }
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
                                  const float* eps_or_null, uint64_t seed, float* eps_out,
                                  float* ability, double* part_term, double* out_term,
                                  cudaStream_t st) {
  const int grid = person_grid(d.num_person);
  person_forward_kernel<<<grid, 256, 0, st>>>(d.num_person, d.ability_dim, d.elbo_form,
                                              d.person_offset, amu, alv, eps_or_null, seed, eps_out,
                                              ability, part_term);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sum_term_kernel<<<1, 32, 0, st>>>(part_term, grid, out_term);
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

cudaError_t launch_negate(float* v, int n, cudaStream_t st) {
  negate_kernel<<<(n + 255) / 256, 256, 0, st>>>(v, n);
  note_launch();
  return cudaGetLastError();
}

}  // namespace vibo
