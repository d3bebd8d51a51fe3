// Parameter-side chain of the unconditional model as two small kernels, so a
// training step is ~10 graph nodes instead of ~85 tiny PyTorch kernels:
//
//   forward  : item_feat = mu + exp(logvar/2) * eps_item       (models.py:359-361, :506-510)
//              table     = mlp([0]), mlp([1])                  (models.py:575-582 on the 2 distinct
//                                                               cell inputs of AbilityInferenceNetwork)
//              item_term = KL(q(d) || N(0,1))                  (utils.py:85-88, models.py:429)  or
//                          -(log p(d) - log q(d))              (utils.py:59-67, models.py:434-441)
//   backward : chain rule of the above given d loss / d table, d loss / d item_feat
//              (from vibo_fused_elbo) and d loss / d item_term.
//
// Block 0 runs the 1 -> H -> H -> 2D MLP on the two rows (forward keeps the
// hidden activations for the backward); block 1 handles the (I, F) item side.
// Everything is deterministic (fixed summation order, no atomics).
#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

constexpr int kParamThreads = 256;
constexpr int kMaxHidden = 256;

__device__ __forceinline__ float elu(float a) { return a > 0.0f ? a : expm1f(a); }
// d ELU / d a expressed through the activation h = ELU(a): 1 for a > 0, exp(a) = h + 1 otherwise
__device__ __forceinline__ float elu_grad_from_out(float h) { return h > 0.0f ? 1.0f : h + 1.0f; }

__global__ void __launch_bounds__(kParamThreads)
param_forward_kernel(int I, int F, int D, int H, int form, const float* __restrict__ mu,
                     const float* __restrict__ lv, const float* __restrict__ eps,
                     const float* __restrict__ w0, const float* __restrict__ b0,
                     const float* __restrict__ w2, const float* __restrict__ b2,
                     const float* __restrict__ w4, const float* __restrict__ b4,
                     float* __restrict__ item_feat, float* __restrict__ table,
                     float* __restrict__ hidden /*[2 layers][2 rows][H]*/, double* __restrict__ item_term) {
  __shared__ float s_h[2][2][kMaxHidden];
  __shared__ double s_red[kParamThreads / 32];
  const int t = threadIdx.x;
  if (blockIdx.x == 0) {
    for (int k = t; k < 2 * H; k += blockDim.x) {
      const int r = k / H, h = k % H;
      s_h[0][r][h] = elu(fmaf(w0[h], (float)r, b0[h]));  // cell input is r itself: 0 or 1
    }
    __syncthreads();
    for (int k = t; k < 2 * H; k += blockDim.x) {
      const int r = k / H, h = k % H;
      float a = b2[h];
      for (int j = 0; j < H; ++j) a = fmaf(w2[(size_t)h * H + j], s_h[0][r][j], a);
      s_h[1][r][h] = elu(a);
    }
    __syncthreads();
    for (int k = t; k < 2 * 2 * D; k += blockDim.x) {
      const int r = k / (2 * D), o = k % (2 * D);
      float a = b4[o];
      for (int j = 0; j < H; ++j) a = fmaf(w4[(size_t)o * H + j], s_h[1][r][j], a);
      table[r * 2 * D + o] = a;
    }
    for (int k = t; k < 2 * 2 * H; k += blockDim.x) hidden[k] = (&s_h[0][0][0])[(k / H) * kMaxHidden + (k % H)];
    return;
  }
  double acc = 0.0;
  for (int k = t; k < I * F; k += blockDim.x) {
    const float m = mu[k], l = lv[k], e = eps[k];
    const float d = fmaf(e, expf(0.5f * l), m);
    item_feat[k] = d;
    if (form == VIBO_ELBO_KL) acc += (double)(-0.5f * (1.0f + l - m * m - expf(l)));
    else acc += (double)(0.5f * d * d - 0.5f * e * e - 0.5f * l);  // -(log p(d) - log q(d))
  }
  acc = warp_sum(acc);
  if ((t & 31) == 0) s_red[t >> 5] = acc;
  __syncthreads();
  if (t == 0) {
    double s = 0.0;
    for (int w = 0; w < kParamThreads / 32; ++w) s += s_red[w];
    *item_term = s;
  }
}

__global__ void __launch_bounds__(kParamThreads)
param_backward_kernel(int I, int F, int D, int H, int form, const float* __restrict__ mu,
                      const float* __restrict__ lv, const float* __restrict__ eps,
                      const float* __restrict__ w2, const float* __restrict__ w4,
                      const float* __restrict__ hidden, const float* __restrict__ g_table,
                      const float* __restrict__ g_item, const float* __restrict__ g_term_ptr,
                      float* __restrict__ g_mu, float* __restrict__ g_lv, float* __restrict__ g_w0,
                      float* __restrict__ g_b0, float* __restrict__ g_w2, float* __restrict__ g_b2,
                      float* __restrict__ g_w4, float* __restrict__ g_b4) {
  __shared__ float s_h[2][2][kMaxHidden];   // activations h1, h2
  __shared__ float s_ga[2][2][kMaxHidden];  // gradients w.r.t. pre-activations a1, a2
  __shared__ float s_go[2][2 * VIBO_MAX_ABILITY_DIM];
  const int t = threadIdx.x;
  if (blockIdx.x == 0) {
    for (int k = t; k < 2 * 2 * H; k += blockDim.x) (&s_h[0][0][0])[(k / H) * kMaxHidden + (k % H)] = hidden[k];
    for (int k = t; k < 2 * 2 * D; k += blockDim.x) s_go[k / (2 * D)][k % (2 * D)] = g_table[k];
    __syncthreads();
    // layer 4: out = w4 h2 + b4
    for (int k = t; k < 2 * D * H; k += blockDim.x) {
      const int o = k / H, j = k % H;
      g_w4[k] = s_go[0][o] * s_h[1][0][j] + s_go[1][o] * s_h[1][1][j];
    }
    for (int o = t; o < 2 * D; o += blockDim.x) g_b4[o] = s_go[0][o] + s_go[1][o];
    for (int k = t; k < 2 * H; k += blockDim.x) {
      const int r = k / H, j = k % H;
      float g = 0.0f;
      for (int o = 0; o < 2 * D; ++o) g = fmaf(w4[(size_t)o * H + j], s_go[r][o], g);
      s_ga[1][r][j] = g * elu_grad_from_out(s_h[1][r][j]);
    }
    __syncthreads();
    // layer 2: a2 = w2 h1 + b2
    for (int k = t; k < H * H; k += blockDim.x) {
      const int h = k / H, j = k % H;
      g_w2[k] = s_ga[1][0][h] * s_h[0][0][j] + s_ga[1][1][h] * s_h[0][1][j];
    }
    for (int h = t; h < H; h += blockDim.x) g_b2[h] = s_ga[1][0][h] + s_ga[1][1][h];
    for (int k = t; k < 2 * H; k += blockDim.x) {
      const int r = k / H, j = k % H;
      float g = 0.0f;
      for (int h = 0; h < H; ++h) g = fmaf(w2[(size_t)h * H + j], s_ga[1][r][h], g);
      s_ga[0][r][j] = g * elu_grad_from_out(s_h[0][r][j]);
    }
    __syncthreads();
    // layer 0: a1 = w0 * r + b0 with r = 0, 1
    for (int h = t; h < H; h += blockDim.x) {
      g_w0[h] = s_ga[0][1][h];
      g_b0[h] = s_ga[0][0][h] + s_ga[0][1][h];
    }
    return;
  }
  const float gt = *g_term_ptr;  // d loss / d item_term
  for (int k = t; k < I * F; k += blockDim.x) {
    const float m = mu[k], l = lv[k], e = eps[k];
    const float sd = expf(0.5f * l);
    float gd = g_item[k];  // d loss / d item_feat through the kernels
    float gm_direct, gl_direct;
    if (form == VIBO_ELBO_KL) {
      gm_direct = gt * m;          // gt already carries beta (and the 1/world_size weight)
      gl_direct = gt * 0.5f * (expf(l) - 1.0f);
    } else {
      gd += gt * fmaf(e, sd, m);   // d/d d of 1/2 d^2
      gm_direct = 0.0f;
      gl_direct = -0.5f * gt;
    }
    g_mu[k] = gd + gm_direct;
    g_lv[k] = gd * 0.5f * e * sd + gl_direct;
  }
}

cudaError_t launch_param_forward(int I, int F, int D, int H, int form, const float* mu, const float* lv,
                                 const float* eps, const float* w0, const float* b0, const float* w2,
                                 const float* b2, const float* w4, const float* b4, float* item_feat,
                                 float* table, float* hidden, double* item_term, cudaStream_t st) {
  if (H > kMaxHidden) return cudaErrorInvalidValue;
  param_forward_kernel<<<2, kParamThreads, 0, st>>>(I, F, D, H, form, mu, lv, eps, w0, b0, w2, b2, w4, b4,
                                                    item_feat, table, hidden, item_term);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_param_backward(int I, int F, int D, int H, int form, const float* mu,
                                  const float* lv, const float* eps, const float* w2, const float* w4,
                                  const float* hidden, const float* g_table, const float* g_item,
                                  const float* g_term, float* g_mu, float* g_lv, float* g_w0, float* g_b0,
                                  float* g_w2, float* g_b2, float* g_w4, float* g_b4, cudaStream_t st) {
  if (H > kMaxHidden) return cudaErrorInvalidValue;
  param_backward_kernel<<<2, kParamThreads, 0, st>>>(I, F, D, H, form, mu, lv, eps, w2, w4, hidden,
                                                     g_table, g_item, g_term, g_mu, g_lv, g_w0, g_b0, g_w2,
                                                     g_b2, g_w4, g_b4);
  note_launch();
  return cudaGetLastError();
}

}  // namespace vibo
