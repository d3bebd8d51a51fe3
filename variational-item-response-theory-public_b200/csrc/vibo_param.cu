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
constexpr uint64_t kItemNoiseStream = 1ull << 62;  // "person index" range of the item noise

// Copy a row-major (rows, cols) matrix into shared memory with row pitch `ld`: every thread first
// issues up to 16 INDEPENDENT coalesced loads, then stores them, so a 64 x 64 matrix costs one
// global-memory latency instead of sixteen dependent ones.
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int rows, int cols, int ld) {
  constexpr int B = 16;
  const int n = rows * cols, nt = blockDim.x;
  for (int k0 = threadIdx.x; k0 < n; k0 += nt * B) {
    float v[B];
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const int k = k0 + u * nt;
      v[u] = k < n ? src[k] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const int k = k0 + u * nt;
      if (k < n) dst[(k / cols) * ld + (k % cols)] = v[u];
    }
  }
}

__device__ __forceinline__ float elu(float a) { return a > 0.0f ? a : expm1f(a); }
// d ELU / d a expressed through the activation h = ELU(a): 1 for a > 0, exp(a) = h + 1 otherwise
__device__ __forceinline__ float elu_grad_from_out(float h) { return h > 0.0f ? 1.0f : h + 1.0f; }

__global__ void __launch_bounds__(kParamThreads)
param_forward_kernel(int I, int F, int D, int H, int form, const float* __restrict__ mu,
                     const float* __restrict__ lv, const float* __restrict__ eps,
                     const uint64_t* __restrict__ seed_state, float* __restrict__ eps_out,
                     const float* __restrict__ w0, const float* __restrict__ b0,
                     const float* __restrict__ w2, const float* __restrict__ b2,
                     const float* __restrict__ w4, const float* __restrict__ b4,
                     float* __restrict__ item_feat, float* __restrict__ table,
                     float* __restrict__ hidden /*[2 layers][2 rows][H]*/, double* __restrict__ item_term,
                     bool staged) {
  __shared__ float s_h[2][2][kMaxHidden];
  __shared__ double s_red[kParamThreads / 32];
  const int t = threadIdx.x;
  if (blockIdx.x == 0) {
    // The H x H and 2D x H weights are staged into shared memory with ONE round of independent,
    // coalesced loads per thread (rows padded to H + 1 words: conflict-free both row- and
    // column-wise); dot products then run at shared-memory latency instead of one dependent
    // global load per term.  `staged` is false only for hidden sizes that do not fit.
    extern __shared__ float s_w[];
    const int ld = staged ? H + 1 : H;
    float* s_w4 = s_w + (size_t)H * (H + 1);
    if (staged) {
      stage_rows(s_w, w2, H, H, ld);
      stage_rows(s_w4, w4, 2 * D, H, ld);
    }
    const float* W2 = staged ? s_w : w2;
    const float* W4 = staged ? s_w4 : w4;
    for (int k = t; k < 2 * H; k += blockDim.x) {
      const int r = k / H, h = k % H;
      s_h[0][r][h] = elu(fmaf(w0[h], (float)r, b0[h]));  // cell input is r itself: 0 or 1
    }
    __syncthreads();
    for (int k = t; k < 2 * H; k += blockDim.x) {
      const int r = k / H, h = k % H;
      float a = b2[h];
      for (int j = 0; j < H; ++j) a = fmaf(W2[(size_t)h * ld + j], s_h[0][r][j], a);
      s_h[1][r][h] = elu(a);
    }
    __syncthreads();
    for (int k = t; k < 2 * 2 * D; k += blockDim.x) {
      const int r = k / (2 * D), o = k % (2 * D);
      float a = b4[o];
      for (int j = 0; j < H; ++j) a = fmaf(W4[(size_t)o * ld + j], s_h[1][r][j], a);
      table[r * 2 * D + o] = a;
    }
    for (int k = t; k < 2 * 2 * H; k += blockDim.x) hidden[k] = (&s_h[0][0][0])[(k / H) * kMaxHidden + (k % H)];
    return;
  }
  double acc = 0.0;
  const uint64_t key = seed_state != nullptr ? seed_state[0] + seed_state[1] : 0;
  for (int k = t; k < I * F; k += blockDim.x) {
    const float m = mu[k], l = lv[k];
    float e;
    if (eps != nullptr) {
      e = eps[k];
    } else {
      // ONE global item draw per step (models.py:361), identical on every rank: Philox(seed + step)
      // on a counter range disjoint from the persons' (same stream as vibo_philox_normal with
      // person_offset = kItemNoiseStream, ability_dim = 1)
      float nrm[4];
      philox_normal4(key, kItemNoiseStream + (uint64_t)k, 0u, nrm);
      e = nrm[0];
      eps_out[k] = e;
    }
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

__device__ __forceinline__ void
param_backward_body(int I, int F, int D, int H, int form, const float* __restrict__ mu,
                    const float* __restrict__ lv, const float* __restrict__ eps,
                    const float* __restrict__ w2, const float* __restrict__ w4,
                    const float* __restrict__ hidden, const float* __restrict__ g_table,
                    const float* __restrict__ g_item, const float gt,
                    float* __restrict__ g_mu, float* __restrict__ g_lv, float* __restrict__ g_w0,
                    float* __restrict__ g_b0, float* __restrict__ g_w2, float* __restrict__ g_b2,
                    float* __restrict__ g_w4, float* __restrict__ g_b4, bool staged) {
  __shared__ float s_h[2][2][kMaxHidden];   // activations h1, h2
  __shared__ float s_ga[2][2][kMaxHidden];  // gradients w.r.t. pre-activations a1, a2
  __shared__ float s_go[2][2 * VIBO_MAX_ABILITY_DIM];
  const int t = threadIdx.x;
  if (blockIdx.x == 0) {
    extern __shared__ float s_w[];  // staged w2 [H][H+1] | w4 [2D][H+1] (see param_forward_kernel)
    const int ld = staged ? H + 1 : H;
    float* s_w4 = s_w + (size_t)H * (H + 1);
    if (staged) {
      stage_rows(s_w, w2, H, H, ld);
      stage_rows(s_w4, w4, 2 * D, H, ld);
    }
    const float* W2 = staged ? s_w : w2;
    const float* W4 = staged ? s_w4 : w4;
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
      for (int o = 0; o < 2 * D; ++o) g = fmaf(W4[(size_t)o * ld + j], s_go[r][o], g);
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
      for (int h = 0; h < H; ++h) g = fmaf(W2[(size_t)h * ld + j], s_ga[1][r][h], g);
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
  // gt = d loss / d item_term
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

__global__ void __launch_bounds__(kParamThreads)
param_backward_kernel(int I, int F, int D, int H, int form, const float* __restrict__ mu,
                      const float* __restrict__ lv, const float* __restrict__ eps,
                      const float* __restrict__ w2, const float* __restrict__ w4,
                      const float* __restrict__ hidden, const float* __restrict__ g_table,
                      const float* __restrict__ g_item, const float* __restrict__ g_term_ptr,
                      float* __restrict__ g_mu, float* __restrict__ g_lv, float* __restrict__ g_w0,
                      float* __restrict__ g_b0, float* __restrict__ g_w2, float* __restrict__ g_b2,
                      float* __restrict__ g_w4, float* __restrict__ g_b4, bool staged) {
  param_backward_body(I, F, D, H, form, mu, lv, eps, w2, w4, hidden, g_table, g_item, *g_term_ptr, g_mu, g_lv,
                      g_w0, g_b0, g_w2, g_b2, g_w4, g_b4, staged);
}

// Tail of a fused step (vibo_step_tail): loss assembly, the parameter chain rule, and the step
// counters, in ONE launch:
//   loss = -LL + beta KL_theta + item_scale beta KL_item            (VIBO_ELBO_KL,  models.py:428-430)
//   loss = -LL - person_term + item_scale item_term                 (VIBO_ELBO_SAMPLE, :433-441)
// (item_scale = 1 / world_size in a person-sharded run), written to out[0]; the gradients of
// every parameter go to the caller's pointers (normally views of out[1:], the flat buffer the
// all-reduce and Adam work on).  counters[0..n) are incremented by one by block 0 AFTER the
// kernels that read them in this step (stream order): {seed, step}[1], the Adam step count.
__global__ void __launch_bounds__(kParamThreads)
step_tail_kernel(int I, int F, int D, int H, int form, float beta, float item_scale,
                 const double* __restrict__ scalars, const double* __restrict__ item_term,
                 float* __restrict__ loss_out, int64_t* counter0, int64_t* counter1, bool grad,
                 const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                 const float* __restrict__ w2, const float* __restrict__ w4,
                 const float* __restrict__ hidden, const float* __restrict__ g_table,
                 const float* __restrict__ g_item, float* __restrict__ g_mu, float* __restrict__ g_lv,
                 float* __restrict__ g_w0, float* __restrict__ g_b0, float* __restrict__ g_w2,
                 float* __restrict__ g_b2, float* __restrict__ g_w4, float* __restrict__ g_b4, bool staged) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double ll = scalars[0], term = scalars[1], it = *item_term;
    const double loss = form == VIBO_ELBO_KL ? -ll + (double)beta * term + (double)item_scale * (double)beta * it
                                             : -ll - term + (double)item_scale * it;
    *loss_out = (float)loss;
    if (counter0 != nullptr) *counter0 += 1;
    if (counter1 != nullptr) *counter1 += 1;
  }
  if (!grad) return;
  const float gt = form == VIBO_ELBO_KL ? item_scale * beta : item_scale;
  param_backward_body(I, F, D, H, form, mu, lv, eps, w2, w4, hidden, g_table, g_item, gt, g_mu, g_lv, g_w0, g_b0,
                      g_w2, g_b2, g_w4, g_b4, staged);
}

// torch.optim.Adam (vibo.py:221: defaults betas (0.9, 0.999), eps 1e-8, no weight decay, no
// amsgrad) over flat buffers; `step` is the 1-based count of THIS update, read from device memory
// (incremented by step_tail_kernel), so the launch is identical on every CUDA-graph replay.
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(int n, float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
            float* __restrict__ v, const int64_t* __restrict__ step, float lr, float b1, float b2, float eps) {
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) adam_bias_corrections(step, b1, b2, &s_bc[0], &s_bc[1]);
  __syncthreads();
  const float bc2_sqrt = s_bc[1];
  const float step_size = lr / s_bc[0];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    adam_update(grad[k], param + k, m + k, v + k, step_size, bc2_sqrt, b1, b2, eps);
}

// Dynamic shared memory of the staged weights; 0 (not staged) when they do not fit.
static size_t staged_bytes(int H, int D) {
  const size_t b = ((size_t)H * (H + 1) + (size_t)2 * D * (H + 1)) * sizeof(float);
  return b <= 160 * 1024 ? b : 0;
}
template <typename K>
static cudaError_t opt_in_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_param_forward(int I, int F, int D, int H, int form, const float* mu, const float* lv,
                                 const float* eps, const float* w0, const float* b0, const float* w2,
                                 const float* b2, const float* w4, const float* b4, float* item_feat,
                                 float* table, float* hidden, double* item_term, cudaStream_t st,
                                 const uint64_t* seed_state, float* eps_out) {
  if (H > kMaxHidden) return cudaErrorInvalidValue;
  if (eps == nullptr && (seed_state == nullptr || eps_out == nullptr)) return cudaErrorInvalidValue;
  const size_t sb = staged_bytes(H, D);
  cudaError_t e0 = opt_in_smem(param_forward_kernel, sb);
  if (e0 != cudaSuccess) return e0;
  param_forward_kernel<<<2, kParamThreads, sb, st>>>(I, F, D, H, form, mu, lv, eps, seed_state, eps_out, w0, b0,
                                                     w2, b2, w4, b4, item_feat, table, hidden, item_term, sb > 0);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_param_backward(int I, int F, int D, int H, int form, const float* mu,
                                  const float* lv, const float* eps, const float* w2, const float* w4,
                                  const float* hidden, const float* g_table, const float* g_item,
                                  const float* g_term, float* g_mu, float* g_lv, float* g_w0, float* g_b0,
                                  float* g_w2, float* g_b2, float* g_w4, float* g_b4, cudaStream_t st) {
  if (H > kMaxHidden) return cudaErrorInvalidValue;
  const size_t sb = staged_bytes(H, D);
  cudaError_t e0 = opt_in_smem(param_backward_kernel, sb);
  if (e0 != cudaSuccess) return e0;
  param_backward_kernel<<<2, kParamThreads, sb, st>>>(I, F, D, H, form, mu, lv, eps, w2, w4, hidden,
                                                      g_table, g_item, g_term, g_mu, g_lv, g_w0, g_b0, g_w2,
                                                      g_b2, g_w4, g_b4, sb > 0);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_step_tail(int I, int F, int D, int H, int form, float beta, float item_scale,
                             const double* scalars, const double* item_term, float* loss_out, int64_t* counter0,
                             int64_t* counter1, bool grad, const float* mu, const float* lv, const float* eps,
                             const float* w2, const float* w4, const float* hidden, const float* g_table,
                             const float* g_item, float* g_mu, float* g_lv, float* g_w0, float* g_b0,
                             float* g_w2, float* g_b2, float* g_w4, float* g_b4, cudaStream_t st) {
  if (H > kMaxHidden) return cudaErrorInvalidValue;
  const size_t sb = grad ? staged_bytes(H, D) : 0;
  cudaError_t e0 = opt_in_smem(step_tail_kernel, sb);
  if (e0 != cudaSuccess) return e0;
  step_tail_kernel<<<grad ? 2 : 1, kParamThreads, sb, st>>>(I, F, D, H, form, beta, item_scale, scalars, item_term,
                                                             loss_out, counter0, counter1, grad, mu, lv, eps, w2,
                                                             w4, hidden, g_table, g_item, g_mu, g_lv, g_w0, g_b0,
                                                             g_w2, g_b2, g_w4, g_b4, sb > 0);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_adam(int n, float* param, const float* grad, float* m, float* v, const int64_t* step, float lr,
                        float b1, float b2, float eps, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  int blocks = (n + 255) / 256;
  if (blocks > 296) blocks = 296;
  adam_kernel<<<blocks, 256, 0, st>>>(n, param, grad, m, v, step, lr, b1, b2, eps);
  note_launch();
  return cudaGetLastError();
}

}  // namespace vibo
