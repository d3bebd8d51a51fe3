// Shared device helpers for the VIBO ELBO kernels (sm_100a).
//
// Math follows SURVEY.md Appendix A, which restates the reference's
// src/torch_core/models.py:337-443, :551-766 and src/utils.py:46-113.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vibo {

// torch.finfo(float32).eps: clamp of Bernoulli probabilities inside
// torch.distributions (reached from src/utils.py:46-49).
__device__ constexpr float kEps32 = 1.1920929e-07f;
// log((1 - eps32) / eps32): |logit| at which a 1PL/2PL cell hits the clamp.
__device__ constexpr float kLogitClamp = 15.942385f;
// product_of_experts eps, src/utils.py:105.
__device__ constexpr float kPoeEps = 1e-8f;
__device__ constexpr float kHalfLog2Pi = 0.9189385332046727f;

__host__ __device__ constexpr int item_width(int model, int D) {
  return model == 1 ? 1 : (model == 2 ? D + 1 : D + 2);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Result of one observed cell of the link + Bernoulli log-likelihood.
struct CellGrad {
  float ll;    // log Bernoulli(x; p~)
  float dz;    // d ll / d z
  float dgam;  // d ll / d gamma (3PL guess logit), 0 otherwise
};

// 1PL / 2PL cell: p = sigmoid(z); torch clamps p to [eps32, 1 - eps32]
// (utils.py:46-49 -> Bernoulli(probs).log_prob), which in logit space is
// ll = -softplus(-t), t = clamp((2x-1) z, +-kLogitClamp), with zero gradient
// outside the clamp.  PRECISE selects libm-accurate exp/log1p; otherwise the
// MUFU approximations (ex2 / lg2 / rcp) are used.
template <bool PRECISE>
__device__ __forceinline__ CellGrad cell_logistic(float z, bool x1) {
  const float u = x1 ? z : -z;
  const float t = fminf(fmaxf(u, -kLogitClamp), kLogitClamp);
  CellGrad c;
  float e, w, L;
  if (PRECISE) {
    e = expf(-fabsf(t));
    w = 1.0f + e;
    L = log1pf(e);
  } else {
    e = __expf(-fabsf(t));
    w = 1.0f + e;
    L = __logf(w);
  }
  c.ll = -(fmaxf(-t, 0.0f) + L);
  const float sig_neg = (t >= 0.0f ? e : 1.0f) * (PRECISE ? 1.0f / w : __frcp_rn(w));  // sigmoid(-t)
  const float du = (t == u) ? sig_neg : 0.0f;
  c.dz = x1 ? du : -du;
  c.dgam = 0.0f;
  return c;
}

// 3PL cell: p = g + (1-g) sigmoid(z)  (models.py:753-766).  1 - p is formed as
// (1-g) sigmoid(-z) so it keeps full relative precision near p -> 1.
template <bool PRECISE>
__device__ __forceinline__ CellGrad cell_3pl(float z, float g, bool x1) {
  const float e = PRECISE ? expf(-fabsf(z)) : __expf(-fabsf(z));
  const float r = 1.0f / (1.0f + e);
  const float s = z >= 0.0f ? r : e * r;       // sigmoid(z)
  const float sn = z >= 0.0f ? e * r : r;      // sigmoid(-z)
  const float p = fmaf(1.0f - g, s, g);
  const float q = (1.0f - g) * sn;             // 1 - p
  const bool inside = (p >= kEps32) && (q >= kEps32);
  const float pc = fminf(fmaxf(p, kEps32), 1.0f - kEps32);
  const float qc = fminf(fmaxf(q, kEps32), 1.0f - kEps32);
  CellGrad c;
  float dp;
  if (x1) {
    c.ll = PRECISE ? logf(pc) : __logf(pc);
    dp = 1.0f / pc;
  } else {
    c.ll = PRECISE ? logf(qc) : __logf(qc);
    dp = -1.0f / qc;
  }
  if (!inside) dp = 0.0f;
  c.dz = dp * (1.0f - g) * s * sn;
  c.dgam = dp * sn * g * (1.0f - g);
  return c;
}

// ---------------------------------------------------------------------------
// Counter-based noise for the reparameterised ability draw (models.py:506-510
// uses randn_like; here the draw is keyed by the GLOBAL person index so a
// person's noise does not depend on how persons are sharded over GPUs).
// Philox4x32-10 (Salmon et al. 2011): counter = (person_lo, person_hi, block, 0),
// key = (seed_lo, seed_hi); block b yields the normals for dims 4b .. 4b+3
// through Box-Muller.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// torch.optim.Adam (amsgrad off, no weight decay), one element; bc1 = 1 - b1^t, bc2_sqrt = sqrt(1 - b2^t).
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / bc1 * m / (sqrt(v) / bc2_sqrt + eps)
__device__ __forceinline__ void adam_update(float g, float* __restrict__ param, float* __restrict__ m,
                                            float* __restrict__ v, float step_size, float bc2_sqrt, float b1,
                                            float b2, float eps) {
  const float mk = fmaf(b1, *m, (1.0f - b1) * g);
  const float vk = fmaf(b2, *v, (1.0f - b2) * g * g);
  *m = mk;
  *v = vk;
  const float denom = sqrtf(vk) / bc2_sqrt + eps;
  *param -= step_size * (mk / denom);
}
__device__ __forceinline__ void adam_bias_corrections(const int64_t* step, float b1, float b2, float* bc1,
                                                      float* bc2_sqrt) {
  const double t = (double)*step;
  *bc1 = (float)(1.0 - pow((double)b1, t));
  *bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
}

// Four standard normals for (seed, person, block): Box-Muller on the four Philox words with the
// MUFU approximations (lg2 / sin / cos; absolute error ~1e-6, irrelevant for noise).  The angle is
// taken in [-pi, pi] where sin.approx / cos.approx are most accurate:
// (cos, sin)(2 pi u) = -(cos, sin)(pi (2u - 1)).
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t person, uint32_t block,
                                               float out[4]) {
  uint32_t c[4] = {(uint32_t)person, (uint32_t)(person >> 32), block, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
    const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;       // [0, 1]
    const float rad = sqrtf(-2.0f * __logf(u1));
    const float ang = 3.14159265358979f * fmaf(2.0f, u2, -1.0f);
    out[2 * h] = -rad * __cosf(ang);
    out[2 * h + 1] = -rad * __sinf(ang);
  }
}

}  // namespace vibo
