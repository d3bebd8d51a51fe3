// Single-pass fused ELBO kernel for the unconditional encoder (sm_100a).
//
// One pass over the (P, I) response + mask rows computes, per person,
//   counts -> product-of-experts posterior (models.py:596-629, utils.py:105-113)
//   theta = mu + sd * eps                 (models.py:506-510)
//   z_ij, masked Bernoulli log-lik       (models.py:729-766, utils.py:46-49)
//   KL(q||N(0,1)) or log p - log q       (utils.py:85-88 / models.py:433-435)
// and, when GRAD, d loss_k / d item_feat and d loss_k / d table in the same
// read of the rows (SURVEY.md Appendix A "Backward").
//
// Data movement: chunks of R rows (R % 16 == 0, so the float32 response block,
// the uint8 mask block and the eps block all start 16-byte aligned whatever I
// is) stream into a ring of shared-memory stages with 1-D TMA bulk copies
// (cp.async.bulk, SASS UBLKCP) completing on mbarriers.  There is no dedicated
// producer warp (a 17th warp would cap registers at 96/thread): the warp that
// is LAST to finish a stage re-arms that stage's barrier and issues the copy
// of the chunk that will next occupy it, so the refill starts the moment the
// slot is free.  The warps read their rows with conflict-free 128-bit LDS;
// each row is read from HBM exactly once.
//
// Work layout: a sub-group of LPP lanes owns one person (32/LPP persons per
// warp at a time).  Lane q of the sub-group owns the 4-item groups
// g = q + LPP*k (k < NG), i.e. items 4g..4g+3, for EVERY person the warp
// processes, so item parameters come from shared memory as float4 and the
// per-item gradient accumulators stay in registers for the whole kernel: the
// cross-person reductions need no atomics and the result is deterministic.
#pragma once

#include "vibo_common.cuh"
#include "vibo_kernels.h"

namespace vibo {

constexpr int kFusedConsumerWarps = 16;
constexpr int kFusedThreads = kFusedConsumerWarps * 32;
// The warps of a CTA form kFusedTeams independent teams; each team streams its
// own interleaved sequence of row chunks through its own ring of stages, so a
// CTA keeps several smaller bulk copies in flight at staggered times instead
// of one large one.
constexpr int kFusedTeamWarps = 4;
constexpr int kFusedTeams = kFusedConsumerWarps / kFusedTeamWarps;

struct FusedParams {
  int64_t P;
  int I;
  int R;        // rows per stage
  int nstage;   // ring depth per team
  int form;     // VIBO_ELBO_*
  int missing_policy;
  float beta;
  int debug;    // 0 normal; 1 stream only (no math); 2 math only (no refills)  [VIBO_FUSED_DEBUG]
  int mask_off, eps_off, stage_bytes;   // fused_smem_layout of this launch (host-computed copies: one constant load each)
  const float* resp;
  const uint8_t* mask;
  const float* eps;        // (P, D) standard normals the rows are staged from
  float* eps_draw;         // non-null: == eps, a scratch this kernel FILLS first (fused_draw_noise)
  uint64_t seed;           // Philox key when seed_dev is null
  const uint64_t* seed_dev;  // device {seed, step}: key = seed + step, read at run time (CUDA-graph replays)
  int64_t person_offset;   // global index of row 0
  const float* item_feat;  // (I, F)
  const float* table;      // (2, 1, 2D)
  float* out_mu;           // (P, D) or null
  float* out_lv;
  float* out_theta;
  double* part_scalar;     // [grid][2]
  float* part_table;       // [grid][4D]  (A0 | A1 | B0 | B1)
  float* part_item;        // [grid][I*F]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar_addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_addr(smem_u32(bar), parity); }
// 1-D TMA bulk copy global -> shared, completing `bytes` on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// Shared-memory loads by 32-bit shared-window address (+ immediate offset):
// keeps one base register per operand instead of letting the compiler
// re-derive the generic->shared conversion for every group.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kFusedHdr = 512;
__host__ __device__ inline size_t fused_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Shared-memory layout (bytes): barriers | item params (SoA floats) | stages.
struct FusedSmem {
  size_t params_off, stage_off, stage_bytes, resp_bytes, mask_off, eps_off, total;
};
__host__ __device__ inline FusedSmem fused_smem_layout(int I, int D, int model, int R, int nstage,
                                                       int nteams = kFusedTeams, int scratch_bytes = 0) {
  FusedSmem L;
  const int nparam = model == 1 ? 1 : (model == 2 ? D + 1 : D + 2);
  // header: [0,128) up to 16 'full' mbarriers; [128,192) up to 16 stage counters; [256,384) up to 16
  // 'empty' mbarriers; scratch from kFusedHdr
  L.params_off = kFusedHdr + scratch_bytes;
  L.stage_off = fused_align(L.params_off + (size_t)nparam * I * 4, 128);
  L.resp_bytes = (size_t)R * I * 4;
  L.mask_off = L.resp_bytes;
  L.eps_off = fused_align(L.mask_off + (size_t)R * I, 16);
  L.stage_bytes = fused_align(L.eps_off + (size_t)R * D * 4, 128);
  L.total = L.stage_off + (size_t)nstage * nteams * L.stage_bytes;
  return L;
}

// Fill stage memory `st` with chunk c (rows [c*R, c*R + rows)) and make
// `bar` complete when the data has landed.  Called by one whole warp.
template <int D>
__device__ __forceinline__ void fused_issue_chunk(const FusedParams& p, const FusedSmem& L, int64_t c,
                                                  unsigned char* st, uint64_t* bar, int lane) {
  const int I = p.I, R = p.R;
  const int64_t row0 = c * R;
  const int rows = (int)((p.P - row0 < R) ? p.P - row0 : R);
  const uint32_t b_resp = (uint32_t)rows * I * 4, b_mask = (uint32_t)rows * I, b_eps = (uint32_t)rows * D * 4;
  if (((b_mask | b_eps) & 15u) == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar, b_resp + b_mask + b_eps);
      bulk_g2s(st, p.resp + row0 * I, b_resp, bar);
      bulk_g2s(st + L.mask_off, p.mask + row0 * I, b_mask, bar);
      bulk_g2s(st + L.eps_off, p.eps + row0 * D, b_eps, bar);
    }
  } else {
    // ragged tail chunk: sizes are not 16-byte multiples, copy by hand
    const float* gr = p.resp + row0 * I;
    float* sr = reinterpret_cast<float*>(st);
    for (int k = lane; k < rows * I; k += 32) sr[k] = gr[k];
    const uint8_t* gm = p.mask + row0 * I;
    uint8_t* sm = st + L.mask_off;
    for (int k = lane; k < rows * I; k += 32) sm[k] = gm[k];
    const float* ge = p.eps + row0 * D;
    float* se = reinterpret_cast<float*>(st + L.eps_off);
    for (int k = lane; k < rows * D; k += 32) se[k] = ge[k];
    // the hand-copied cells are performed before the arrive that publishes them (an mbarrier arrive is not
    // ordered behind the warp's own in-flight shared-memory accesses on B200)
    __threadfence_block();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  }
}

// In-kernel reparameterisation noise (models.py:506-510 randn_like): before its first bulk copy a
// CTA fills the scratch rows IT will stage -- chunks (blockIdx.x * NQ + t) + k * gridDim.x * NQ of
// its NQ teams -- with Philox normals keyed by the GLOBAL person index, all threads in parallel
// (one person per thread at a time), and then reads them back through the same TMA path as
// caller-supplied noise.  No separate launch; the 4 D bytes per person stay in L2.  Generic-proxy
// stores followed by async-proxy (bulk copy) loads: every writer issues fence.proxy.async before
// the CTA barrier that precedes the first copy.
template <int D>
__device__ __forceinline__ void fused_draw_noise(const FusedParams& p, int NQ) {
  if (p.eps_draw == nullptr) return;
  const uint64_t key = p.seed_dev != nullptr ? p.seed_dev[0] + p.seed_dev[1] : p.seed;
  const int R = p.R;
  const int64_t n_chunks = (p.P + R - 1) / R;
  const int per_round = NQ * R;  // rows of one round of the CTA's teams
  for (int64_t u = threadIdx.x;; u += blockDim.x) {
    const int64_t k = u / per_round;
    const int rem = (int)(u - k * per_round);
    const int64_t c0 = (int64_t)blockIdx.x * NQ + k * (int64_t)gridDim.x * NQ;
    if (c0 >= n_chunks) break;
    const int64_t c = c0 + rem / R;
    const int64_t row = c * R + rem % R;
    if (c >= n_chunks || row >= p.P) continue;
    float nrm[4];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if ((d & 3) == 0) philox_normal4(key, (uint64_t)(p.person_offset + row), (uint32_t)(d >> 2), nrm);
      p.eps_draw[row * D + d] = nrm[d & 3];
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");
}

// One group of 4 consecutive items of one person: link, Bernoulli
// log-likelihood and (GRAD) the per-cell gradient terms.
//
// 1PL/2PL, with zc = clamp(z, +-kLogitClamp) (the eps32 clamp of
// torch.distributions in logit space) and e = exp(-|zc|):
//   ll = x zc - softplus(zc) = (x - 1/2) zc - |zc|/2 - log(1 + e)
//   d ll / d z = x - sigmoid(zc) = (x - 1/2) - copysign(1/(1+e) - 1/2, zc),  0 where z != zc
// accumulated as s1 += (x-1/2) zc, s2 += |zc|, s3 += log2(1+e).
// FULL: every cell of the row is observed (mask not consulted).  Otherwise a
// missing cell is turned into the neutral cell (x = 1/2, z = 0), whose only
// contribution -- log2(2) = 1 to s3 -- is subtracted via the caller's count.
template <int MODEL, int D, bool GRAD, bool FULL>
__device__ __forceinline__ void fused_group(const float4& x4, uint32_t m4, const float4& b4,
                                            const float4 (&a4)[MODEL == 1 ? 1 : D], const float4& g4,
                                            const float (&th)[D], float tsum, float (&gth)[D],
                                            float* __restrict__ acc, float& s1, float& s2, float& s3) {
  constexpr int F = item_width(MODEL, D);
  constexpr int DA = MODEL == 1 ? 0 : D;
  constexpr float kNegLog2e = -1.4426950408889634f;
  const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
  const float bs[4] = {b4.x, b4.y, b4.z, b4.w};
  const float gs[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float av[DA > 0 ? DA : 1];
#pragma unroll
    for (int d = 0; d < DA; ++d) av[d] = c == 0 ? a4[d].x : (c == 1 ? a4[d].y : (c == 2 ? a4[d].z : a4[d].w));
    float z = bs[c];
    if (MODEL == 1) {
      z += tsum;
    } else {
#pragma unroll
      for (int d = 0; d < DA; ++d) z = fmaf(-th[d], av[d], z);
    }
    const bool o = FULL ? true : ((m4 >> (8 * c)) & 0xffu) != 0;
    float dz = 0.0f, dgam = 0.0f;
    if (MODEL == 3) {
      const CellGrad cg = cell_3pl<false>(z, gs[c], xs[c] > 0.5f);
      s1 += o ? cg.ll : 0.0f;
      dz = o ? cg.dz : 0.0f;
      dgam = o ? cg.dgam : 0.0f;
    } else {
      float xm = xs[c] - 0.5f;
      if (!FULL) {
        xm = o ? xm : 0.0f;
        z = o ? z : 0.0f;
      }
      const float zc = fminf(fmaxf(z, -kLogitClamp), kLogitClamp);
      const float e = ex2_approx(fabsf(zc) * kNegLog2e);
      const float w = 1.0f + e;
      s1 = fmaf(xm, zc, s1);
      s2 += fabsf(zc);
      s3 += lg2_approx(w);
      if (GRAD) {
        const float h = rcp_approx(w) - 0.5f;                    // sigmoid(|zc|) - 1/2
        const float cs = __uint_as_float((__float_as_uint(h) & 0x7fffffffu) | (__float_as_uint(zc) & 0x80000000u));
        dz = (z == zc) ? xm - cs : 0.0f;                         // zero outside the eps32 clamp
      }
    }
    if (GRAD) {
      if (MODEL == 1) {
        gth[0] -= dz;  // d loss_k / d theta_d = -sum_j dz (same for every d)
        acc[c * F] += dz;
      } else {
#pragma unroll
        for (int d = 0; d < DA; ++d) {
          gth[d] = fmaf(dz, av[d], gth[d]);
          acc[c * F + d] = fmaf(dz, th[d], acc[c * F + d]);
        }
        acc[c * F + D] += dz;
        if (MODEL == 3) acc[c * F + D + 1] += dgam;
      }
    }
  }
}

// Pass 2 over one person's row.  Lane q owns groups q + LPP*k: the first
// `kfull` of them are valid for every lane (warp-uniform test, no divergence);
// at most one further group is valid for lanes q < ntail only, and it always
// accumulates into the last register slot (slot NG-1 is free whenever a
// ragged group exists).
template <int MODEL, int D, int LPP, int NG, bool GRAD, bool FULL>
__device__ __forceinline__ void fused_pass2(uint32_t xp, uint32_t mp, uint32_t pp, int I4, int kfull, bool has_tail,
                                            const float (&th)[D], float tsum, float (&gth)[D],
                                            float (&acc)[GRAD ? NG * 4 * item_width(MODEL, D) : 1],
                                            float& s1, float& s2, float& s3) {
  constexpr int F = item_width(MODEL, D);
  constexpr int DA = MODEL == 1 ? 0 : D;
  float nmiss_lane = 0.0f;
  // xp / mp / pp: shared-window byte addresses of this lane's first group in the
  // response row, the mask row and the item-parameter arrays (I4 float4 each)
  auto one = [&](int gi, int slot) {
    const float4 x4 = lds128(xp + gi * 16);
    uint32_t m4 = 0x01010101u;
    if (!FULL) {
      m4 = lds32(mp + gi * 4);
      if (MODEL != 3) nmiss_lane += (float)(4 - __popc(m4 & 0x01010101u));
    }
    float4 a4[DA > 0 ? DA : 1];
#pragma unroll
    for (int d = 0; d < DA; ++d) a4[d] = lds128(pp + (d * I4 + gi) * 16);
    const float4 b4 = lds128(pp + (DA * I4 + gi) * 16);
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODEL == 3) g4 = lds128(pp + ((DA + 1) * I4 + gi) * 16);
    fused_group<MODEL, D, GRAD, FULL>(x4, m4, b4, a4, g4, th, tsum, gth, GRAD ? &acc[slot * 4 * F] : &acc[0],
                                      s1, s2, s3);
  };
#pragma unroll
  for (int k = 0; k < NG; ++k)
    if (k < kfull) one(LPP * k, k);
  if (has_tail) one(LPP * kfull, NG - 1);
  if (!FULL && MODEL != 3) s3 -= nmiss_lane;  // neutral cells added log2(2) each
}

template <int MODEL, int D, int LPP, int NG, bool GRAD>
__global__ void __launch_bounds__(kFusedThreads, 1) fused_uncond_kernel(const __grid_constant__ FusedParams p) {
  constexpr int F = item_width(MODEL, D);
  constexpr int DA = MODEL == 1 ? 0 : D;  // discrimination arrays
  (void)DA;
  constexpr int NW = kFusedConsumerWarps;
  constexpr int PPW = 32 / LPP;           // persons per warp at a time
  constexpr int IPL = NG * 4;             // items per lane
  constexpr float kLn2 = 0.6931471805599453f;

  extern __shared__ __align__(128) unsigned char smem[];
  const int I = p.I, R = p.R, NS = p.nstage;
  const FusedSmem L = fused_smem_layout(I, D, MODEL, R, NS);
  constexpr int TW = kFusedTeamWarps, NQ = kFusedTeams;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);     // [NQ * NS]
  int* done_cnt = reinterpret_cast<int*>(smem + 128);          // warps finished with each stage
  uint64_t* empty_bar = reinterpret_cast<uint64_t*>(smem + 256);   // [NQ * NS]: every warp of a team arrives when it leaves the stage
  float* s_param = reinterpret_cast<float*>(smem + L.params_off);  // [a_0 | .. | a_{D-1} | b | guess]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / TW, wt = warp % TW;
  const int n_groups = I >> 2;
  const int64_t n_chunks = (p.P + R - 1) / R;
  // team t of CTA b walks chunks (b*NQ + t) + k * (gridDim.x * NQ)
  const int64_t chunk0 = (int64_t)blockIdx.x * NQ + team, chunk_step = (int64_t)gridDim.x * NQ;
  uint64_t* t_full = full_bar + team * NS;
  int* t_done = done_cnt + team * NS;
  uint64_t* t_empty = empty_bar + team * NS;
  unsigned char* t_stage = smem + L.stage_off + (size_t)team * NS * L.stage_bytes;

  // ---- one-time setup ----------------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < NQ * NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], TW);
      done_cnt[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = threadIdx.x; j < I; j += blockDim.x) {
    if (MODEL == 1) {
      s_param[j] = p.item_feat[j];
    } else {
      for (int d = 0; d < D; ++d) s_param[(size_t)d * I + j] = p.item_feat[(size_t)j * F + d];
      s_param[(size_t)D * I + j] = p.item_feat[(size_t)j * F + D];
      if (MODEL == 3)
        s_param[(size_t)(D + 1) * I + j] = 1.0f / (1.0f + expf(-p.item_feat[(size_t)j * F + D + 1]));
    }
  }
  fused_draw_noise<D>(p, NQ);
  __syncthreads();

  // per-CTA results
  // per-lane float accumulators (a few hundred rows each; the cross-lane and
  // cross-CTA sums are done in double)
  float ll_acc = 0.0f, term_acc = 0.0f;
  float tA[2][D], tB[2][D];
  float acc[GRAD ? IPL * F : 1];
#pragma unroll
  for (int d = 0; d < D; ++d) tA[0][d] = tA[1][d] = tB[0][d] = tB[1][d] = 0.0f;
#pragma unroll
  for (int k = 0; k < (GRAD ? IPL * F : 1); ++k) acc[k] = 0.0f;

  // prologue: the first NS chunks of each team
  if (wt == 0) {
    for (int s = 0; s < NS; ++s) {
      const int64_t c = chunk0 + (int64_t)s * chunk_step;
      if (c < n_chunks) fused_issue_chunk<D>(p, L, c, t_stage + (size_t)s * L.stage_bytes, &t_full[s], lane);
    }
  }
  {
    // ======================= consumer warps ==============================
    const int sub = lane / LPP, q = lane % LPP;
    // expert table -> precisions (utils.py:107-108); same for every person
    float tau[2][D], mt[2][D];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float mu = p.table[r * 2 * D + d], lam = p.table[r * 2 * D + D + d];
        tau[r][d] = 1.0f / (expf(lam) + kPoeEps);
        mt[r][d] = mu * tau[r][d];
      }
    const float prior_tau = p.missing_policy == VIBO_MISSING_PRIOR ? 1.0f / (1.0f + kPoeEps) : 0.0f;
    // lane q owns the 4-item groups q + LPP*k: k < kfull for every lane, plus
    // one ragged group (index kfull) for lanes q < ntail
    const int kfull = min(n_groups / LPP, NG);
    const bool has_tail = kfull < NG && q < n_groups - kfull * LPP;
    uint32_t pp = smem_u32(s_param) + q * 16;
    asm volatile("mov.u32 %0, %0;" : "+r"(pp));  // opaque: keep it in a register, do not re-derive per group

    // Loop state is kept incremental and 32-bit (no divisions, no 64-bit row
    // arithmetic per stage): ring slot s and its mbarrier phase, the shared
    // addresses of this sub-group's first row in a stage, and the rows left.
    const uint32_t stage0 = smem_u32(t_stage), bar0 = smem_u32(t_full);
    const int first_row = wt * PPW + sub;                       // row of this sub-group in a stage
    const uint32_t off_x = (uint32_t)first_row * I * 4 + q * 16;
    const uint32_t off_m = (uint32_t)L.mask_off + (uint32_t)first_row * I + q * 4;
    const uint32_t off_e = (uint32_t)L.eps_off + (uint32_t)first_row * D * 4;
    const uint32_t step_x = (uint32_t)TW * PPW * I * 4, step_m = (uint32_t)TW * PPW * I,
                   step_e = (uint32_t)TW * PPW * D * 4;
    int s = 0;
    uint32_t phase = 0;
    int64_t rows_left = p.P - chunk0 * R;                       // rows from this chunk to the end
    const int64_t rows_step = chunk_step * R;
    for (int64_t c = chunk0; c < n_chunks; c += chunk_step, rows_left -= rows_step) {
      if (p.debug != 2 || c == chunk0) mbar_wait_addr(bar0 + (uint32_t)s * 8, phase);
      const uint32_t sb = stage0 + (uint32_t)s * (uint32_t)L.stage_bytes;
      const int rows = rows_left < R ? (int)rows_left : R;
      const int64_t row0 = c * R;
      uint32_t xrow = sb + off_x, mrow = sb + off_m, erow = sb + off_e;
      for (int rbase = wt * PPW; rbase < rows && p.debug != 1;
           rbase += TW * PPW, xrow += step_x, mrow += step_m, erow += step_e) {
        const int r = rbase + sub;
        const bool valid = r < rows;
        // idle sub-groups of a ragged tail re-read the first row of the pass (in range)
        const uint32_t xp = valid ? xrow : xrow - (uint32_t)sub * I * 4;
        const uint32_t mp = valid ? mrow : mrow - (uint32_t)sub * I;
        const uint32_t ep = valid ? erow : erow - (uint32_t)sub * D * 4;

        // ---- pass 1: counts ------------------------------------------------
        float n1f = 0.0f;
        uint32_t mand = 0x01010101u;
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          if (k < kfull) {
            const float4 x = lds128(xp + LPP * k * 16);
            n1f += (x.x + x.y) + (x.z + x.w);
            mand &= lds32(mp + LPP * k * 4);
          }
        }
        if (has_tail) {
          const float4 x = lds128(xp + LPP * kfull * 16);
          n1f += (x.x + x.y) + (x.z + x.w);
          mand &= lds32(mp + LPP * kfull * 4);
        }
        const bool full_obs = __all_sync(0xffffffffu, mand == 0x01010101u);
        float nobsf = (float)I;
        if (!full_obs) {
          int n1 = 0, nobs = 0;
          auto count = [&](int gi) {
            const float4 x = lds128(xp + gi * 16);
            const uint32_t m = lds32(mp + gi * 4);
            const bool o0 = (m & 0xffu) != 0, o1 = (m & 0xff00u) != 0, o2 = (m & 0xff0000u) != 0,
                       o3 = (m & 0xff000000u) != 0;
            nobs += (int)o0 + (int)o1 + (int)o2 + (int)o3;
            n1 += (int)(o0 && x.x > 0.5f) + (int)(o1 && x.y > 0.5f) + (int)(o2 && x.z > 0.5f) +
                  (int)(o3 && x.w > 0.5f);
          };
#pragma unroll
          for (int k = 0; k < NG; ++k)
            if (k < kfull) count(LPP * k);
          if (has_tail) count(LPP * kfull);
          n1f = (float)n1;
          nobsf = (float)nobs;
        }
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) {
          n1f += __shfl_xor_sync(0xffffffffu, n1f, o);
          if (!full_obs) nobsf += __shfl_xor_sync(0xffffffffu, nobsf, o);
        }
        const float n0f = nobsf - n1f, nmiss = (float)I - nobsf;

        // ---- per-person posterior and draw ---------------------------------
        float amu[D], Ssum[D], sd[D], th[D], epsv[D], alv[D];
        float tsum = 0.0f;
        float term = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float S = fmaf(n0f, tau[0][d], fmaf(n1f, tau[1][d], nmiss * prior_tau));
          const float N = fmaf(n0f, mt[0][d], n1f * mt[1][d]);
          const float invS = __fdividef(1.0f, S);  // = exp(logvar)
          Ssum[d] = invS;
          amu[d] = N * invS;
          alv[d] = -kLn2 * lg2_approx(S);          // log(1 / S)
          sd[d] = rsqrtf(S);                       // exp(logvar / 2)
          epsv[d] = __uint_as_float(lds32(ep + d * 4));
          th[d] = fmaf(epsv[d], sd[d], amu[d]);
          tsum += th[d];
          if (p.form == VIBO_ELBO_KL) {
            term += -0.5f * (1.0f + alv[d] - amu[d] * amu[d] - invS);
          } else {
            // log p(theta) - log q(theta); theta - mu = sd * eps exactly
            term += -0.5f * th[d] * th[d] + 0.5f * epsv[d] * epsv[d] + 0.5f * alv[d];
          }
        }
        if (valid && q == 0) {
          term_acc += term;
          const int64_t row = row0 + r;
          if (p.out_mu != nullptr) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              p.out_mu[row * D + d] = amu[d];
              p.out_lv[row * D + d] = alv[d];
              p.out_theta[row * D + d] = th[d];
            }
          }
        }

        // ---- pass 2: link, log-likelihood, gradients -----------------------
        float gth[D];
#pragma unroll
        for (int d = 0; d < D; ++d) gth[d] = 0.0f;
        float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        if (!valid) {
          // idle sub-group of a ragged tail: contributes nothing (its lanes
          // still take part in the shuffles below)
        } else if (full_obs)
          fused_pass2<MODEL, D, LPP, NG, GRAD, true>(xp, mp, pp, n_groups, kfull, has_tail, th, tsum, gth, acc, s1, s2, s3);
        else
          fused_pass2<MODEL, D, LPP, NG, GRAD, false>(xp, mp, pp, n_groups, kfull, has_tail, th, tsum, gth, acc, s1, s2, s3);
        if (valid) ll_acc += MODEL == 3 ? s1 : s1 - 0.5f * s2 - kLn2 * s3;

        // ---- per-person backward -------------------------------------------
        if (GRAD) {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            float gv = MODEL == 1 ? gth[0] : gth[d];
#pragma unroll
            for (int o = LPP / 2; o > 0; o >>= 1) gv += __shfl_xor_sync(0xffffffffu, gv, o);
            float g_mu, g_lv;
            if (p.form == VIBO_ELBO_KL) {
              g_mu = fmaf(p.beta, amu[d], gv);
              g_lv = 0.5f * gv * epsv[d] * sd[d] + 0.5f * p.beta * (Ssum[d] - 1.0f);
            } else {
              gv += th[d];
              g_mu = gv;
              g_lv = 0.5f * gv * epsv[d] * sd[d] - 0.5f;
            }
            const float GN = g_mu * Ssum[d];  // Ssum holds 1 / S
            const float GS = -(g_mu * amu[d] + g_lv) * Ssum[d];
            if (valid && q == 0) {
              tA[0][d] = fmaf(n0f, GN, tA[0][d]);
              tA[1][d] = fmaf(n1f, GN, tA[1][d]);
              tB[0][d] = fmaf(n0f, GS, tB[0][d]);
              tB[1][d] = fmaf(n1f, GS, tB[1][d]);
            }
          }
        }
      }
      // the last warp of the team to leave the stage refills it with the chunk NS iterations ahead
      // (VIBO_FUSED_DEBUG=3 adds a team barrier first, see vibo_fused2_kernel.cuh)
      if (p.debug == 3) asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(kFusedTeamWarps * 32) : "memory");
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
      if (last && p.debug != 2) {
        const int64_t cn = c + (int64_t)NS * chunk_step;
        if (cn < n_chunks) {
          mbar_wait(&t_empty[s], phase);
          fused_issue_chunk<D>(p, L, cn, t_stage + (size_t)s * L.stage_bytes, &t_full[s], lane);
        }
      }
      if (p.debug != 2) {
        if (++s == NS) {
          s = 0;
          phase ^= 1u;
        }
      }
    }
  }

  // ---- CTA-level combine (deterministic order) -----------------------------
  __syncthreads();  // every stage consumed; stage memory is free for reuse
  double* s_d = reinterpret_cast<double*>(smem + L.stage_off);          // [NW+1][2]
  float* s_t = reinterpret_cast<float*>(smem + L.stage_off + 1024);     // [NW+1][4D]
  float* s_item = reinterpret_cast<float*>(smem + L.stage_off + 4096);  // [I*F]
  {
    const double a = warp_sum((double)ll_acc), b = warp_sum((double)term_acc);
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
  if (GRAD) {
    for (int k = threadIdx.x; k < I * F; k += blockDim.x) s_item[k] = 0.0f;
    // fold the PPW sub-groups of a warp (same items, different persons)
#pragma unroll
    for (int k = 0; k < IPL * F; ++k) {
#pragma unroll
      for (int o = LPP; o < 32; o <<= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
  }
  __syncthreads();
  if (GRAD) {
    for (int w = 0; w < NW; ++w) {
      if (warp == w && lane < LPP) {
        const int kf = min(n_groups / LPP, NG);
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          // slot k holds group lane + LPP*k, except that the last slot holds the ragged group
          const int g = lane + LPP * ((k == NG - 1 && kf < NG) ? kf : k);
          if ((k < kf || k == NG - 1) && g < n_groups) {
#pragma unroll
            for (int cidx = 0; cidx < 4; ++cidx)
#pragma unroll
              for (int f = 0; f < F; ++f) s_item[(size_t)(4 * g + cidx) * F + f] += acc[(k * 4 + cidx) * F + f];
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
    for (int w = 0; w < NW; ++w) {
      a += s_d[w * 2];
      b += s_d[w * 2 + 1];
    }
    p.part_scalar[(size_t)blockIdx.x * 2] = a;
    p.part_scalar[(size_t)blockIdx.x * 2 + 1] = b;
  }
  if (GRAD && threadIdx.x < 4 * D) {
    float v = 0.0f;
    for (int w = 0; w < NW; ++w) v += s_t[w * 4 * D + threadIdx.x];
    p.part_table[(size_t)blockIdx.x * 4 * D + threadIdx.x] = v;
  }
}

// Launcher for one (MODEL, D); picks LPP / NG from I.  Defined per
// instantiation file so the variants compile in parallel.
template <int MODEL, int D>
cudaError_t launch_fused_md(const FusedParams& p, int grid, size_t smem, bool grad, cudaStream_t st);

template <int MODEL, int D, int LPP, int NG>
static cudaError_t launch_fused_cfg(const FusedParams& p, int grid, size_t smem, bool grad, cudaStream_t st) {
  // the opt-in to > 48 KB dynamic shared memory is per device and per function: set it on every
  // launch (a host-side attribute write, ~1 us) instead of caching it in process-wide state
  cudaError_t e;
  if (grad) {
    auto k = fused_uncond_kernel<MODEL, D, LPP, NG, true>;
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<grid, kFusedThreads, smem, st>>>(p);
  } else {
    auto k = fused_uncond_kernel<MODEL, D, LPP, NG, false>;
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<grid, kFusedThreads, smem, st>>>(p);
  }
  return cudaGetLastError();
}

// lanes per person and 4-item groups per lane for I items (I % 4 == 0, I <= 1024)
inline void fused_pick(int I, int* lpp, int* ng) {
  const int groups = I / 4;
  int l = groups <= 8 * 8 ? 8 : (groups <= 16 * 8 ? 16 : 32);
  int per = (groups + l - 1) / l;
  *lpp = l;
  *ng = per <= 4 ? 4 : 8;
}

#define VIBO_FUSED_INSTANTIATE(MODEL, D)                                                              \
  template <>                                                                                         \
  cudaError_t launch_fused_md<MODEL, D>(const FusedParams& p, int grid, size_t smem, bool grad,       \
                                        cudaStream_t st) {                                            \
    int lpp, ng;                                                                                      \
    fused_pick(p.I, &lpp, &ng);                                                                       \
    if (lpp == 8 && ng == 4) return launch_fused_cfg<MODEL, D, 8, 4>(p, grid, smem, grad, st);        \
    if (lpp == 8 && ng == 8) return launch_fused_cfg<MODEL, D, 8, 8>(p, grid, smem, grad, st);        \
    if (lpp == 16 && ng == 4) return launch_fused_cfg<MODEL, D, 16, 4>(p, grid, smem, grad, st);      \
    if (lpp == 16 && ng == 8) return launch_fused_cfg<MODEL, D, 16, 8>(p, grid, smem, grad, st);      \
    if (lpp == 32 && ng == 4) return launch_fused_cfg<MODEL, D, 32, 4>(p, grid, smem, grad, st);      \
    return launch_fused_cfg<MODEL, D, 32, 8>(p, grid, smem, grad, st);                                \
  }

}  // namespace vibo
