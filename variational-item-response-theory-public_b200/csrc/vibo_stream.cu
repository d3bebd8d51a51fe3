// Planning and launch of the slab-stream kernels (vibo_stream_kernel.cuh).
#include <cstdlib>

#include "vibo_stream_kernel.cuh"

namespace vibo {

namespace {

constexpr size_t kStreamSmemCap = 227 * 1024;

inline size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int gcd_i(int a, int b) {
  while (b) {
    const int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// kind: 0 encode, 1 link, 2 encode backward.  n_parr: staged per-person arrays.
static int env_int(const char* name) {
  const char* v = getenv(name);
  return v != nullptr ? atoi(v) : 0;
}

StreamPlan stream_plan(const vibo_desc& d, int kind, int n_parr) {
  StreamPlan pl;
  const int I = d.num_item, D = d.ability_dim;
  const char* off = getenv("VIBO_DISABLE_STREAM");
  if (off != nullptr && off[0] == '1') return pl;
  // items per lane: the register budget of the lane-owned state bounds M
  int M = I <= 512 ? 1 : (I <= 1024 ? 2 : 4);
  if (I > 2048) return pl;
  const int F = d.irt_model == 1 ? 1 : (d.irt_model == 2 ? D + 1 : D + 2);
  const int regs = kind == 0 ? (d.conditional ? 4 * D : 0) : (kind == 1 ? 2 * F + 2 : 4 * D);
  if (kind == 1) {
    // the link kernel works on f32x2 pairs of items; 4 items per lane halve the share of the
    // per-row overhead (theta loads, transpose-reduce) when the lane-owned state still fits
    const char* m4 = getenv("VIBO_LINK_M");
    const int want = m4 != nullptr ? atoi(m4) : 4;
    M = (want >= 4 && 4 * regs <= 64 && I > 256) ? 4 : (M < 2 ? 2 : M);
  }
  if (M * regs > 64) return pl;   // beyond this the lane-owned state spills
  const int NW = (I + 32 * M - 1) / (32 * M);
  // rows per stage: a multiple of (a) the alignment quantum of the bulk copies
  // (mask rows are I bytes, per-person rows 4 D bytes) and (b) 8 (reduce tile)
  const int q_mask = 16 / gcd_i(I, 16), q_parr = 4 / gcd_i(D, 4);
  int rq = 8;
  while (rq % q_mask != 0 || rq % q_parr != 0) rq += 8;
  const size_t row_bytes = (size_t)I * 5 + (size_t)n_parr * D * 4;
  // CTAs per SM: aim at >= 16 resident warps
  int ctas = 16 / NW;
  if (ctas < 1) ctas = 1;
  if (ctas > 8) ctas = 8;
  // tuning overrides (experiments only): VIBO_STREAM_CTAS / _R / _NS replace the planned values
  const int env_ctas = env_int("VIBO_STREAM_CTAS"), env_r = env_int("VIBO_STREAM_R"), env_ns = env_int("VIBO_STREAM_NS");
  // narrow rows (at most three slab warps per CTA): the per-chunk latency chain (copy -> wait -> reduce ->
  // barrier) dominates, so run many small CTAs with two ~8 KB stages each (measured on 428478 x 95 against
  // five CTAs with 23 KB stages: encode 127 -> 90 us, encode backward 116 -> 97 us, link with gradients
  // 140 -> 124 us, link 107 -> 89 us; 8 / 12 / 14 / 16 CTAs were all slower than 10)
  const bool narrow = NW <= 3;
  if (narrow) ctas = 10;
  if (env_ctas > 0) ctas = env_ctas;
  const size_t budget = kStreamSmemCap / ctas - 1024;
  const int Qmax = 2 * D;
  int NS = 4, R = rq;
  auto total = [&](int R_, int NS_, StreamPlan* o) {
    const size_t red = up((size_t)2 * R_ * NW * Qmax * 4 + 256, 128);
    const size_t info = up((size_t)R_ * 2 * D * 4, 128);
    const size_t mask_off = (size_t)R_ * I * 4;
    const size_t parr_off = up(mask_off + (size_t)R_ * I, 16);
    const size_t stage = up(parr_off + (size_t)n_parr * R_ * D * 4, 128);
    if (o) {
      o->red_off = 128;
      o->info_off = (uint32_t)(128 + red);
      o->stage_off = (uint32_t)(128 + red + info);
      o->mask_off = (uint32_t)mask_off;
      o->parr_off = (uint32_t)parr_off;
      o->stage_bytes = (uint32_t)stage;
    }
    return 128 + red + info + (size_t)NS_ * stage;
  };
  // grow R while a stage stays <= 24 KB and R <= 64, then fit NS
  // long rows (one CTA per SM): 16-row stages halve the number of CTA-wide barriers per row
  const size_t stage_cap = (ctas == 1 ? 84 : (narrow ? 8 : 24)) * 1024;
  while ((size_t)(R + rq) * row_bytes <= stage_cap && R + rq <= 64 && total(R + rq, 2, nullptr) <= budget) R += rq;
  // small problems: keep enough chunks to occupy the machine
  while (R > rq && (d.num_person + R - 1) / R < 2 * (int64_t)sm_count() * ctas) R -= rq;
  if (env_r > 0 && env_r % (rq < 8 ? rq : 4) == 0 && env_r % q_mask == 0 && env_r % q_parr == 0) R = env_r;
  if (env_ns > 0 && env_ns <= kStreamMaxStages) NS = env_ns;
  while (NS > 2 && total(R, NS, nullptr) > budget) --NS;
  if (total(R, NS, nullptr) > kStreamSmemCap) return pl;
  pl.smem = total(R, NS, &pl);
  if (pl.smem > budget) ctas = (int)(kStreamSmemCap / pl.smem) > 0 ? (int)(kStreamSmemCap / pl.smem) : 1;
  pl.M = M;
  pl.NW = NW;
  pl.R = R;
  pl.NS = NS;
  const int64_t n_chunks = (d.num_person + R - 1) / R;
  int64_t g = (int64_t)sm_count() * ctas;
  if (g > n_chunks) g = n_chunks;
  if (g < 1) g = 1;
  pl.grid = (int)g;
  pl.ok = true;
  return pl;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

StreamParams make_params(const vibo_desc& d, const StreamPlan& pl, const float* resp, const uint8_t* mask,
                         int n_parr, const float* const* parr) {
  StreamParams p;
  p.P = d.num_person;
  p.I = d.num_item;
  p.R = pl.R;
  p.NS = pl.NS;
  p.n_parr = n_parr;
  p.D = d.ability_dim;
  p.mask_off = pl.mask_off;
  p.parr_off = pl.parr_off;
  p.stage_bytes = pl.stage_bytes;
  p.red_off = pl.red_off;
  p.info_off = pl.info_off;
  p.stage_off = pl.stage_off;
  p.resp = resp;
  p.mask = mask;
  for (int a = 0; a < kStreamMaxParr; ++a) p.parr[a] = a < n_parr ? parr[a] : nullptr;
  return p;
}

template <typename K>
cudaError_t set_smem(K kernel, size_t smem) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int D, int M>
cudaError_t run_encode(const vibo_desc& d, const StreamPlan& pl, const StreamParams& p, const float* table,
                       float* mu, float* lv, float* S, float* counts, cudaStream_t st) {
  constexpr int NR = (2 * D <= 8) ? 8 : 4;
  cudaError_t e;
  if (d.conditional) {
    auto k = encode_stream_kernel<D, M, NR, true>;
    if ((e = set_smem(k, pl.smem)) != cudaSuccess) return e;
    k<<<pl.grid, pl.NW * 32, pl.smem, st>>>(p, d.missing_policy, table, mu, lv, S, nullptr);
  } else {
    auto k = encode_stream_kernel<D, M, 8, false>;
    if ((e = set_smem(k, pl.smem)) != cudaSuccess) return e;
    k<<<pl.grid, pl.NW * 32, pl.smem, st>>>(p, d.missing_policy, table, mu, lv, S, counts);
  }
  return cudaGetLastError();
}

// ---- tensor-core conditional encode --------------------------------------------
struct MmaPlan {
  bool ok = false;
  int KS = 0, R = 16, NS = 0, grid = 0;
  size_t smem = 0;
  uint32_t mask_off = 0, parr_off = 0, stage_bytes = 0, red_off = 0, info_off = 0, stage_off = 0;
};
MmaPlan mma_plan(const vibo_desc& d) {
  MmaPlan pl;
  const int I = d.num_item, D = d.ability_dim;
  const char* off = getenv("VIBO_DISABLE_MMA");
  if (off != nullptr && off[0] == '1') return pl;
  if (!d.conditional || I > 1024 || I < 16) return pl;
  const int ksteps = (I + 15) / 16;
  int KS = (ksteps + kMmaWarps - 1) / kMmaWarps;
  KS = KS <= 1 ? 1 : (KS <= 2 ? 2 : (KS <= 4 ? 4 : 8));
  const int NT = (6 * D + 7) / 8, QC = 2 * D + 1;
  // rows per stage: multiples of 16 that keep the bulk copies 16-byte multiples
  const int q_mask = 16 / gcd_i(I, 16);
  int R = 16;
  while (R % q_mask != 0) R += 16;
  const size_t red = up((size_t)kMmaWarps * 16 * NT * 8 * 4, 128);
  const size_t info = up(((size_t)kMmaWarps * 16 * QC + 2 * D + (size_t)I * 2 * D) * 4, 128);
  const size_t mask_off = (size_t)R * I * 4;
  const size_t stage = up(mask_off + (size_t)R * I, 128);
  int NS = 4;
  while (NS > 2 && 128 + red + info + (size_t)NS * stage > kStreamSmemCap) --NS;
  const size_t total = 128 + red + info + (size_t)NS * stage;
  if (total > kStreamSmemCap) return pl;
  pl.KS = KS;
  pl.R = R;
  pl.NS = NS;
  pl.smem = total;
  pl.red_off = 128;
  pl.info_off = (uint32_t)(128 + red);
  pl.stage_off = (uint32_t)(128 + red + info);
  pl.mask_off = (uint32_t)mask_off;
  pl.parr_off = (uint32_t)stage;
  pl.stage_bytes = (uint32_t)stage;
  const int64_t n_chunks = (d.num_person + R - 1) / R;
  int ctas = (int)(kStreamSmemCap / total);
  if (ctas < 1) ctas = 1;
  if (ctas > 4) ctas = 4;
  int64_t g = (int64_t)sm_count() * ctas;
  if (g > n_chunks) g = n_chunks;
  if (g < 1) g = 1;
  pl.grid = (int)g;
  pl.ok = true;
  return pl;
}

// ---- tensor-core encode backward (conditional table) ----------------------------
struct BwdMmaPlan {
  bool ok = false;
  int MT = 0, grid = 0;
  StreamPlan sp;
};
BwdMmaPlan bwd_mma_plan(const vibo_desc& d) {
  BwdMmaPlan pl;
  const int I = d.num_item, D = d.ability_dim;
  const char* off = getenv("VIBO_DISABLE_MMA");
  if (off != nullptr && off[0] == '1') return pl;
  if (!d.conditional || I > 1024 || I < 16) return pl;
  const int NT = (4 * D + 7) / 8, NC = NT * 8, NW = 16;
  const int MT = I <= 256 ? 1 : (I <= 512 ? 2 : 4);
  if (MT * NT * 8 > 96) return pl;   // accumulator registers
  const int q_mask = 16 / gcd_i(I, 16), q_parr = 4 / gcd_i(D, 4);
  int rq = 8;
  while (rq % q_mask != 0 || rq % q_parr != 0) rq += 8;
  int R = rq;
  const size_t row_bytes = (size_t)I * 5 + 4 * (size_t)D * 4;
  while ((size_t)(R + rq) * row_bytes <= 84 * 1024 && R + rq <= 64) R += rq;
  while (R > rq && (d.num_person + R - 1) / R < 2 * (int64_t)sm_count()) R -= rq;
  const size_t red = up((size_t)2 * R * NC * 4 + 256, 128), info = 128;   // G tile double-buffered; 2 x 16 warp flags
  const size_t mask_off = (size_t)R * I * 4;
  const size_t parr_off = up(mask_off + (size_t)R * I, 16);
  const size_t stage = up(parr_off + 4 * (size_t)R * D * 4, 128);
  const size_t tile = (size_t)NW * 16 * MT * NC * 4;   // epilogue staging reuses the stages
  int NS = 5;
  while (NS > 2 && 128 + red + info + (size_t)NS * stage > kStreamSmemCap) --NS;
  size_t total = 128 + red + info + (size_t)NS * stage;
  if (total > kStreamSmemCap) return pl;
  if ((size_t)NS * stage < tile) total = 128 + red + info + tile;
  if (total > kStreamSmemCap) return pl;
  pl.sp.M = 0; pl.sp.NW = NW; pl.sp.R = R; pl.sp.NS = NS; pl.sp.smem = total;
  pl.sp.red_off = 128;
  pl.sp.info_off = (uint32_t)(128 + red);
  pl.sp.stage_off = (uint32_t)(128 + red + info);
  pl.sp.mask_off = (uint32_t)mask_off;
  pl.sp.parr_off = (uint32_t)parr_off;
  pl.sp.stage_bytes = (uint32_t)stage;
  const int64_t n_chunks = (d.num_person + R - 1) / R;
  int64_t g = sm_count();
  if (g > n_chunks) g = n_chunks;
  if (g < 1) g = 1;
  pl.grid = pl.sp.grid = (int)g;
  pl.MT = MT;
  pl.ok = true;
  return pl;
}

template <int D, int M>
cudaError_t run_encode_bwd(const vibo_desc& d, const StreamPlan& pl, const StreamParams& p, float* part,
                           cudaStream_t st) {
  auto k = encode_bwd_stream_kernel<D, M>;
  cudaError_t e;
  if ((e = set_smem(k, pl.smem)) != cudaSuccess) return e;
  k<<<pl.grid, pl.NW * 32, pl.smem, st>>>(p, d.conditional, part);
  return cudaGetLastError();
}

#define VIBO_STREAM_SWITCH_D(D_, ...)                     \
  switch (D_) {                                           \
    case 1: { constexpr int kD = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int kD = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int kD = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int kD = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int kD = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int kD = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int kD = 7; __VA_ARGS__; } break; \
    case 8: { constexpr int kD = 8; __VA_ARGS__; } break; \
    default: return cudaErrorInvalidValue;                \
  }
#define VIBO_STREAM_SWITCH_M(M_, ...)                     \
  switch (M_) {                                           \
    case 1: { constexpr int kM = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int kM = 2; __VA_ARGS__; } break; \
    case 4: { constexpr int kM = 4; __VA_ARGS__; } break; \
    default: return cudaErrorInvalidValue;                \
  }

}  // namespace

// ---- public (library-internal) entry points ------------------------------------
// Each returns cudaErrorNotSupported when the configuration / pointers are not
// covered; the caller then uses the legacy kernels of vibo_general.cu.

cudaError_t stream_encode(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* table,
                          float* mu, float* lv, float* S, int* grid_out, cudaStream_t st) {
  if (!aligned16(resp) || !aligned16(mask)) return cudaErrorNotSupported;
  {
    // conditional posterior, D <= 5, I % 4 == 0: tcgen05 / TMA streaming kernel (vibo_tc5_encode.cu)
    const cudaError_t e5 = tc5_encode(d, resp, mask, table, mu, lv, S, st);
    if (e5 != cudaErrorNotSupported) {
      if (grid_out) *grid_out = sm_count();
      return e5;
    }
  }
  const MmaPlan mp = mma_plan(d);
  if (mp.ok) {
    StreamPlan sp;
    sp.R = mp.R; sp.NS = mp.NS; sp.mask_off = mp.mask_off; sp.parr_off = mp.parr_off;
    sp.stage_bytes = mp.stage_bytes; sp.red_off = mp.red_off; sp.info_off = mp.info_off; sp.stage_off = mp.stage_off;
    const StreamParams p = make_params(d, sp, resp, mask, 0, nullptr);
    cudaError_t e = cudaSuccess;
    e = stream_encode_mma_run(d.ability_dim, mp.KS, (d.num_item & 1) == 0, mp.grid, mp.smem, p, d.missing_policy,
                              table, mu, lv, S, st);
    if (grid_out) *grid_out = mp.grid;
    return e;
  }
  const StreamPlan pl = stream_plan(d, 0, 0);
  if (!pl.ok) return cudaErrorNotSupported;
  const StreamParams p = make_params(d, pl, resp, mask, 0, nullptr);
  cudaError_t e = cudaSuccess;
  VIBO_STREAM_SWITCH_D(d.ability_dim,
                       VIBO_STREAM_SWITCH_M(pl.M, e = (run_encode<kD, kM>(d, pl, p, table, mu, lv, S, nullptr, st))));
  if (grid_out) *grid_out = pl.grid;
  return e;
}

// Unconditional posterior AND the per-person counts (n1, n_observed) it was formed from, in one pass:
// the backward then needs no second pass over the rows (encode_bwd_counts_kernel, vibo_general.cu).
cudaError_t stream_encode_counts(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* table,
                                 float* mu, float* lv, float* S, float* counts, cudaStream_t st) {
  if (d.conditional || counts == nullptr || !aligned16(resp) || !aligned16(mask)) return cudaErrorNotSupported;
  const StreamPlan pl = stream_plan(d, 0, 0);
  if (!pl.ok) return cudaErrorNotSupported;
  const StreamParams p = make_params(d, pl, resp, mask, 0, nullptr);
  cudaError_t e = cudaSuccess;
  VIBO_STREAM_SWITCH_D(d.ability_dim,
                       VIBO_STREAM_SWITCH_M(pl.M, e = (run_encode<kD, kM>(d, pl, p, table, mu, lv, S, counts, st))));
  return e;
}

// (n1, n_observed) per person, (P, 2) floats: the sufficient statistics of an unconditional encoder.
cudaError_t stream_counts(const vibo_desc& d0, const float* resp, const uint8_t* mask, float* counts,
                          cudaStream_t st) {
  if (!aligned16(resp) || !aligned16(mask)) return cudaErrorNotSupported;
  vibo_desc d = d0;
  d.conditional = 0;
  d.ability_dim = 1;
  const StreamPlan pl = stream_plan(d, 0, 0);
  if (!pl.ok) return cudaErrorNotSupported;
  const StreamParams p = make_params(d, pl, resp, mask, 0, nullptr);
  cudaError_t e = cudaErrorInvalidValue;
#define VIBO_COUNTS_CASE(M_)                                                                   \
  case M_: {                                                                                   \
    auto k = encode_stream_kernel<1, M_, 8, false>;                                            \
    if ((e = set_smem(k, pl.smem)) != cudaSuccess) return e;                                   \
    k<<<pl.grid, pl.NW * 32, pl.smem, st>>>(p, d.missing_policy, nullptr, nullptr, nullptr, nullptr, counts); \
    e = cudaGetLastError();                                                                    \
  } break;
  switch (pl.M) {
    VIBO_COUNTS_CASE(1)
    VIBO_COUNTS_CASE(2)
    VIBO_COUNTS_CASE(4)
    default: break;
  }
#undef VIBO_COUNTS_CASE
  return e;
}

cudaError_t stream_link(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* ability,
                        const float* item_feat, double* part_ll, float* g_ability, float* part_g, bool grad,
                        int* grid_out, cudaStream_t st) {
  if (!aligned16(resp) || !aligned16(mask) || !aligned16(ability)) return cudaErrorNotSupported;
  const StreamPlan pl = stream_plan(d, 1, 1);
  if (!pl.ok) return cudaErrorNotSupported;
  const float* parr[1] = {ability};
  const StreamParams p = make_params(d, pl, resp, mask, 1, parr);
  cudaError_t e = cudaSuccess;
  switch (d.irt_model) {
    case 1: e = stream_link_run1(pl, p, d.ability_dim, item_feat, part_ll, g_ability, part_g, grad, st); break;
    case 2: e = stream_link_run2(pl, p, d.ability_dim, item_feat, part_ll, g_ability, part_g, grad, st); break;
    case 3: e = stream_link_run3(pl, p, d.ability_dim, item_feat, part_ll, g_ability, part_g, grad, st); break;
    default: return cudaErrorInvalidValue;
  }
  if (grid_out) *grid_out = pl.grid;
  return e;
}

cudaError_t stream_encode_bwd(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* amu,
                              const float* S, const float* g_mu, const float* g_lv, float* part, int* grid_out,
                              cudaStream_t st) {
  if (!aligned16(resp) || !aligned16(mask) || !aligned16(amu) || !aligned16(S) || !aligned16(g_mu) ||
      !aligned16(g_lv))
    return cudaErrorNotSupported;
  {
    // conditional posterior, D <= 5, I % 4 == 0: tcgen05 / TMA kernel (vibo_tc5_encode_bwd.cu)
    const cudaError_t e5 = tc5_encode_bwd(d, resp, mask, amu, S, g_mu, g_lv, part, grid_out, st);
    if (e5 != cudaErrorNotSupported) return e5;
  }
  const float* parr[4] = {amu, S, g_mu, g_lv};
  const BwdMmaPlan mp = bwd_mma_plan(d);
  if (mp.ok) {
    const StreamParams p = make_params(d, mp.sp, resp, mask, 4, parr);
    cudaError_t e = cudaSuccess;
    e = stream_encode_bwd_mma_run(d.ability_dim, mp.MT, mp.grid, mp.sp.smem, p, part, st);
    if (grid_out) *grid_out = mp.grid;
    return e;
  }
  const StreamPlan pl = stream_plan(d, 2, 4);
  if (!pl.ok) return cudaErrorNotSupported;
  const StreamParams p = make_params(d, pl, resp, mask, 4, parr);
  cudaError_t e = cudaSuccess;
  VIBO_STREAM_SWITCH_D(d.ability_dim, VIBO_STREAM_SWITCH_M(pl.M, e = (run_encode_bwd<kD, kM>(d, pl, p, part, st))));
  if (grid_out) *grid_out = pl.grid;
  return e;
}

}  // namespace vibo
