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
  while ((size_t)(R + rq) * row_bytes <= 24 * 1024 && R + rq <= 64 && total(R + rq, 2, nullptr) <= budget) R += rq;
  // small problems: keep enough chunks to occupy the machine
  while (R > rq && (d.num_person + R - 1) / R < 2 * (int64_t)sm_count() * ctas) R -= rq;
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
                       float* mu, float* lv, float* S, cudaStream_t st) {
  constexpr int NR = (2 * D <= 8) ? 8 : 4;
  cudaError_t e;
  if (d.conditional) {
    auto k = encode_stream_kernel<D, M, NR, true>;
    if ((e = set_smem(k, pl.smem)) != cudaSuccess) return e;
    k<<<pl.grid, pl.NW * 32, pl.smem, st>>>(p, d.missing_policy, table, mu, lv, S);
  } else {
    auto k = encode_stream_kernel<D, M, 8, false>;
    if ((e = set_smem(k, pl.smem)) != cudaSuccess) return e;
    k<<<pl.grid, pl.NW * 32, pl.smem, st>>>(p, d.missing_policy, table, mu, lv, S);
  }
  return cudaGetLastError();
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
  const StreamPlan pl = stream_plan(d, 0, 0);
  if (!pl.ok) return cudaErrorNotSupported;
  const StreamParams p = make_params(d, pl, resp, mask, 0, nullptr);
  cudaError_t e = cudaSuccess;
  VIBO_STREAM_SWITCH_D(d.ability_dim,
                       VIBO_STREAM_SWITCH_M(pl.M, e = (run_encode<kD, kM>(d, pl, p, table, mu, lv, S, st))));
  if (grid_out) *grid_out = pl.grid;
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
  const StreamPlan pl = stream_plan(d, 2, 4);
  if (!pl.ok) return cudaErrorNotSupported;
  const float* parr[4] = {amu, S, g_mu, g_lv};
  const StreamParams p = make_params(d, pl, resp, mask, 4, parr);
  cudaError_t e = cudaSuccess;
  VIBO_STREAM_SWITCH_D(d.ability_dim, VIBO_STREAM_SWITCH_M(pl.M, e = (run_encode_bwd<kD, kM>(d, pl, p, part, st))));
  if (grid_out) *grid_out = pl.grid;
  return e;
}

}  // namespace vibo
