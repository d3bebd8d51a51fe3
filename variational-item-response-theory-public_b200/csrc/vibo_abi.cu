// C ABI of libvibo_b200.so (see include/vibo_b200.h for the contract).
#include <cstdio>
#include <cstring>
#include <string>

#include "vibo_kernels.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* where) {
  return fail(VIBO_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

#define VIBO_CUDA(call, where)                         \
  do {                                                 \
    cudaError_t e__ = (call);                          \
    if (e__ != cudaSuccess) return cuda_fail(e__, where); \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int check_desc(const vibo_desc* d) {
  if (d == nullptr) return fail(VIBO_ERR_BAD_ARGUMENT, "desc is NULL");
  if (d->num_person < 0 || d->num_item <= 0)
    return fail(VIBO_ERR_BAD_ARGUMENT, "num_person must be >= 0 and num_item > 0");
  if (d->irt_model < 1 || d->irt_model > 3)
    return fail(VIBO_ERR_BAD_ARGUMENT, "irt_model must be 1, 2 or 3");
  if (d->ability_dim < 1) return fail(VIBO_ERR_BAD_ARGUMENT, "ability_dim must be >= 1");
  if (d->ability_dim > VIBO_MAX_ABILITY_DIM)
    return fail(VIBO_ERR_UNSUPPORTED, "ability_dim > VIBO_MAX_ABILITY_DIM (8)");
  if (d->missing_policy != VIBO_MISSING_PRIOR && d->missing_policy != VIBO_MISSING_DROP)
    return fail(VIBO_ERR_BAD_ARGUMENT, "missing_policy must be VIBO_MISSING_PRIOR or VIBO_MISSING_DROP");
  if (d->elbo_form != VIBO_ELBO_KL && d->elbo_form != VIBO_ELBO_SAMPLE)
    return fail(VIBO_ERR_BAD_ARGUMENT, "elbo_form must be VIBO_ELBO_KL or VIBO_ELBO_SAMPLE");
  if (d->conditional != 0 && d->conditional != 1)
    return fail(VIBO_ERR_BAD_ARGUMENT, "conditional must be 0 or 1");
  return VIBO_OK;
}

int check_items(const vibo_desc* d) {
  if (d->num_item > vibo::general_max_items(d->ability_dim)) {
    char buf[160];
    snprintf(buf, sizeof buf, "num_item %d exceeds the limit %d for ability_dim %d", d->num_item,
             vibo::general_max_items(d->ability_dim), d->ability_dim);
    return fail(VIBO_ERR_UNSUPPORTED, buf);
  }
  return VIBO_OK;
}

// Carves the workspace of the composed (multi-pass) path.
struct Carve {
  char* base;
  size_t off = 0, cap;
  Carve(void* p, size_t n) : base(static_cast<char*>(p)), cap(n) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

size_t general_workspace_bytes(const vibo_desc* d) {
  const size_t P = (size_t)d->num_person, I = (size_t)d->num_item, D = (size_t)d->ability_dim;
  const size_t F = (size_t)vibo::item_width_host(d->irt_model, d->ability_dim);
  const size_t G = (size_t)vibo::sm_count() * 8;  // upper bound of general_grid / person_grid
  size_t b = 0;
  b += align_up(8 * (G + 8), 256) * 3;                      // double partials
  b += (align_up(4 * P * D, 256) + 256) * 8;                // per-person float scratch
  b += align_up(4 * G * I * F, 256) + 256;                  // item-gradient partials
  b += align_up(4 * G * 2 * I * 2 * D, 256) + 256;          // expert-table partials
  b += align_up(4 * P * 2, 256) + 256;                      // per-person counts (unconditional backward)
  return b + 4096;
}

}  // namespace

namespace vibo {
int set_last_error(int code, const char* msg) { return fail(code, msg); }
int item_width_host(int model, int D) { return model == 1 ? 1 : (model == 2 ? D + 1 : D + 2); }
}  // namespace vibo

extern "C" {

int vibo_version(void) { return VIBO_B200_VERSION; }

int vibo_max_items(const vibo_desc* desc) {
  if (desc == nullptr || desc->ability_dim < 1 || desc->ability_dim > VIBO_MAX_ABILITY_DIM) return 0;
  return vibo::general_max_items(desc->ability_dim);
}

int vibo_single_pass(const vibo_desc* desc) {
  if (check_desc(desc) != VIBO_OK) return 0;
  return vibo::fused_supported(*desc, nullptr, nullptr) ? 1 : 0;
}

uint64_t vibo_launch_count(void) { return (uint64_t)vibo::launch_count(); }

int vibo_profile_begin(void) {
  vibo::profile_begin();
  return VIBO_OK;
}

int vibo_profile_end(int* n_launches, double* total_ms) { return vibo::profile_end(n_launches, total_ms); }

const char* vibo_last_error(void) { return g_last_error.c_str(); }

size_t vibo_workspace_bytes(const vibo_desc* desc) {
  if (check_desc(desc) != VIBO_OK) return 0;
  const size_t a = general_workspace_bytes(desc);
  const size_t b = vibo::fused_workspace_bytes(*desc);
  return a > b ? a : b;
}

static int fused_elbo_impl(const vibo_desc* desc, const float* response, const uint8_t* mask,
                           const float* table, const float* item_feat, const float* eps_ability,
                           uint64_t seed, const uint64_t* seed_dev, float beta, double* out_scalars,
                           float* ability_mu, float* ability_logvar, float* ability, float* g_table,
                           float* g_item, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!response || !mask || !table || !item_feat || !out_scalars)
    return fail(VIBO_ERR_BAD_ARGUMENT, "response, mask, table, item_feat and out_scalars are required");
  if ((g_table == nullptr) != (g_item == nullptr))
    return fail(VIBO_ERR_BAD_ARGUMENT, "g_table and g_item must both be given or both be NULL");
  if (workspace == nullptr || workspace_bytes < vibo_workspace_bytes(desc))
    return fail(VIBO_ERR_WORKSPACE, "workspace smaller than vibo_workspace_bytes(desc)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const vibo_desc& d = *desc;
  if (d.num_person == 0) {
    VIBO_CUDA(cudaMemsetAsync(out_scalars, 0, 2 * sizeof(double), st), "memset");
    const int F = vibo::item_width_host(d.irt_model, d.ability_dim);
    if (g_item) {
      VIBO_CUDA(cudaMemsetAsync(g_item, 0, sizeof(float) * d.num_item * F, st), "memset");
      VIBO_CUDA(cudaMemsetAsync(g_table, 0,
                                sizeof(float) * 2 * (d.conditional ? d.num_item : 1) * 2 * d.ability_dim, st),
                "memset");
    }
    return VIBO_OK;
  }
  if (vibo::fused_supported(d, response, mask)) {
    VIBO_CUDA(vibo::launch_fused(d, response, mask, table, item_feat, eps_ability, seed, seed_dev, beta,
                                 out_scalars, ability_mu, ability_logvar, ability, g_table, g_item,
                                 workspace, workspace_bytes, false, st),
              "fused kernel");
    return VIBO_OK;
  }
  if (int rc = check_items(desc)) return rc;

  const bool grad = g_item != nullptr;
  if (!grad && d.conditional) {
    // forward-only evaluation of the conditional posterior: ONE pass (tcgen05 encode, link from on-chip bits)
    const cudaError_t e5 = vibo::tc5_eval(d, response, mask, table, item_feat, eps_ability, seed, seed_dev,
                                          out_scalars, ability_mu, ability_logvar, ability, workspace,
                                          workspace_bytes, st);
    if (e5 == cudaSuccess) return VIBO_OK;
    if (e5 != cudaErrorNotSupported) return cuda_fail(e5, "tc5_eval");
    (void)cudaGetLastError();
  }

  // Composition of the general kernels (three passes over the rows).
  const size_t PD = (size_t)d.num_person * d.ability_dim;
  const size_t F = (size_t)vibo::item_width_host(d.irt_model, d.ability_dim);
  const size_t G = (size_t)vibo::sm_count() * 8;
  Carve ws(workspace, workspace_bytes);
  double* part_ll = ws.take<double>(G + 8);
  double* part_term = ws.take<double>(G + 8);
  float* S = ws.take<float>(PD);
  float* amu = ability_mu ? ability_mu : ws.take<float>(PD);
  float* alv = ability_logvar ? ability_logvar : ws.take<float>(PD);
  float* th = ability ? ability : ws.take<float>(PD);
  float* eps_buf = ws.take<float>(PD);
  float* g_ab = ws.take<float>(PD);
  float* g_mu = ws.take<float>(PD);
  float* g_lv = ws.take<float>(PD);
  float* part_g = ws.take<float>(G * d.num_item * F);
  float* part_ab = ws.take<float>(G * 2 * d.num_item * 2 * d.ability_dim);
  float* counts = ws.take<float>((size_t)d.num_person * 2);
  if (!ws.ok()) return fail(VIBO_ERR_WORKSPACE, "workspace carve overflow");

  // unconditional posterior with gradients: the encode pass also leaves the per-person counts, from which the
  // backward forms the table gradient without a third pass over the rows
  bool have_counts = false;
  if (grad && !d.conditional) {
    const cudaError_t ec = vibo::launch_encode_counts(d, response, mask, table, amu, alv, S, counts, st);
    if (ec == cudaSuccess) have_counts = true;
    else if (ec != cudaErrorNotSupported) return cuda_fail(ec, "encode");
    else (void)cudaGetLastError();
  }
  if (!have_counts) VIBO_CUDA(vibo::launch_encode(d, response, mask, table, amu, alv, S, st), "encode");
  VIBO_CUDA(vibo::launch_person_forward(d, amu, alv, eps_ability, seed, seed_dev,
                                        eps_ability ? nullptr : eps_buf, th, part_term,
                                        out_scalars + 1, st),
            "person_forward");
  VIBO_CUDA(vibo::launch_link(d, response, mask, th, item_feat, out_scalars, grad ? g_ab : nullptr,
                              g_item, part_ll, part_g, st),
            "link");
  if (grad) {
    VIBO_CUDA(vibo::launch_negate(g_item, (int)(d.num_item * F), st), "negate");
    VIBO_CUDA(vibo::launch_person_backward(d, beta, amu, alv, eps_ability ? eps_ability : eps_buf, th,
                                           g_ab, g_mu, g_lv, st),
              "person_backward");
    if (have_counts) {
      VIBO_CUDA(vibo::launch_encode_bwd_counts(d, counts, table, amu, S, g_mu, g_lv, g_table, part_ab, st),
                "encode_backward (counts)");
    } else {
      VIBO_CUDA(vibo::launch_encode_bwd(d, response, mask, table, amu, S, g_mu, g_lv, g_table, part_ab, st),
                "encode_backward");
    }
  }
  return VIBO_OK;
}

int vibo_fused_elbo(const vibo_desc* desc, const float* response, const uint8_t* mask,
                    const float* table, const float* item_feat, const float* eps_ability,
                    uint64_t seed, float beta, double* out_scalars, float* ability_mu,
                    float* ability_logvar, float* ability, float* g_table, float* g_item,
                    void* workspace, size_t workspace_bytes, void* stream) {
  return fused_elbo_impl(desc, response, mask, table, item_feat, eps_ability, seed, nullptr, beta,
                         out_scalars, ability_mu, ability_logvar, ability, g_table, g_item, workspace,
                         workspace_bytes, stream);
}

int vibo_fused_elbo_graph(const vibo_desc* desc, const float* response, const uint8_t* mask,
                          const float* table, const float* item_feat, const uint64_t* seed_state,
                          float beta, double* out_scalars, float* ability_mu, float* ability_logvar,
                          float* ability, float* g_table, float* g_item, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (seed_state == nullptr) return fail(VIBO_ERR_BAD_ARGUMENT, "seed_state is required");
  return fused_elbo_impl(desc, response, mask, table, item_feat, nullptr, 0, seed_state, beta,
                         out_scalars, ability_mu, ability_logvar, ability, g_table, g_item, workspace,
                         workspace_bytes, stream);
}

int vibo_philox_normal(const vibo_desc* desc, uint64_t seed, const uint64_t* seed_state, float* eps,
                       void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!eps) return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (desc->num_person == 0) return VIBO_OK;
  VIBO_CUDA(vibo::launch_philox_fill(desc->num_person, desc->ability_dim, desc->person_offset, seed,
                                     seed_state, eps, static_cast<cudaStream_t>(stream)),
            "philox_fill");
  return VIBO_OK;
}

int vibo_encode(const vibo_desc* desc, const float* response, const uint8_t* mask,
                const float* table, float* ability_mu, float* ability_logvar,
                float* precision_sum, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!response || !mask || !table || !ability_mu || !ability_logvar)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (int rc = check_items(desc)) return rc;
  if (desc->num_person == 0) return VIBO_OK;
  VIBO_CUDA(vibo::launch_encode(*desc, response, mask, table, ability_mu, ability_logvar,
                                precision_sum, static_cast<cudaStream_t>(stream)),
            "encode");
  return VIBO_OK;
}

int vibo_encode_counts(const vibo_desc* desc, const float* response, const uint8_t* mask,
                       const float* table, float* ability_mu, float* ability_logvar,
                       float* precision_sum, float* counts, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!response || !mask || !table || !ability_mu || !ability_logvar || !precision_sum || !counts)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (desc->conditional) return fail(VIBO_ERR_UNSUPPORTED, "vibo_encode_counts: unconditional posterior only");
  if (int rc = check_items(desc)) return rc;
  if (desc->num_person == 0) return VIBO_OK;
  const cudaError_t e = vibo::launch_encode_counts(*desc, response, mask, table, ability_mu, ability_logvar,
                                                   precision_sum, counts, static_cast<cudaStream_t>(stream));
  if (e == cudaErrorNotSupported) {
    (void)cudaGetLastError();
    return fail(VIBO_ERR_UNSUPPORTED, "vibo_encode_counts: rows not 16-byte aligned (use vibo_encode)");
  }
  VIBO_CUDA(e, "encode_counts");
  return VIBO_OK;
}

int vibo_encode_backward_counts(const vibo_desc* desc, const float* counts, const float* table,
                                const float* ability_mu, const float* precision_sum,
                                const float* g_ability_mu, const float* g_ability_logvar, float* g_table,
                                void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!counts || !table || !ability_mu || !precision_sum || !g_ability_mu || !g_ability_logvar || !g_table)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (desc->conditional)
    return fail(VIBO_ERR_UNSUPPORTED, "vibo_encode_backward_counts: unconditional posterior only");
  if (workspace == nullptr || workspace_bytes < vibo_workspace_bytes(desc))
    return fail(VIBO_ERR_WORKSPACE, "workspace smaller than vibo_workspace_bytes(desc)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (desc->num_person == 0) {
    VIBO_CUDA(cudaMemsetAsync(g_table, 0, sizeof(float) * 2 * 2 * desc->ability_dim, st), "memset");
    return VIBO_OK;
  }
  Carve ws(workspace, workspace_bytes);
  float* part = ws.take<float>((size_t)vibo::sm_count() * 4 * 4 * desc->ability_dim);
  VIBO_CUDA(vibo::launch_encode_bwd_counts(*desc, counts, table, ability_mu, precision_sum, g_ability_mu,
                                           g_ability_logvar, g_table, part, st),
            "encode_backward_counts");
  return VIBO_OK;
}

int vibo_person_counts(const vibo_desc* desc, const float* response, const uint8_t* mask, float* counts,
                       void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!response || !mask || !counts) return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (desc->num_person == 0) return VIBO_OK;
  VIBO_CUDA(vibo::launch_person_counts(*desc, response, mask, counts, static_cast<cudaStream_t>(stream)),
            "person_counts");
  return VIBO_OK;
}

int vibo_encode_backward(const vibo_desc* desc, const float* response, const uint8_t* mask,
                         const float* table, const float* ability_mu,
                         const float* precision_sum, const float* g_ability_mu,
                         const float* g_ability_logvar, float* g_table, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!response || !mask || !table || !ability_mu || !precision_sum || !g_ability_mu ||
      !g_ability_logvar || !g_table)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (int rc = check_items(desc)) return rc;
  if (workspace == nullptr || workspace_bytes < vibo_workspace_bytes(desc))
    return fail(VIBO_ERR_WORKSPACE, "workspace smaller than vibo_workspace_bytes(desc)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const vibo_desc& d = *desc;
  if (d.num_person == 0) {
    VIBO_CUDA(cudaMemsetAsync(g_table, 0,
                              sizeof(float) * 2 * (d.conditional ? d.num_item : 1) * 2 * d.ability_dim, st),
              "memset");
    return VIBO_OK;
  }
  Carve ws(workspace, workspace_bytes);
  const size_t G = (size_t)vibo::sm_count() * 8;
  float* part_ab = ws.take<float>(G * 2 * d.num_item * 2 * d.ability_dim);
  VIBO_CUDA(vibo::launch_encode_bwd(d, response, mask, table, ability_mu, precision_sum,
                                    g_ability_mu, g_ability_logvar, g_table, part_ab, st),
            "encode_backward");
  return VIBO_OK;
}

int vibo_link_loglik(const vibo_desc* desc, const float* response, const uint8_t* mask,
                     const float* ability, const float* item_feat, double* out_ll,
                     float* g_ability, float* g_item, void* workspace, size_t workspace_bytes,
                     void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!response || !mask || !ability || !item_feat || !out_ll)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if ((g_ability == nullptr) != (g_item == nullptr))
    return fail(VIBO_ERR_BAD_ARGUMENT, "g_ability and g_item must both be given or both be NULL");
  if (int rc = check_items(desc)) return rc;
  if (workspace == nullptr || workspace_bytes < vibo_workspace_bytes(desc))
    return fail(VIBO_ERR_WORKSPACE, "workspace smaller than vibo_workspace_bytes(desc)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const vibo_desc& d = *desc;
  const size_t F = (size_t)vibo::item_width_host(d.irt_model, d.ability_dim);
  if (d.num_person == 0) {
    VIBO_CUDA(cudaMemsetAsync(out_ll, 0, sizeof(double), st), "memset");
    if (g_item) VIBO_CUDA(cudaMemsetAsync(g_item, 0, sizeof(float) * d.num_item * F, st), "memset");
    return VIBO_OK;
  }
  Carve ws(workspace, workspace_bytes);
  const size_t G = (size_t)vibo::sm_count() * 8;
  double* part_ll = ws.take<double>(G + 8);
  float* part_g = ws.take<float>(G * d.num_item * F);
  VIBO_CUDA(vibo::launch_link(d, response, mask, ability, item_feat, out_ll, g_ability, g_item,
                              part_ll, part_g, st),
            "link");
  return VIBO_OK;
}

int vibo_decode(const vibo_desc* desc, const float* ability, const float* item_feat,
                float* response_mu, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!ability || !item_feat || !response_mu) return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (desc->num_person == 0) return VIBO_OK;
  VIBO_CUDA(vibo::launch_decode(*desc, ability, item_feat, response_mu,
                                static_cast<cudaStream_t>(stream)),
            "decode");
  return VIBO_OK;
}

int vibo_bernoulli_loglik(const vibo_desc* desc, const float* response, const uint8_t* mask,
                          const float* response_mu, double* out_ll, float* g_prob,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (!response || !mask || !response_mu || !out_ll) return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (workspace == nullptr || workspace_bytes < vibo_workspace_bytes(desc))
    return fail(VIBO_ERR_WORKSPACE, "workspace smaller than vibo_workspace_bytes(desc)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (desc->num_person == 0) {
    VIBO_CUDA(cudaMemsetAsync(out_ll, 0, sizeof(double), st), "memset");
    return VIBO_OK;
  }
  Carve ws(workspace, workspace_bytes);
  double* part_ll = ws.take<double>((size_t)vibo::sm_count() * 8 + 8);
  VIBO_CUDA(vibo::launch_bernoulli_ll(*desc, response, mask, response_mu, out_ll, g_prob, part_ll, st),
            "bernoulli_loglik");
  return VIBO_OK;
}

int vibo_flow_person_forward(const vibo_desc* desc, int n_flows, const float* ability_mu,
                             const float* ability_logvar, const float* eps, const float* uhat,
                             const float* w, const float* b, float* ability_0, float* ability_k,
                             double* out_term, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (n_flows < 1 || n_flows > 8) return fail(VIBO_ERR_UNSUPPORTED, "n_flows must be in 1..8");
  if (!ability_mu || !ability_logvar || !eps || !uhat || !w || !b || !ability_k || !out_term)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (workspace == nullptr || workspace_bytes < vibo_workspace_bytes(desc))
    return fail(VIBO_ERR_WORKSPACE, "workspace smaller than vibo_workspace_bytes(desc)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (desc->num_person == 0) {
    VIBO_CUDA(cudaMemsetAsync(out_term, 0, sizeof(double), st), "memset");
    return VIBO_OK;
  }
  Carve ws(workspace, workspace_bytes);
  double* part_term = ws.take<double>((size_t)vibo::sm_count() * 8 + 8);
  VIBO_CUDA(vibo::launch_flow_person_forward(desc->num_person, desc->ability_dim, n_flows, ability_mu,
                                             ability_logvar, eps, uhat, w, b, ability_0, ability_k, part_term,
                                             out_term, st),
            "flow_person_forward");
  return VIBO_OK;
}

int vibo_flow_person_backward(const vibo_desc* desc, int n_flows, const float* ability_mu,
                              const float* ability_logvar, const float* eps, const float* uhat,
                              const float* w, const float* b, const float* g_ability_k,
                              const float* g_term, float* g_ability_mu, float* g_ability_logvar,
                              float* g_uhat, float* g_w, float* g_b, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (n_flows < 1 || n_flows > 8) return fail(VIBO_ERR_UNSUPPORTED, "n_flows must be in 1..8");
  if (!ability_mu || !ability_logvar || !eps || !uhat || !w || !b || !g_ability_k || !g_term ||
      !g_ability_mu || !g_ability_logvar || !g_uhat || !g_w || !g_b)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (workspace == nullptr || workspace_bytes < vibo_workspace_bytes(desc))
    return fail(VIBO_ERR_WORKSPACE, "workspace smaller than vibo_workspace_bytes(desc)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = desc->ability_dim, n = n_flows * (2 * D + 1);
  if (desc->num_person == 0) {
    VIBO_CUDA(cudaMemsetAsync(g_uhat, 0, sizeof(float) * n_flows * D, st), "memset");
    VIBO_CUDA(cudaMemsetAsync(g_w, 0, sizeof(float) * n_flows * D, st), "memset");
    VIBO_CUDA(cudaMemsetAsync(g_b, 0, sizeof(float) * n_flows, st), "memset");
    return VIBO_OK;
  }
  Carve ws(workspace, workspace_bytes);
  float* part_g = ws.take<float>((size_t)vibo::sm_count() * 8 * n);
  if (!ws.ok()) return fail(VIBO_ERR_WORKSPACE, "workspace carve overflow");
  VIBO_CUDA(vibo::launch_flow_person_backward(desc->num_person, D, n_flows, ability_mu, ability_logvar, eps,
                                              uhat, w, b, g_ability_k, g_term, g_ability_mu, g_ability_logvar,
                                              g_uhat, g_w, g_b, part_g, st),
            "flow_person_backward");
  return VIBO_OK;
}

int vibo_planar_params_forward(int n_flows, int dim, const float* const* u, const float* const* w,
                               const float* const* b, float* uhat, float* w_out, float* b_out, void* stream) {
  if (n_flows < 1 || n_flows > 8 || dim < 1 || dim > VIBO_MAX_ABILITY_DIM)
    return fail(VIBO_ERR_UNSUPPORTED, "n_flows must be in 1..8 and dim in 1..8");
  if (!u || !w || !b || !uhat || !w_out || !b_out) return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  for (int k = 0; k < n_flows; ++k)
    if (!u[k] || !w[k] || !b[k]) return fail(VIBO_ERR_BAD_ARGUMENT, "NULL parameter pointer");
  VIBO_CUDA(vibo::launch_planar_params_forward(n_flows, dim, u, w, b, uhat, w_out, b_out,
                                               static_cast<cudaStream_t>(stream)),
            "planar_params_forward");
  return VIBO_OK;
}

int vibo_planar_params_backward(int n_flows, int dim, const float* const* u, const float* const* w,
                                const float* g_uhat, const float* g_w_out, const float* g_b_out, float* g_u,
                                float* g_w, float* g_b, void* stream) {
  if (n_flows < 1 || n_flows > 8 || dim < 1 || dim > VIBO_MAX_ABILITY_DIM)
    return fail(VIBO_ERR_UNSUPPORTED, "n_flows must be in 1..8 and dim in 1..8");
  if (!u || !w || !g_uhat || !g_w_out || !g_b_out || !g_u || !g_w || !g_b)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  for (int k = 0; k < n_flows; ++k)
    if (!u[k] || !w[k]) return fail(VIBO_ERR_BAD_ARGUMENT, "NULL parameter pointer");
  VIBO_CUDA(vibo::launch_planar_params_backward(n_flows, dim, u, w, g_uhat, g_w_out, g_b_out, g_u, g_w, g_b,
                                                static_cast<cudaStream_t>(stream)),
            "planar_params_backward");
  return VIBO_OK;
}

int vibo_param_forward(const vibo_desc* desc, int hidden_dim, const float* mu_lookup,
                       const float* logvar_lookup, const float* eps_item, const float* w0,
                       const float* b0, const float* w2, const float* b2, const float* w4,
                       const float* b4, float* item_feat, float* table, float* hidden,
                       double* item_term, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (desc->conditional) return fail(VIBO_ERR_UNSUPPORTED, "vibo_param_forward covers the unconditional encoder only");
  if (hidden_dim < 1 || hidden_dim > 256) return fail(VIBO_ERR_UNSUPPORTED, "hidden_dim must be in 1..256");
  if (!mu_lookup || !logvar_lookup || !eps_item || !w0 || !b0 || !w2 || !b2 || !w4 || !b4 || !item_feat ||
      !table || !hidden || !item_term)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  const int F = vibo::item_width_host(desc->irt_model, desc->ability_dim);
  VIBO_CUDA(vibo::launch_param_forward(desc->num_item, F, desc->ability_dim, hidden_dim, desc->elbo_form,
                                       mu_lookup, logvar_lookup, eps_item, w0, b0, w2, b2, w4, b4, item_feat,
                                       table, hidden, item_term, static_cast<cudaStream_t>(stream)),
            "param_forward");
  return VIBO_OK;
}

int vibo_percell_mlp(int64_t num_person, int num_item, int hidden_dim, int u_rows, int v_rows, const float* u,
                     const float* v, const float* z, const float* w0, const float* w2, const float* c2,
                     const float* w4, float c4, float* out, void* stream) {
  if (num_person < 0 || num_item <= 0) return fail(VIBO_ERR_BAD_ARGUMENT, "num_person must be >= 0 and num_item > 0");
  if (hidden_dim != 64) return fail(VIBO_ERR_UNSUPPORTED, "vibo_percell_mlp is built for hidden_dim 64");
  if ((u_rows != 1 && u_rows != num_item) || (v_rows != 1 && v_rows != num_person))
    return fail(VIBO_ERR_BAD_ARGUMENT, "u_rows must be 1 or num_item, v_rows 1 or num_person");
  if (!u || !v || !w2 || !c2 || !w4 || !out || ((z == nullptr) != (w0 == nullptr)))
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer (z and w0 come together)");
  if ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(v)) & 15)
    return fail(VIBO_ERR_MISALIGNED, "u and v must be 16-byte aligned");
  if (num_person == 0) return VIBO_OK;
  VIBO_CUDA(vibo::launch_percell_mlp(num_person, num_item, u_rows, v_rows, u, v, z, w0, w2, c2, w4, c4, out,
                                     static_cast<cudaStream_t>(stream)),
            "percell_mlp");
  return VIBO_OK;
}

size_t vibo_log_marginal_workspace_bytes(int num_samples) {
  return num_samples > 0 ? vibo::log_marginal_workspace_bytes(num_samples) : 0;
}

int vibo_log_marginal(const vibo_desc* desc, const float* response, const uint8_t* mask, const float* table,
                      const float* item_mu, const float* item_logvar, int num_samples,
                      const float* eps_item, const float* eps_ability, uint64_t seed,
                      const uint64_t* seed_state, double* out_log_weights, double* out_logp,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (desc->conditional)
    return fail(VIBO_ERR_UNSUPPORTED, "vibo_log_marginal covers the unconditional posterior (the table does not "
                                      "depend on the item sample); compose per-sample calls otherwise");
  if (num_samples < 1 || num_samples > 65536) return fail(VIBO_ERR_BAD_ARGUMENT, "num_samples must be in 1..65536");
  if (!response || !mask || !table || !item_mu || !item_logvar || !out_log_weights || !out_logp)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (int rc = check_items(desc)) return rc;
  if (desc->num_person == 0) return fail(VIBO_ERR_BAD_ARGUMENT, "num_person must be > 0");
  if (workspace == nullptr || workspace_bytes < vibo::log_marginal_workspace_bytes(num_samples))
    return fail(VIBO_ERR_WORKSPACE, "workspace smaller than vibo_log_marginal_workspace_bytes(num_samples)");
  VIBO_CUDA(vibo::launch_log_marginal(*desc, num_samples, response, mask, table, item_mu, item_logvar, eps_item,
                                      eps_ability, seed, seed_state, out_log_weights, out_logp, workspace,
                                      workspace_bytes, static_cast<cudaStream_t>(stream)),
            "log_marginal");
  return VIBO_OK;
}

int vibo_predictive_mean(const vibo_desc* desc, const float* ability_mu, const float* ability_logvar,
                         const float* item_mu, const float* item_logvar, int num_samples, uint64_t seed,
                         const uint64_t* seed_state, float* out_mean, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (num_samples < 1) return fail(VIBO_ERR_BAD_ARGUMENT, "num_samples must be >= 1");
  if (!ability_mu || !ability_logvar || !item_mu || !item_logvar || !out_mean)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  if (int rc = check_items(desc)) return rc;
  if (desc->num_person == 0) return VIBO_OK;
  VIBO_CUDA(vibo::launch_predictive_mean(*desc, num_samples, ability_mu, ability_logvar, item_mu, item_logvar, seed,
                                         seed_state, out_mean, static_cast<cudaStream_t>(stream)),
            "predictive_mean");
  return VIBO_OK;
}

int vibo_param_forward_draw(const vibo_desc* desc, int hidden_dim, const float* mu_lookup,
                            const float* logvar_lookup, const uint64_t* seed_state, const float* w0,
                            const float* b0, const float* w2, const float* b2, const float* w4,
                            const float* b4, float* eps_item_out, float* item_feat, float* table,
                            float* hidden, double* item_term, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (desc->conditional) return fail(VIBO_ERR_UNSUPPORTED, "vibo_param_forward_draw covers the unconditional encoder only");
  if (hidden_dim < 1 || hidden_dim > 256) return fail(VIBO_ERR_UNSUPPORTED, "hidden_dim must be in 1..256");
  if (!mu_lookup || !logvar_lookup || !seed_state || !w0 || !b0 || !w2 || !b2 || !w4 || !b4 || !eps_item_out ||
      !item_feat || !table || !hidden || !item_term)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  const int F = vibo::item_width_host(desc->irt_model, desc->ability_dim);
  VIBO_CUDA(vibo::launch_param_forward(desc->num_item, F, desc->ability_dim, hidden_dim, desc->elbo_form,
                                       mu_lookup, logvar_lookup, nullptr, w0, b0, w2, b2, w4, b4, item_feat,
                                       table, hidden, item_term, static_cast<cudaStream_t>(stream), seed_state,
                                       eps_item_out),
            "param_forward");
  return VIBO_OK;
}

int vibo_step_tail(const vibo_desc* desc, int hidden_dim, float beta, float item_scale,
                   const double* scalars, const double* item_term, float* loss_out, int64_t* counter0,
                   int64_t* counter1, const float* mu_lookup, const float* logvar_lookup,
                   const float* eps_item, const float* w2, const float* w4, const float* hidden,
                   const float* g_table, const float* g_item, float* g_mu_lookup,
                   float* g_logvar_lookup, float* g_w0, float* g_b0, float* g_w2, float* g_b2,
                   float* g_w4, float* g_b4, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (desc->conditional) return fail(VIBO_ERR_UNSUPPORTED, "vibo_step_tail covers the unconditional encoder only");
  if (hidden_dim < 1 || hidden_dim > 256) return fail(VIBO_ERR_UNSUPPORTED, "hidden_dim must be in 1..256");
  if (!scalars || !item_term || !loss_out) return fail(VIBO_ERR_BAD_ARGUMENT, "scalars, item_term and loss_out are required");
  const bool grad = g_mu_lookup != nullptr;
  if (grad && (!mu_lookup || !logvar_lookup || !eps_item || !w2 || !w4 || !hidden || !g_table || !g_item ||
               !g_logvar_lookup || !g_w0 || !g_b0 || !g_w2 || !g_b2 || !g_w4 || !g_b4))
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer (all gradient outputs and their inputs, or none)");
  const int F = vibo::item_width_host(desc->irt_model, desc->ability_dim);
  VIBO_CUDA(vibo::launch_step_tail(desc->num_item, F, desc->ability_dim, hidden_dim, desc->elbo_form, beta,
                                   item_scale, scalars, item_term, loss_out, counter0, counter1, grad, mu_lookup,
                                   logvar_lookup, eps_item, w2, w4, hidden, g_table, g_item, g_mu_lookup,
                                   g_logvar_lookup, g_w0, g_b0, g_w2, g_b2, g_w4, g_b4,
                                   static_cast<cudaStream_t>(stream)),
            "step_tail");
  return VIBO_OK;
}

int vibo_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   const int64_t* step, float lr, float beta1, float beta2, float eps, void* stream) {
  if (n < 0 || n > 0x7fffffff) return fail(VIBO_ERR_BAD_ARGUMENT, "n out of range");
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step) return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  VIBO_CUDA(vibo::launch_adam((int)n, param, grad, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps,
                              static_cast<cudaStream_t>(stream)),
            "adam");
  return VIBO_OK;
}

int vibo_param_backward(const vibo_desc* desc, int hidden_dim, const float* mu_lookup,
                        const float* logvar_lookup, const float* eps_item, const float* w2,
                        const float* w4, const float* hidden, const float* g_table,
                        const float* g_item, const float* g_item_term, float* g_mu_lookup,
                        float* g_logvar_lookup, float* g_w0, float* g_b0, float* g_w2, float* g_b2,
                        float* g_w4, float* g_b4, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (desc->conditional) return fail(VIBO_ERR_UNSUPPORTED, "vibo_param_backward covers the unconditional encoder only");
  if (hidden_dim < 1 || hidden_dim > 256) return fail(VIBO_ERR_UNSUPPORTED, "hidden_dim must be in 1..256");
  if (!mu_lookup || !logvar_lookup || !eps_item || !w2 || !w4 || !hidden || !g_table || !g_item ||
      !g_item_term || !g_mu_lookup || !g_logvar_lookup || !g_w0 || !g_b0 || !g_w2 || !g_b2 || !g_w4 || !g_b4)
    return fail(VIBO_ERR_BAD_ARGUMENT, "NULL pointer");
  const int F = vibo::item_width_host(desc->irt_model, desc->ability_dim);
  VIBO_CUDA(vibo::launch_param_backward(desc->num_item, F, desc->ability_dim, hidden_dim, desc->elbo_form,
                                        mu_lookup, logvar_lookup, eps_item, w2, w4, hidden, g_table, g_item,
                                        g_item_term, g_mu_lookup, g_logvar_lookup, g_w0, g_b0, g_w2, g_b2,
                                        g_w4, g_b4, static_cast<cudaStream_t>(stream)),
            "param_backward");
  return VIBO_OK;
}

}  // extern "C"
