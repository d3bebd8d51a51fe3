// Internal launcher declarations shared by the .cu translation units.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/vibo_b200.h"

namespace vibo {

int sm_count();
void note_launch(int n = 1);  // counts kernel launches (vibo_launch_count)
int item_width_host(int model, int D);
int set_last_error(int code, const char* msg);  // records vibo_last_error() for this thread, returns code
int general_max_items(int D);
int general_grid(int64_t P, int I, int D);
int bernoulli_grid(int64_t n);

// vibo_general.cu
cudaError_t launch_encode(const vibo_desc& d, const float* resp, const uint8_t* mask,
                          const float* table, float* mu, float* lv, float* S, cudaStream_t st);
cudaError_t launch_encode_bwd(const vibo_desc& d, const float* resp, const uint8_t* mask,
                              const float* table, const float* amu, const float* S,
                              const float* g_mu, const float* g_lv, float* g_table, float* part,
                              cudaStream_t st);
// Unconditional posterior: the forward also returns the per-person counts (n1, n_observed), the backward
// works from them alone (no second pass over the rows).  cudaErrorNotSupported: use the pair above.
// narrow rows (I <= 256), unconditional: one warp per row with direct loads; mu / counts may each be NULL
cudaError_t launch_encode_rows_uncond(const vibo_desc& d, const float* resp, const uint8_t* mask,
                                      const float* table, float* mu, float* lv, float* S, float* counts,
                                      cudaStream_t st);
cudaError_t launch_encode_counts(const vibo_desc& d, const float* resp, const uint8_t* mask,
                                 const float* table, float* mu, float* lv, float* S, float* counts,
                                 cudaStream_t st);
cudaError_t launch_encode_bwd_counts(const vibo_desc& d, const float* counts, const float* table,
                                     const float* amu, const float* S, const float* g_mu, const float* g_lv,
                                     float* g_table, float* part, cudaStream_t st);
cudaError_t launch_link(const vibo_desc& d, const float* resp, const uint8_t* mask,
                        const float* ability, const float* item_feat, double* out_ll,
                        float* g_ability, float* g_item, double* part_ll, float* part_g,
                        cudaStream_t st);
cudaError_t launch_decode(const vibo_desc& d, const float* ability, const float* item_feat,
                          float* out, cudaStream_t st);
cudaError_t launch_bernoulli_ll(const vibo_desc& d, const float* resp, const uint8_t* mask,
                                const float* prob, double* out_ll, float* g_prob, double* part_ll,
                                cudaStream_t st);

// vibo_stream.cu: slab-stream kernels (TMA-staged row tiles).  Return
// cudaErrorNotSupported when the configuration / pointer alignment is not
// covered; the launchers above then fall back to the legacy kernels.
cudaError_t stream_encode(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* table,
                          float* mu, float* lv, float* S, int* grid_out, cudaStream_t st);
cudaError_t stream_counts(const vibo_desc& d, const float* resp, const uint8_t* mask, float* counts,
                          cudaStream_t st);
cudaError_t stream_encode_counts(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* table,
                                 float* mu, float* lv, float* S, float* counts, cudaStream_t st);
cudaError_t launch_person_counts(const vibo_desc& d, const float* resp, const uint8_t* mask, float* counts,
                                 cudaStream_t st);
cudaError_t stream_link(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* ability,
                        const float* item_feat, double* part_ll, float* g_ability, float* part_g, bool grad,
                        int* grid_out, cudaStream_t st);
cudaError_t stream_encode_bwd(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* amu,
                              const float* S, const float* g_mu, const float* g_lv, float* part, int* grid_out,
                              cudaStream_t st);

// vibo_person.cu: per-person prior / reparameterisation math of the
// multi-pass composition of vibo_fused_elbo.
int person_grid(int64_t P);
cudaError_t launch_person_forward(const vibo_desc& d, const float* amu, const float* alv,
                                  const float* eps_or_null, uint64_t seed, const uint64_t* seed_dev,
                                  float* eps_out, float* ability, double* part_term, double* out_term,
                                  cudaStream_t st);
cudaError_t launch_person_backward(const vibo_desc& d, float beta, const float* amu,
                                   const float* alv, const float* eps, const float* ability,
                                   const float* g_ll_ability, float* g_mu, float* g_lv,
                                   cudaStream_t st);
cudaError_t launch_philox_fill(int64_t P, int D, int64_t person_offset, uint64_t seed,
                               const uint64_t* seed_dev, float* eps, cudaStream_t st);
cudaError_t launch_negate(float* v, int n, cudaStream_t st);
int flow_grid(int64_t P);
cudaError_t launch_flow_person_forward(int64_t P, int D, int K, const float* amu, const float* alv,
                                       const float* eps, const float* uhat, const float* w, const float* b,
                                       float* ability0, float* ability_k, double* part_term, double* out_term,
                                       cudaStream_t st);
cudaError_t launch_planar_params_forward(int K, int D, const float* const* u, const float* const* w,
                                         const float* const* b, float* uhat, float* w_out, float* b_out,
                                         cudaStream_t st);
cudaError_t launch_planar_params_backward(int K, int D, const float* const* u, const float* const* w,
                                          const float* g_uhat, const float* g_w_out, const float* g_b_out,
                                          float* g_u, float* g_w, float* g_b, cudaStream_t st);
cudaError_t launch_flow_person_backward(int64_t P, int D, int K, const float* amu, const float* alv,
                                        const float* eps, const float* uhat, const float* w, const float* b,
                                        const float* g_ability_k, const float* g_term, float* g_mu, float* g_lv,
                                        float* g_uhat, float* g_w, float* g_b, float* part_g, cudaStream_t st);
cudaError_t launch_accumulate(float* dst, const float* src, int n, double* dst2, const double* src2,
                              int n2, cudaStream_t st);

// vibo_param.cu: parameter-side chain of the unconditional model
cudaError_t launch_param_forward(int I, int F, int D, int H, int form, const float* mu, const float* lv,
                                 const float* eps, const float* w0, const float* b0, const float* w2,
                                 const float* b2, const float* w4, const float* b4, float* item_feat,
                                 float* table, float* hidden, double* item_term, cudaStream_t st,
                                 const uint64_t* seed_state = nullptr, float* eps_out = nullptr);
cudaError_t launch_step_tail(int I, int F, int D, int H, int form, float beta, float item_scale,
                             const double* scalars, const double* item_term, float* loss_out, int64_t* counter0,
                             int64_t* counter1, bool grad, const float* mu, const float* lv, const float* eps,
                             const float* w2, const float* w4, const float* hidden, const float* g_table,
                             const float* g_item, float* g_mu, float* g_lv, float* g_w0, float* g_b0,
                             float* g_w2, float* g_b2, float* g_w4, float* g_b4, cudaStream_t st);
cudaError_t launch_adam(int n, float* param, const float* grad, float* m, float* v, const int64_t* step, float lr,
                        float b1, float b2, float eps, cudaStream_t st);
cudaError_t launch_param_backward(int I, int F, int D, int H, int form, const float* mu, const float* lv,
                                  const float* eps, const float* w2, const float* w4, const float* hidden,
                                  const float* g_table, const float* g_item, const float* g_term,
                                  float* g_mu, float* g_lv, float* g_w0, float* g_b0, float* g_w2,
                                  float* g_b2, float* g_w4, float* g_b4, cudaStream_t st);

// vibo_tc5_encode.cu: conditional encode on tcgen05 (TF32, 2-D TMA, TMEM accumulators)
cudaError_t tc5_encode(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* table, float* mu,
                       float* lv, float* S, cudaStream_t st);

cudaError_t tc5_encode_bwd(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* amu,
                           const float* S, const float* g_mu, const float* g_lv, float* part, int* grid_out,
                           cudaStream_t st);

// vibo_tc5_eval.cu: single-pass forward ELBO of the conditional posterior (tcgen05 encode + on-chip link)
size_t tc5_eval_workspace_bytes();
cudaError_t tc5_eval(const vibo_desc& d, const float* resp, const uint8_t* mask, const float* table,
                     const float* item_feat, const float* eps, uint64_t seed, const uint64_t* seed_dev,
                     double* out_scalars, float* amu, float* alv, float* ability, void* ws, size_t ws_bytes,
                     cudaStream_t st);

// vibo_percell.cu: per-cell MLP of the nonlinear generative models on tcgen05 / TMEM (hidden width 64)
cudaError_t launch_percell_mlp(int64_t P, int I, int u_rows, int v_rows, const float* U, const float* V,
                               const float* Z, const float* w0, const float* W2, const float* c2, const float* w4,
                               float c4, float* out, cudaStream_t st);

// vibo_pack.cu: one-byte-per-cell row format
cudaError_t launch_pack(int64_t n, const float* resp, const uint8_t* mask, int8_t* out, cudaStream_t st);
cudaError_t launch_unpack(int64_t n, const int8_t* in, float* resp, uint8_t* mask, cudaStream_t st);

// vibo_sample.cu: sample-loop kernels (IWAE log-marginal, posterior-predictive mean)
size_t log_marginal_workspace_bytes(int S);
cudaError_t launch_log_marginal(const vibo_desc& d, int S, const float* resp, const uint8_t* mask, const float* table,
                                const float* item_mu, const float* item_lv, const float* eps_item,
                                const float* eps_ability, uint64_t seed, const uint64_t* seed_dev, double* log_w,
                                double* out_logp, void* ws, size_t ws_bytes, cudaStream_t st);
cudaError_t launch_predictive_mean(const vibo_desc& d, int S, const float* amu, const float* alv, const float* item_mu,
                                   const float* item_lv, uint64_t seed, const uint64_t* seed_dev, float* out_mean,
                                   cudaStream_t st);

// vibo_fused.cu: single-pass kernel.  Returns false when the configuration is
// not covered (caller composes the general kernels instead).
void profile_begin();
int profile_end(int* n, double* total_ms);
unsigned long long launch_count();
bool fused_supported(const vibo_desc& d, const float* resp, const uint8_t* mask);
size_t fused_workspace_bytes(const vibo_desc& d);
cudaError_t launch_fused(const vibo_desc& d, const float* resp, const uint8_t* mask,
                         const float* table, const float* item_feat, const float* eps,
                         uint64_t seed, const uint64_t* seed_dev, float beta, double* out_scalars,
                         float* amu, float* alv, float* ability, float* g_table, float* g_item, void* ws,
                         size_t ws_bytes, bool accumulate, cudaStream_t st);

}  // namespace vibo
