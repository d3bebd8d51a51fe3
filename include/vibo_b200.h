/*
 * vibo_b200.h -- C ABI of the B200-native VIBO ELBO engine (libvibo_b200.so).
 *
 * Drop-in boundary for ONE hot path of mhw32/variational-item-response-theory:
 * the amortized ELBO forward/backward of src/torch_core/models.py
 * (VIBO_{1,2,3}PL.forward :337-354, .encode :356-371, .decode :373-378,
 * .elbo :380-443) together with the helpers of src/utils.py (:46-49, :59-67,
 * :85-88, :105-113) it calls.  The reference has no FFI of its own (100 %
 * Python on stock PyTorch ops); these entry points are what a ctypes binding
 * for that path binds -- see INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch / C++ types cross the boundary.
 *   - Every pointer is a DEVICE pointer on the current CUDA device unless the
 *     parameter name ends in _host.  All buffers are caller-owned and
 *     contiguous row-major; the library allocates nothing persistent.
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*); no entry
 *     point synchronises except the *_host ones, which say so.
 *   - Return value: 0 on success, a negative vibo_status otherwise; the
 *     message for the calling thread's last failure is vibo_last_error().
 *   - P persons (rows), I items (columns), D ability dims, F item-feature
 *     width: 1 (1PL), D+1 (2PL), D+2 (3PL)   [models.py:331, :523, :538].
 *   - response: float32 (P, I), values 0/1; missing cells hold anything
 *     (-1 by convention, src/config.py:14).   mask: uint8/bool (P, I),
 *     non-zero = observed  [src/datasets.py:928-940 emits exactly these].
 *   - table: float32 (2, It, 2D): raw encoder outputs (mean | log-variance)
 *     of AbilityInferenceNetwork.mlp (models.py:575-582, :599) for response
 *     value r = 0, 1; It = 1 for the unconditional encoder, It = I for
 *     ConditionalAbilityInferenceNetwork (models.py:695-710).
 */
#ifndef VIBO_B200_H_
#define VIBO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIBO_B200_VERSION 100 /* 0.1.0 */
#define VIBO_MAX_ABILITY_DIM 8

typedef enum {
  VIBO_OK = 0,
  VIBO_ERR_BAD_ARGUMENT = -1, /* null pointer, negative size, bad enum      */
  VIBO_ERR_UNSUPPORTED = -2,  /* D > VIBO_MAX_ABILITY_DIM, I too large ...  */
  VIBO_ERR_MISALIGNED = -3,   /* pointer not aligned as documented          */
  VIBO_ERR_WORKSPACE = -4,    /* workspace too small                        */
  VIBO_ERR_CUDA = -5          /* a CUDA runtime call failed                 */
} vibo_status;

/* replace_missing_with_prior (models.py:614-618) vs --drop-missing */
#define VIBO_MISSING_PRIOR 0
#define VIBO_MISSING_DROP 1
/* elbo(): use_kl_divergence=True (models.py:427-430) vs False (:432-441) */
#define VIBO_ELBO_KL 0
#define VIBO_ELBO_SAMPLE 1

typedef struct {
  int64_t num_person;     /* P: rows held by this call (one shard)            */
  int32_t num_item;       /* I                                                */
  int32_t ability_dim;    /* D, 1..VIBO_MAX_ABILITY_DIM                       */
  int32_t irt_model;      /* 1, 2, 3  (VIBO_1PL / 2PL / 3PL)                  */
  int32_t conditional;    /* 0: table is (2,1,2D); 1: table is (2,I,2D)       */
  int32_t missing_policy; /* VIBO_MISSING_*                                   */
  int32_t elbo_form;      /* VIBO_ELBO_*                                      */
  int64_t person_offset;  /* global index of row 0 (keys the in-kernel Philox) */
} vibo_desc;

int vibo_version(void);
const char* vibo_last_error(void);

/* Bytes of scratch the calls below need for `desc` (an upper bound valid for
 * every entry point).  The scratch holds per-CTA partial sums; it need not be
 * zeroed. */
size_t vibo_workspace_bytes(const vibo_desc* desc);

/*
 * Fused ELBO: replaces, in ONE pass over the response matrix,
 *   encode   models.py:364-368   (PoE ability posterior + reparameterised draw)
 *   decode   models.py:373, :729-766 (IRT link)
 *   elbo     models.py:399 (masked Bernoulli LL, utils.py:46-49) and the
 *            per-person prior term :428 (KL, utils.py:85-88) or :433-435.
 * response_mu is never materialised.
 *
 *   eps_ability  (P, D) standard-normal draws (models.py:509), or NULL to
 *                draw them in-kernel with Philox4x32-10 keyed by
 *                (seed, person_offset + row).
 *   out_scalars  double[2]: { LL = sum_ij o_ij ll_ij ,  person_term } where
 *                person_term = KL(q(theta)||N(0,1)) summed over persons
 *                (VIBO_ELBO_KL) or sum_i log p(theta_i) - log q(theta_i)
 *                (VIBO_ELBO_SAMPLE).
 *   ability_mu, ability_logvar, ability   (P, D) each, or NULL to skip.
 *   g_table (2, It, 2D), g_item (I, F): gradients of
 *                loss_k = -LL + beta * KL                 (VIBO_ELBO_KL)
 *                loss_k = -LL - person_term               (VIBO_ELBO_SAMPLE)
 *                w.r.t. table and item_feat (link path only); both NULL for a
 *                forward-only evaluation.  Overwritten, not accumulated.
 */
int vibo_fused_elbo(const vibo_desc* desc, const float* response, const uint8_t* mask,
                    const float* table, const float* item_feat, const float* eps_ability,
                    uint64_t seed, float beta, double* out_scalars, float* ability_mu,
                    float* ability_logvar, float* ability, float* g_table, float* g_item,
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * vibo_fused_elbo for CUDA-graph replays: the reparameterisation noise is always drawn in-kernel
 * and the Philox key is read FROM DEVICE MEMORY when the kernel runs:
 *   key = seed_state[0] + seed_state[1]      (device uint64[2] = {seed, step})
 * so a captured step draws fresh noise on every replay once the caller bumps seed_state[1]
 * (the role torch's global generator plays for models.py:509 in the reference loop,
 * vibo.py:243-268), still keyed by the global person index.
 */
int vibo_fused_elbo_graph(const vibo_desc* desc, const float* response, const uint8_t* mask,
                          const float* table, const float* item_feat, const uint64_t* seed_state,
                          float beta, double* out_scalars, float* ability_mu, float* ability_logvar,
                          float* ability, float* g_table, float* g_item, void* workspace,
                          size_t workspace_bytes, void* stream);

/* The (P, D) standard normals vibo_fused_elbo draws in-kernel when eps_ability is NULL, as a
 * tensor: Philox4x32-10, counter = (person_offset + row, block), key = seed, or
 * seed_state[0] + seed_state[1] read on the device when seed_state is non-NULL.  Used by the
 * composed paths (flows, mean merge) so that their noise is keyed by person too. */
int vibo_philox_normal(const vibo_desc* desc, uint64_t seed, const uint64_t* seed_state, float* eps,
                       void* stream);

/*
 * Same computation with response / mask in HOST memory (pinned or pageable):
 * rows are streamed host->device in person chunks on an internal copy stream,
 * overlapped with the kernel on the previous chunk; results (out_scalars and
 * the gradient buffers) are DEVICE buffers as above and `out_scalars_host`
 * (double[2], may be NULL) receives a copy.  Synchronises `stream` before
 * returning.  staging: device scratch of vibo_host_staging_bytes() bytes.
 */
size_t vibo_host_staging_bytes(const vibo_desc* desc, int64_t chunk_person);
/*
 * Host-compressed route of vibo_fused_elbo_host: for chunks of >= 2^20 cells, a share of every chunk's
 * rows (vibo_host_pack_share(): starts at 0.72 and follows the machine, VIBO_HOST_PACK_FRACTION fixes it,
 * 0 with fewer than 6 host threads per process) is packed to 1 B/cell by the library's host thread pool while the rest of the chunk
 * is in flight over PCIe in the reference layout, then sent and expanded on the device.  The two routes
 * use different resources (CPU cores + host DRAM vs the PCIe link), so their throughputs add; results are
 * identical.  Requires 0 / 1 responses where observed (the Bernoulli model of this path).
 */
double vibo_host_pack_share(const vibo_desc* desc, int64_t chunk_person);
int vibo_fused_elbo_host(const vibo_desc* desc, const float* response_host,
                         const uint8_t* mask_host, const float* table, const float* item_feat,
                         const float* eps_ability, uint64_t seed, float beta,
                         double* out_scalars, double* out_scalars_host, float* g_table,
                         float* g_item, int64_t chunk_person, void* staging,
                         size_t staging_bytes, void* workspace, size_t workspace_bytes,
                         void* stream);

/*
 * Packed row format: one signed byte per cell, -1 = missing (src/config.py:14), 0 / 1 = observed
 * response -- the (response, mask) pair of src/datasets.py:928-940 in 1 byte instead of 5.
 *   vibo_pack / vibo_unpack   device-side conversion between the pair and the packed rows (P, I).
 *   vibo_pack_host            the same conversion of HOST rows on the caller's CPU cores (a pool of
 *       vibo_host_threads() threads: hardware concurrency, VIBO_HOST_THREADS overrides) -- the
 *       one-off step at dataset load for the packed entry point below.
 *   vibo_fused_elbo_host_packed   vibo_fused_elbo_host with the rows in HOST memory in the packed
 *       format: PCIe carries 1 B/cell; each chunk is expanded on the device (one streaming kernel)
 *       before the row kernels read it.  Same staging size query (vibo_host_staging_bytes).
 */
int vibo_pack(const vibo_desc* desc, const float* response, const uint8_t* mask, int8_t* packed, void* stream);
int vibo_pack_host(const vibo_desc* desc, const float* response_host, const uint8_t* mask_host,
                   int8_t* packed_host);
int vibo_host_threads(void);
int vibo_unpack(const vibo_desc* desc, const int8_t* packed, float* response, uint8_t* mask, void* stream);
int vibo_fused_elbo_host_packed(const vibo_desc* desc, const int8_t* packed_host, const float* table,
                                const float* item_feat, const float* eps_ability, uint64_t seed,
                                float beta, double* out_scalars, double* out_scalars_host,
                                float* g_table, float* g_item, int64_t chunk_person, void* staging,
                                size_t staging_bytes, void* workspace, size_t workspace_bytes,
                                void* stream);

/*
 * encode: product-of-experts ability posterior only.
 * Replaces AbilityInferenceNetwork._forward_product (models.py:596-629) +
 * product_of_experts (utils.py:105-113).  precision_sum (P, D) = sum of expert
 * precisions (saved for the backward).
 */
int vibo_encode(const vibo_desc* desc, const float* response, const uint8_t* mask,
                const float* table, float* ability_mu, float* ability_logvar,
                float* precision_sum, void* stream);

/*
 * Backward of vibo_encode: given d loss / d ability_mu and d loss /
 * d ability_logvar (P, D), writes d loss / d table (2, It, 2D).
 */
int vibo_encode_backward(const vibo_desc* desc, const float* response, const uint8_t* mask,
                         const float* table, const float* ability_mu,
                         const float* precision_sum, const float* g_ability_mu,
                         const float* g_ability_logvar, float* g_table, void* workspace,
                         size_t workspace_bytes, void* stream);

/*
 * The same pair for the UNCONDITIONAL posterior with one pass over the rows instead of two: the
 * unconditional table has one expert per response value (models.py:575-582 evaluated on r in {0, 1}), so
 * the posterior and its backward depend on a row only through counts (P, 2) = (observed ones, observed
 * cells).  vibo_encode_counts writes them next to the posterior; vibo_encode_backward_counts forms
 * d loss / d table (2, 1, 2D) from them without reading response / mask again.
 * VIBO_ERR_UNSUPPORTED for a conditional posterior or rows that are not 16-byte aligned.
 */
int vibo_encode_counts(const vibo_desc* desc, const float* response, const uint8_t* mask,
                       const float* table, float* ability_mu, float* ability_logvar,
                       float* precision_sum, float* counts, void* stream);
int vibo_encode_backward_counts(const vibo_desc* desc, const float* counts, const float* table,
                                const float* ability_mu, const float* precision_sum,
                                const float* g_ability_mu, const float* g_ability_logvar, float* g_table,
                                void* workspace, size_t workspace_bytes, void* stream);

/*
 * link + log-likelihood without materialising response_mu:
 *   LL = sum_ij o_ij log Bernoulli(x_ij ; irt_model(ability, item_feat))
 * (models.py:729-766 + utils.py:46-49 + the .sum() of models.py:399).
 * out_ll: double[1].  g_ability (P, D), g_item (I, F): d LL / d ability and
 * d LL / d item_feat, or both NULL.
 */
int vibo_link_loglik(const vibo_desc* desc, const float* response, const uint8_t* mask,
                     const float* ability, const float* item_feat, double* out_ll,
                     float* g_ability, float* g_item, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Per-person sufficient statistics of an unconditional encoder: counts (P, 2) =
 * (number of observed responses equal to 1, number of observed responses).  The masked mean
 * of `--ability-merge mean` (models.py:631-650) and the unconditional product of experts
 * (models.py:596-629) depend on a row only through these counts. */
int vibo_person_counts(const vibo_desc* desc, const float* response, const uint8_t* mask, float* counts,
                       void* stream);

/* decode: response_mu (P, I) = irt_model_{1,2,3}pl(ability, item_feat),
 * models.py:729-766 (API-parity path; the fused entry never needs it). */
int vibo_decode(const vibo_desc* desc, const float* ability, const float* item_feat,
                float* response_mu, void* stream);

/* masked_bernoulli_log_pdf(...).sum() on a materialised response_mu
 * (utils.py:46-49, models.py:399).  out_ll: double[1];  g_prob (P, I) =
 * d LL / d response_mu or NULL. */
int vibo_bernoulli_loglik(const vibo_desc* desc, const float* response, const uint8_t* mask,
                          const float* response_mu, double* out_ll, float* g_prob,
                          void* workspace, size_t workspace_bytes, void* stream);

/*
 * Parameter-side chain of the UNCONDITIONAL model (conditional == 0), i.e. the
 * tiny tensors around the row kernels, as two launches:
 *
 *   forward   item_feat (I, F) = mu_lookup + exp(logvar_lookup / 2) * eps_item
 *                                             [models.py:359-361, :506-510]
 *             table (2, 1, 2D) = ability_encoder.mlp on the two possible cell
 *                                inputs 0 and 1   [models.py:575-582, :599]
 *             item_term (double[1]) = KL(q(item) || N(0,1))  [utils.py:85-88]
 *                  (VIBO_ELBO_KL) or -(log p(item) - log q(item)) [utils.py:59-67]
 *             hidden (2, 2, H): activations kept for the backward.
 *   backward  gradients of every parameter given g_table = d loss / d table,
 *             g_item = d loss / d item_feat (both from vibo_fused_elbo) and
 *             g_item_term = d loss / d item_term (device float[1]).
 *
 * w0 (H, 1), b0 (H), w2 (H, H), b2 (H), w4 (2D, H), b4 (2D) are the
 * ability_encoder.mlp.{0,2,4}.{weight,bias} tensors; hidden_dim <= 256.
 */
int vibo_param_forward(const vibo_desc* desc, int hidden_dim, const float* mu_lookup,
                       const float* logvar_lookup, const float* eps_item, const float* w0,
                       const float* b0, const float* w2, const float* b2, const float* w4,
                       const float* b4, float* item_feat, float* table, float* hidden,
                       double* item_term, void* stream);
int vibo_param_backward(const vibo_desc* desc, int hidden_dim, const float* mu_lookup,
                        const float* logvar_lookup, const float* eps_item, const float* w2,
                        const float* w4, const float* hidden, const float* g_table,
                        const float* g_item, const float* g_item_term, float* g_mu_lookup,
                        float* g_logvar_lookup, float* g_w0, float* g_b0, float* g_w2, float* g_b2,
                        float* g_w4, float* g_b4, void* stream);

/*
 * The training / evaluation step of vibo.py:243-268 (:285-312) for the UNCONDITIONAL model as a
 * handful of launches with no host involvement between them (so the whole step is one CUDA
 * graph): vibo_param_forward_draw -> vibo_fused_elbo_graph -> vibo_step_tail
 * [-> vibo_comm_allreduce] -> vibo_adam_step.
 *
 *   vibo_param_forward_draw  vibo_param_forward with the item noise drawn in-kernel:
 *             eps_item_out (I, F) = Philox(seed_state[0] + seed_state[1]) on the counter range
 *             [2^62, 2^62 + I F) (== vibo_philox_normal with person_offset 2^62, ability_dim 1,
 *             num_person I F): ONE global draw per step, identical on every rank (models.py:361).
 *   vibo_step_tail  loss_out[0] = -LL + beta KL_theta + item_scale beta KL_item  (VIBO_ELBO_KL) or
 *             -LL - person_term + item_scale item_term (VIBO_ELBO_SAMPLE) from out_scalars of the
 *             fused entry and item_term of the forward; item_scale = 1 / world_size when persons
 *             are sharded.  With the gradient pointers given it also runs vibo_param_backward
 *             (g_item_term = item_scale * beta, resp. item_scale).  counter0 / counter1 (device
 *             int64, may be NULL) are incremented by one: seed_state + 1 (the step) and the Adam
 *             step count, after every kernel of the step that reads them.
 *   vibo_adam_step  torch.optim.Adam's update (vibo.py:221; no weight decay, no amsgrad) on flat
 *             buffers; `step` (device int64) is the 1-based index of this update.
 */
int vibo_param_forward_draw(const vibo_desc* desc, int hidden_dim, const float* mu_lookup,
                            const float* logvar_lookup, const uint64_t* seed_state, const float* w0,
                            const float* b0, const float* w2, const float* b2, const float* w4,
                            const float* b4, float* eps_item_out, float* item_feat, float* table,
                            float* hidden, double* item_term, void* stream);
int vibo_step_tail(const vibo_desc* desc, int hidden_dim, float beta, float item_scale,
                   const double* scalars, const double* item_term, float* loss_out, int64_t* counter0,
                   int64_t* counter1, const float* mu_lookup, const float* logvar_lookup,
                   const float* eps_item, const float* w2, const float* w4, const float* hidden,
                   const float* g_table, const float* g_item, float* g_mu_lookup,
                   float* g_logvar_lookup, float* g_w0, float* g_b0, float* g_w2, float* g_b2,
                   float* g_w4, float* g_b4, void* stream);
int vibo_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   const int64_t* step, float lr, float beta1, float beta2, float eps, void* stream);

/*
 * Monte-Carlo estimators of the evaluation closures with the sample loop INSIDE the kernel (each
 * row tile is read once and scored against all samples from shared memory):
 *
 *   vibo_log_marginal   VIBO_*PL.log_marginal (models.py:445-504) for the unconditional posterior:
 *       log w_s = sum_i [ LL_i(theta_is, d_s) + log p(theta_is) - log q(theta_is) ] + log p(d_s) - log q(d_s)
 *       (elbo(..., use_kl_divergence=False), models.py:432-441) with a fresh item draw d_s and fresh
 *       ability draws theta_is per sample;  out_logp[0] = logsumexp_s(log w_s) - log(num_samples),
 *       out_log_weights[s] = log w_s (doubles).  table (2, 1, 2D) as for vibo_fused_elbo;
 *       item_mu / item_logvar (I, F) are item_encoder.{mu,logvar}_lookup.  Noise: eps_item
 *       (S, I, F) and eps_ability (S, P, D) (either may be NULL: Philox with stream id s + 1,
 *       key = seed, or seed_state[0] + seed_state[1] when seed_state is non-NULL; abilities keyed
 *       by person_offset + row).
 *   vibo_predictive_mean   mean over num_samples posterior draws of irt_model(theta_s, d_s)
 *       (sample_posterior_predictive, vibo.py:349-390, followed by .mean(0), :515), given the
 *       ability posterior (P, D) and the item posterior (I, F); out_mean (P, I).
 */
size_t vibo_log_marginal_workspace_bytes(int num_samples);
int vibo_log_marginal(const vibo_desc* desc, const float* response, const uint8_t* mask, const float* table,
                      const float* item_mu, const float* item_logvar, int num_samples,
                      const float* eps_item, const float* eps_ability, uint64_t seed,
                      const uint64_t* seed_state, double* out_log_weights, double* out_logp,
                      void* workspace, size_t workspace_bytes, void* stream);
int vibo_predictive_mean(const vibo_desc* desc, const float* ability_mu, const float* ability_logvar,
                         const float* item_mu, const float* item_logvar, int num_samples, uint64_t seed,
                         const uint64_t* seed_state, float* out_mean, void* stream);

/*
 * Per-cell MLP of the nonlinear generative models (--generative-model link | deep | residual;
 * LinkedIRT / DeepIRT / ResidualIRT, models.py:769-919) on the tcgen05 tensor cores: for every
 * cell (person i, item j)
 *     out_ij = w4 . ELU( W2 ELU( u_j + v_i + z_ij w0 ) + c2 ) + c4,        hidden_dim = 64
 * u (u_rows, 64) with u_rows = num_item or 1 (one row broadcast to every item), v (v_rows, 64)
 * likewise per person, z (P, I) with w0 (64) or both NULL; W2 (64, 64) row-major [out][in]
 * (nn.Linear.weight), c2, w4 (64); out (P, I).  link: u = link.0.bias, z = the IRT logit,
 * w0 = link.0.weight[:, 0]; deep / residual: u = mlp_concat.0 on the item half (+ bias),
 * v = mlp_concat.0 on the ability half.  The hidden activations are carried as bf16 hi + lo pairs
 * (three tcgen05.mma products, fp32 accumulation in TMEM): relative error ~1e-5.  Forward only
 * (evaluation / predictive sampling); training differentiates through the same math in PyTorch.
 */
int vibo_percell_mlp(int64_t num_person, int num_item, int hidden_dim, int u_rows, int v_rows, const float* u,
                     const float* v, const float* z, const float* w0, const float* w2, const float* c2,
                     const float* w4, float c4, float* out, void* stream);

/*
 * Planar normalizing flows on the abilities (--n-norm-flows K), per person and
 * fused with the reparameterised draw and the person-side terms of the flow
 * form of the ELBO [flows.py:21-41, :58-66; models.py:342-348, :406-424]:
 *
 *   theta_0 = ability_mu + eps * exp(ability_logvar / 2)             [models.py:506-510]
 *   theta_k = theta_{k-1} + uhat_k tanh(w_k . theta_{k-1} + b_k)     [flows.py:21-41]
 *   ldj_k   = log(|1 + (1 - tanh^2) (w_k . uhat_k)| + 1e-8)
 *   out_term (double[1]) = sum_i [ log N(theta_K; 0, 1) - log N(theta_0; mu, exp logvar)
 *                                  + sum_k ldj_k ]                   [models.py:412-424]
 *
 * uhat (K, D) is the invertibility-corrected u of flows.py:26-29, formed by the
 * caller from (u, w) (parameter-only arithmetic); w (K, D), b (K).  ability_0
 * (P, D) may be NULL.  K <= 8.  The backward returns d loss / d (ability_mu,
 * ability_logvar) (P, D) and d loss / d (uhat, w, b) given g_ability_k =
 * d loss / d theta_K (P, D) and g_term = d loss / d out_term (device float[1]).
 */
int vibo_flow_person_forward(const vibo_desc* desc, int n_flows, const float* ability_mu,
                             const float* ability_logvar, const float* eps, const float* uhat,
                             const float* w, const float* b, float* ability_0, float* ability_k,
                             double* out_term, void* workspace, size_t workspace_bytes, void* stream);
int vibo_flow_person_backward(const vibo_desc* desc, int n_flows, const float* ability_mu,
                              const float* ability_logvar, const float* eps, const float* uhat,
                              const float* w, const float* b, const float* g_ability_k,
                              const float* g_term, float* g_ability_mu, float* g_ability_logvar,
                              float* g_uhat, float* g_w, float* g_b, void* workspace,
                              size_t workspace_bytes, void* stream);

/*
 * Planar-flow parameter chain.  The reference keeps each flow's u (D), w (D), b (1) as separate
 * parameters (flows.py:14-17, state_dict keys flows.{k}.{u,w,b}) and corrects u for invertibility inside
 * PlanarFlow.forward (flows.py:26-29):  uhat = u + (softplus(w.u) - 1 - w.u) w / |w|^2.
 * vibo_planar_params_forward gathers the K flows' parameters (u, w, b: HOST arrays of K DEVICE pointers) into
 * the stacked (uhat, w_out (K, D), b_out (K)) the per-row flow kernels take; the backward maps
 * d loss / d (uhat, w_out, b_out) to d loss / d (u, w, b) as stacked (K, D), (K, D), (K) rows.
 * One launch each way instead of ~13 / ~25 elementwise launches of the autograd formulation.
 */
int vibo_planar_params_forward(int n_flows, int dim, const float* const* u, const float* const* w,
                               const float* const* b, float* uhat, float* w_out, float* b_out, void* stream);
int vibo_planar_params_backward(int n_flows, int dim, const float* const* u, const float* const* w,
                                const float* g_uhat, const float* g_w_out, const float* g_b_out, float* g_u,
                                float* g_w, float* g_b, void* stream);

/*
 * Per-step exchange of the person-sharded run (one process per GPU, one node): in-place SUM over
 * the ranks of a small float buffer ([loss | parameter gradients], the `optimizer.step()` input of
 * vibo.py:266-268 when persons are split over GPUs).  One kernel over NVLink peer memory (CUDA
 * IPC), capturable in a CUDA graph; the sum runs in rank order, so all ranks get identical bits.
 *
 *   create    allocates this rank's exchange region (device of the calling thread) sized for
 *             max_floats and writes its VIBO_COMM_HANDLE_BYTES-byte IPC handle to handle_out;
 *   connect   all_handles = the world_size handles in rank order (exchange them with any host
 *             collective, e.g. torch.distributed.all_gather_object); barrier afterwards;
 *   allreduce every rank must issue the same sequence of calls (same n); enqueued on `stream`;
 *   allreduce_adam  the same exchange followed, in the same kernel, by vibo_adam_step on the summed
 *             gradients: data[skip + k] is the gradient of param[k], k < n - skip (the training step
 *             keeps [loss | gradients] in one vector, skip = 1); `step` as for vibo_adam_step;
 *   status    VIBO_OK, or VIBO_ERR_CUDA if a peer failed to arrive within 20 s (synchronises).
 */
#define VIBO_COMM_HANDLE_BYTES 64
typedef struct vibo_comm vibo_comm;
int vibo_comm_create(int rank, int world_size, size_t max_floats, vibo_comm** out, void* handle_out);
int vibo_comm_connect(vibo_comm* comm, const void* all_handles);
int vibo_comm_allreduce(vibo_comm* comm, float* data, size_t n, void* stream);
int vibo_comm_allreduce_adam(vibo_comm* comm, float* data, size_t n, size_t skip, float* param,
                             float* exp_avg, float* exp_avg_sq, const int64_t* step, float lr,
                             float beta1, float beta2, float eps, void* stream);
int vibo_comm_status(vibo_comm* comm);
int vibo_comm_destroy(vibo_comm* comm);
const char* vibo_comm_last_error(void);

/*
 * Measurement hooks (used by bench.py; no effect on results).
 *   vibo_launch_count      kernels this library has launched in this process.
 *   vibo_profile_begin     start bracketing every launch of the fused kernel
 *                          with CUDA events on the launching stream.
 *   vibo_profile_end       stop; synchronises the recorded events and returns
 *                          the number of bracketed launches and their total
 *                          device time in milliseconds.
 *   vibo_single_pass       1 if vibo_fused_elbo would run the single-pass
 *                          kernel for `desc` (16-byte aligned rows assumed),
 *                          0 if it composes the general kernels.
 */
int vibo_single_pass(const vibo_desc* desc);
/* Largest num_item any entry point accepts for desc->ability_dim (2048 for D <= 4, 1024 above):
 * item slabs live in registers of at most 16 warps.  0 if ability_dim is unsupported.  Callers
 * (module constructor, CLI) check this up front instead of failing at the first step. */
int vibo_max_items(const vibo_desc* desc);
uint64_t vibo_launch_count(void);
int vibo_profile_begin(void);
int vibo_profile_end(int* n_launches, double* total_ms);

#ifdef __cplusplus
}
#endif
#endif /* VIBO_B200_H_ */
