"""Kernel-boundary oracle for the VIBO ELBO hot path (numpy, closed form).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the
product package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and
there only as the checker.

What this file is
-----------------
A CPU restatement, in plain numpy, of the arithmetic the reference performs
between "encoder MLP outputs" and "scalar ELBO", written in the collapsed
closed form the CUDA kernels implement, together with
the hand-derived gradients the kernels emit.  Every function cites the
reference lines (relative to the reference checkout) it restates.

Parity pinning
--------------
The reference has no tests or golden vectors for this path, so
the oracle is pinned differentially: ``tests/golden/make_golden.py`` imports
the live reference (``src.torch_core.models``) in the build container, runs it
on seeded inputs with injected noise, and commits inputs + outputs under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this file and
``oracle/reference_port.py`` against those fixtures.

Conventions
-----------
P persons, I items, D ability dims, F item-feature width (1 | D+1 | D+2).
``response`` (P, I) holds 0/1 (anything where mask == 0, -1 by convention,
src/config.py:14).  ``mask`` (P, I) holds 0/1.  ``table`` (2, It, 2D) holds
the raw encoder outputs for response value r in {0, 1}: ``[..., :D]`` is the
expert mean, ``[..., D:]`` the expert log-variance; It == 1 for the
unconditional encoder (same expert for every item), It == I for
``--conditional-posterior``.
"""
from __future__ import annotations

import numpy as np

# torch.finfo(torch.float32).eps: the clamp torch.distributions applies to
# Bernoulli probabilities (torch/distributions/utils.py clamp_probs), which
# is what src/utils.py:46-49 goes through.
EPS32 = float(np.finfo(np.float32).eps)
# |logit| beyond which a 1PL/2PL cell is clamped: log((1 - eps) / eps).
LOGIT_CLAMP = float(np.log((1.0 - EPS32) / EPS32))
POE_EPS = 1e-8  # src/utils.py:105 default eps of product_of_experts
LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)

MISSING_PRIOR = 0  # replace_missing_with_prior=True, src/torch_core/models.py:614-618
MISSING_DROP = 1   # --drop-missing


def item_feat_width(irt_model: int, ability_dim: int) -> int:
    """F for 1PL/2PL/3PL, src/torch_core/models.py:331-332, 523-524, 538-539."""
    return {1: 1, 2: ability_dim + 1, 3: ability_dim + 2}[irt_model]


def _sigmoid(z):
    out = np.empty_like(z)
    pos = z >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-z[pos]))
    ez = np.exp(z[~pos])
    out[~pos] = ez / (1.0 + ez)
    return out


def expert_precision(table, ability_dim):
    """Split raw table into (mu, tau) with tau = 1 / (exp(logvar) + 1e-8).

    src/utils.py:107-108 (``var = exp(logvar) + eps; T = 1 / var``).
    """
    D = ability_dim
    mu = table[..., :D]
    lam = table[..., D:]
    tau = 1.0 / (np.exp(lam) + POE_EPS)
    return mu, lam, tau


def encode(response, mask, table, ability_dim, missing_policy=MISSING_PRIOR):
    """Product-of-experts ability posterior for every person.

    Restates src/torch_core/models.py:596-629 (``_forward_product``) and
    src/utils.py:105-113 (``product_of_experts``) with the per-cell MLP
    replaced by the 2-row (or 2*I-row) expert table: an observed cell
    contributes the expert of its response value; a missing cell contributes
    a N(0, 1) expert (mu 0, precision 1/(1+1e-8)) when
    ``replace_missing_with_prior`` or nothing under ``--drop-missing``.

    Returns dict with S, N (P, D), ability_mu, ability_logvar (P, D).
    """
    dt = table.dtype
    P, I = response.shape
    D = ability_dim
    mu, _, tau = expert_precision(table, D)          # (2, It, D)
    It = table.shape[1]
    obs = mask != 0
    one = obs & (response > 0.5)
    zero = obs & ~(response > 0.5)
    if It == 1:
        n1 = one.sum(1).astype(dt)[:, None]
        n0 = zero.sum(1).astype(dt)[:, None]
        S = n0 * tau[0, 0][None, :] + n1 * tau[1, 0][None, :]
        N = n0 * (mu[0, 0] * tau[0, 0])[None, :] + n1 * (mu[1, 0] * tau[1, 0])[None, :]
    else:
        w1 = one.astype(dt)
        w0 = zero.astype(dt)
        S = w0 @ tau[0] + w1 @ tau[1]
        N = w0 @ (mu[0] * tau[0]) + w1 @ (mu[1] * tau[1])
    if missing_policy == MISSING_PRIOR:
        nmiss = (~obs).sum(1).astype(dt)[:, None]
        S = S + nmiss * dt.type(1.0 / (1.0 + POE_EPS))
    with np.errstate(divide="ignore", invalid="ignore"):
        ability_mu = N / S
        ability_logvar = np.log(1.0 / S)
    return dict(S=S, N=N, ability_mu=ability_mu, ability_logvar=ability_logvar)


def link_logit(ability, item_feat, irt_model):
    """IRT logits z (P, I) and guess g (I,) or None.

    src/torch_core/models.py:729-766: 1PL ``sum_d theta_d + b_j``;
    2PL/3PL ``-theta . a_j + b_j``; 3PL guess ``sigmoid(item_feat[:, D+1])``.
    """
    D = ability.shape[1]
    if irt_model == 1:
        z = ability.sum(1, keepdims=True) + item_feat[:, 0][None, :]
        return z, None
    a = item_feat[:, :D]
    b = item_feat[:, D]
    z = ability @ (-a.T) + b[None, :]
    if irt_model == 2:
        return z, None
    g = _sigmoid(item_feat[:, D + 1])
    return z, g


def decode(ability, item_feat, irt_model):
    """response_mu (P, I): src/torch_core/models.py:373-378, 529-548."""
    z, g = link_logit(ability, item_feat, irt_model)
    s = _sigmoid(z)
    if g is None:
        return s
    return g[None, :] + (1.0 - g[None, :]) * s


def bernoulli_loglik(response, mask, prob):
    """Masked Bernoulli log-likelihood per cell and d ll / d prob.

    src/utils.py:46-49 -> torch Bernoulli(probs).log_prob: probabilities are
    clamped to [eps32, 1 - eps32] (torch/distributions/utils.py clamp_probs),
    turned into logits and fed to -BCEWithLogits.  Mathematically that is
    ``x log p~ + (1 - x) log(1 - p~)``; the clamp passes gradient only where
    eps32 <= p <= 1 - eps32.
    """
    x = (response > 0.5).astype(prob.dtype)
    pc = np.clip(prob, EPS32, 1.0 - EPS32)
    ll = x * np.log(pc) + (1.0 - x) * np.log1p(-pc)
    inside = (prob >= EPS32) & (prob <= 1.0 - EPS32)
    dll_dp = np.where(inside, x / pc - (1.0 - x) / (1.0 - pc), 0.0)
    o = (mask != 0).astype(prob.dtype)
    return ll * o, dll_dp * o


def link_loglik(response, mask, ability, item_feat, irt_model, want_grads=True):
    """LL = sum_ij o_ij ll_ij and its gradients w.r.t. ability and item_feat.

    Restates decode (models.py:729-766) + masked_bernoulli_log_pdf(...).sum()
    (models.py:399, utils.py:46-49).  Gradients are of **LL** (not of the
    loss): ``g_ability = dLL/d theta`` (P, D), ``g_item = dLL/d item_feat``
    (I, F), SURVEY.md Appendix A "Backward".
    """
    D = ability.shape[1]
    z, g = link_logit(ability, item_feat, irt_model)
    s = _sigmoid(z)
    p = s if g is None else g[None, :] + (1.0 - g[None, :]) * s
    ll, dll_dp = bernoulli_loglik(response, mask, p)
    out = dict(ll=ll.sum(), ll_person=ll.sum(1))
    if not want_grads:
        return out
    dp_dz = s * (1.0 - s) if g is None else (1.0 - g[None, :]) * s * (1.0 - s)
    dz = dll_dp * dp_dz                                  # dLL/dz (P, I)
    g_item = np.zeros_like(item_feat)
    if irt_model == 1:
        g_ability = np.repeat(dz.sum(1, keepdims=True), D, axis=1)
        g_item[:, 0] = dz.sum(0)
    else:
        a = item_feat[:, :D]
        g_ability = -(dz @ a)
        g_item[:, :D] = -(dz.T @ ability)
        g_item[:, D] = dz.sum(0)
        if irt_model == 3:
            dgam = (dll_dp * (1.0 - s)).sum(0) * g * (1.0 - g)
            g_item[:, D + 1] = dgam
    out.update(g_ability=g_ability, g_item=g_item, dz=dz)
    return out


def kl_standard_normal(mu, logvar):
    """src/utils.py:85-88, summed over everything."""
    return float((-0.5 * (1.0 + logvar - mu ** 2 - np.exp(logvar))).sum())


def encode_backward(response, mask, table, ability_dim, S, ability_mu, g_mu, g_logvar):
    """d loss / d table given d loss / d (ability_mu, ability_logvar).

    Chain rule through src/utils.py:105-113: with S = sum tau, N = sum mu tau,
    mu_bar = N / S, logvar = -log S:  GN = g_mu / S,
    GS = -(g_mu * mu_bar + g_logvar) / S;  A^r_j = sum_{i: o=1, x=r} GN_i,
    B^r_j likewise with GS;  d/d mu^r_j = tau A,
    d/d lam^r_j = (mu A + B) * (-exp(lam) tau^2).   Missing cells have no
    parameters (prior experts / dropped).
    """
    dt = table.dtype
    D = ability_dim
    mu, lam, tau = expert_precision(table, D)
    It = table.shape[1]
    GN = g_mu / S
    GS = -(g_mu * ability_mu + g_logvar) / S
    obs = mask != 0
    w1 = (obs & (response > 0.5)).astype(dt)
    w0 = (obs & ~(response > 0.5)).astype(dt)
    A = np.stack([w0.T @ GN, w1.T @ GN])                 # (2, I, D)
    B = np.stack([w0.T @ GS, w1.T @ GS])
    if It == 1:
        A = A.sum(1, keepdims=True)
        B = B.sum(1, keepdims=True)
    g_table = np.empty_like(table)
    g_table[..., :D] = tau * A
    g_table[..., D:] = (mu * A + B) * (-np.exp(lam) * tau ** 2)
    return g_table


def planar_params(u, w, g_uhat=None, g_w_out=None):
    """Spec of vibo_planar_params_forward / _backward (reference flows.py:26-29): u, w (K, D) ->
    uhat = u + c w, c = (softplus(a) - 1 - a) / n, a = w.u, n = |w|^2 (torch softplus: threshold 20);
    with (g_uhat, g_w_out) also the gradients of u and w (b passes through unchanged)."""
    a = (u * w).sum(1, keepdims=True)
    n = (w * w).sum(1, keepdims=True)
    sp = np.where(a > 20.0, a, np.log1p(np.exp(np.minimum(a, 20.0))))
    c = (sp - 1.0 - a) / n
    uhat = u + c * w
    if g_uhat is None:
        return uhat
    sig = np.where(a > 20.0, 1.0, 1.0 / (1.0 + np.exp(-a)))
    gw = (g_uhat * w).sum(1, keepdims=True)
    dc_da, dc_dn = (sig - 1.0) / n, -c / n
    g_u = g_uhat + gw * dc_da * w
    g_w = g_w_out + c * g_uhat + gw * (dc_da * u + 2.0 * dc_dn * w)
    return uhat, g_u, g_w


def encode_backward_counts(counts, table, ability_dim, S, ability_mu, g_mu, g_logvar):
    """encode_backward for the UNCONDITIONAL table (one expert per response value, It = 1) from the
    per-person counts (observed ones, observed cells) of person_counts: the row enters A^r, B^r only as
    n^r_i = #{j: o_ij = 1, x_ij = r}, so A^r = sum_i n^r_i GN_i.  Spec of vibo_encode_backward_counts;
    equal to encode_backward on the same inputs (tests/test_oracle_golden.py)."""
    D = ability_dim
    assert table.shape[1] == 1
    mu, lam, tau = expert_precision(table, D)
    GN = g_mu / S
    GS = -(g_mu * ability_mu + g_logvar) / S
    n1 = counts[:, 0:1].astype(table.dtype)
    n0 = counts[:, 1:2].astype(table.dtype) - n1
    A = np.stack([(n0 * GN).sum(0), (n1 * GN).sum(0)])[:, None, :]   # (2, 1, D)
    B = np.stack([(n0 * GS).sum(0), (n1 * GS).sum(0)])[:, None, :]
    g_table = np.empty_like(table)
    g_table[..., :D] = tau * A
    g_table[..., D:] = (mu * A + B) * (-np.exp(lam) * tau ** 2)
    return g_table


ELBO_KL = 0      # use_kl_divergence=True, models.py:427-430
ELBO_SAMPLE = 1  # use_kl_divergence=False, models.py:432-441


def fused_elbo(response, mask, table, item_feat, eps_ability, *, irt_model,
               beta=1.0, missing_policy=MISSING_PRIOR, elbo_form=ELBO_KL,
               want_grads=True):
    """Everything the fused CUDA kernel computes, in one call.

    Forward (SURVEY.md Appendix A steps 3-7 = models.py:364-368, 373,
    399, 428): PoE posterior, reparameterised draw
    ``theta = mu_bar + exp(logvar / 2) * eps`` (models.py:506-510), link,
    masked Bernoulli LL, and the per-person prior term:

    * ELBO_KL:     ``person_term = KL(q(theta_i) || N(0, 1))`` (utils.py:85-88)
    * ELBO_SAMPLE: ``person_term = log p(theta_i) - log q(theta_i)``
      (models.py:433-435, utils.py:59-67)

    The part of the loss the kernel owns is
    ``loss_k = -LL + beta * KL``  (ELBO_KL)   or
    ``loss_k = -LL - person_term`` (ELBO_SAMPLE; beta is not applied to the
    sample form in the reference, models.py:438-441).
    Item-side terms (KL_d or log p(d) - log q(d)) are outside the kernel.

    Gradients returned are of ``loss_k``: ``g_item`` (I, F) through the link
    only, ``g_table`` (2, It, 2D) through the encoder experts.
    """
    dt = table.dtype
    D = eps_ability.shape[1]
    enc = encode(response, mask, table, D, missing_policy)
    amu, alv, S = enc["ability_mu"], enc["ability_logvar"], enc["S"]
    std = np.exp(0.5 * alv)
    theta = amu + std * eps_ability
    lk = link_loglik(response, mask, theta, item_feat, irt_model, want_grads)
    out = dict(ll=float(lk["ll"]), ability_mu=amu, ability_logvar=alv, ability=theta)
    if elbo_form == ELBO_KL:
        out["person_term"] = kl_standard_normal(amu, alv)
        out["loss_k"] = -out["ll"] + beta * out["person_term"]
    else:
        log_p = (-0.5 * theta ** 2 - LOG_SQRT_2PI).sum()
        log_q = (-(theta - amu) ** 2 / (2.0 * std ** 2) - np.log(std) - LOG_SQRT_2PI).sum()
        out["person_term"] = float(log_p - log_q)
        out["loss_k"] = -out["ll"] - out["person_term"]
    if not want_grads:
        return out
    g_theta = -lk["g_ability"]                            # d loss_k / d theta via link
    if elbo_form == ELBO_KL:
        g_mu = g_theta + beta * amu
        g_lv = 0.5 * g_theta * eps_ability * std + 0.5 * beta * (np.exp(alv) - 1.0)
    else:
        g_theta = g_theta + theta                         # -d log p(theta) / d theta
        g_mu = g_theta
        # log q(theta) = -eps^2/2 - logvar/2 - c along the reparameterised path
        g_lv = 0.5 * g_theta * eps_ability * std - 0.5
    out["g_item"] = (-lk["g_item"]).astype(dt)
    out["g_table"] = encode_backward(response, mask, table, D, S, amu, g_mu, g_lv)
    out["g_ability_mu"] = g_mu
    out["g_ability_logvar"] = g_lv
    return out


def person_counts(response, mask):
    """(n0, n1, nmiss) per person: what the unconditional encoder reduces to
    (SURVEY.md finding 2)."""
    obs = mask != 0
    one = obs & (response > 0.5)
    n1 = one.sum(1)
    nobs = obs.sum(1)
    return nobs - n1, n1, response.shape[1] - nobs


# ---------------------------------------------------------------------------
# Planar flows on the abilities, fused per person (spec of vibo_flow_person_forward /
# vibo_flow_person_backward; reference src/torch_core/flows.py:21-41, :58-66 and the flow form
# of the ELBO, src/torch_core/models.py:406-424).  uhat is the invertibility-corrected u
# (flows.py:26-29), formed by the caller.
# ---------------------------------------------------------------------------
def flow_person(ability_mu, ability_logvar, eps, uhat, w, b, g_ability_k=None, g_term=0.0):
    """Forward: theta_0 = mu + eps exp(lv / 2); K planar steps; term = sum_i [log N(theta_K; 0, 1)
    - log N(theta_0; mu, exp lv) + sum_k ldj_k].  With ``g_ability_k`` (d loss / d theta_K) and
    ``g_term`` (d loss / d term) also the closed-form backward."""
    mu, lv, eps = (np.asarray(a, dtype=np.float64) for a in (ability_mu, ability_logvar, eps))
    uhat, w, b = (np.asarray(a, dtype=np.float64) for a in (uhat, w, b))
    K = uhat.shape[0]
    sd = np.exp(0.5 * lv)
    z = mu + eps * sd
    theta0 = z.copy()
    zin, hs = [], []
    wu = (w * uhat).sum(1)
    ldj = np.zeros(mu.shape[0])
    for k in range(K):
        zin.append(z.copy())
        h = np.tanh(z @ w[k] + b[k])
        hs.append(h)
        z = z + h[:, None] * uhat[k][None, :]
        ldj += np.log(np.abs(1.0 + (1.0 - h * h) * wu[k]) + 1e-8)
    term = float((-0.5 * z ** 2 + 0.5 * eps ** 2 + 0.5 * lv).sum() + ldj.sum())
    out = dict(ability_0=theta0, ability_k=z, term=term)
    if g_ability_k is None:
        return out
    c = float(g_term)
    G = np.asarray(g_ability_k, dtype=np.float64) - c * z
    g_uhat, g_w, g_b = np.zeros_like(uhat), np.zeros_like(w), np.zeros_like(b)
    for k in range(K - 1, -1, -1):
        h = hs[k]
        omh = 1.0 - h * h
        q = 1.0 + omh * wu[k]
        dl = c * np.where(q >= 0, 1.0, -1.0) / (np.abs(q) + 1e-8)        # d loss / d s_k
        ga = (G @ uhat[k]) * omh + dl * (-2.0 * h * omh) * wu[k]          # d loss / d a_k
        g_uhat[k] = (G * h[:, None]).sum(0) + (dl * omh).sum() * w[k]
        g_w[k] = (ga[:, None] * zin[k]).sum(0) + (dl * omh).sum() * uhat[k]
        g_b[k] = ga.sum()
        G = G + ga[:, None] * w[k][None, :]
    out.update(g_mu=G, g_logvar=0.5 * G * eps * sd + 0.5 * c, g_uhat=g_uhat, g_w=g_w, g_b=g_b)
    return out


# ---------------------------------------------------------------------------
# --ability-merge mean (reference src/torch_core/models.py:584-594, :631-650) with the table
# collapse: the per-cell hidden vector takes 2 (unconditional) or 2 I (conditional) distinct
# values, so the per-person masked mean is  (indicator or counts) x hidden table / n_observed.
# ---------------------------------------------------------------------------
def mean_merge_hidden(response, mask, hidden_table):
    """hidden_table (2, It, H) -> (P, H) masked mean of the selected hidden vectors.
    Rows without any observed cell give NaN (0 / 0), like the reference's mean of an empty set."""
    x = np.asarray(response, dtype=np.float64)
    o = np.asarray(mask) != 0
    T = np.asarray(hidden_table, dtype=np.float64)
    one = (x > 0.5) & o
    zero = o & ~one
    if T.shape[1] == 1:
        n0, n1, _ = person_counts(response, mask)
        tot = n0[:, None] * T[0, 0][None, :] + n1[:, None] * T[1, 0][None, :]
    else:
        tot = one.astype(np.float64) @ T[1] + zero.astype(np.float64) @ T[0]
    with np.errstate(invalid="ignore", divide="ignore"):
        return tot / o.sum(1, keepdims=True)
