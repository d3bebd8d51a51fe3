"""Literal CPU port of the reference's amortized-ELBO path (torch, fp32 ops).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.
It exists for three jobs:

1. parity oracle at module level: same parameters (reference ``state_dict``
   keys), same inputs, same injected noise -> loss and every parameter
   gradient, computed with the *same op sequence* the reference issues
   (per-cell MLP over ``(P*I, 1[+F])`` rows, product-of-experts, ``torch.mm``
   link, ``torch.distributions.Bernoulli.log_prob``), so fp32 rounding
   behaviour is the reference's;
2. the ``cpu_baseline`` / ``--impl reference`` arm of ``bench.py`` on the GPU
   box, where ``/root/reference`` does not exist: because the op sequence is
   the reference's, its timing is a fair stand-in ("kind": "port");
3. the checker in ``__graft_entry__.smoke()``.

Pinned against the live reference by ``tests/golden/make_golden.py`` ->
``tests/golden/*.npz`` -> ``tests/test_oracle_golden.py``.

The arithmetic lives in PyTorch (unpinned third-party dependency of the
reference, README.md:18-22); this port calls the same torch entry points the
reference calls.  Written functionally (no nn.Module mirror of the reference
classes); each function cites the reference lines it restates, relative to
the reference checkout.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


def item_feat_width(irt_model: int, ability_dim: int) -> int:
    """src/torch_core/models.py:331-332, 523-524, 538-539."""
    return {1: 1, 2: ability_dim + 1, 3: ability_dim + 2}[irt_model]


def init_params(irt_model: int, ability_dim: int, num_item: int, hidden_dim: int = 64,
                conditional: bool = False, n_flows: int = 0,
                generator: Optional[torch.Generator] = None) -> Params:
    """Fresh parameters with the reference's *distributions* (not its RNG
    stream): Linear weights xavier-normal with gain sqrt(2), biases 0
    (models.py:512-518); embeddings N(0, 1) (models.py:718-719); planar-flow
    u, w ~ N(0, 1), b = 1 (flows.py:17-19)."""
    D, H = ability_dim, hidden_dim
    Fw = item_feat_width(irt_model, D)
    in_dim = 1 + (Fw if conditional else 0)

    def xavier(out_f, in_f):
        std = math.sqrt(2.0) * math.sqrt(2.0 / (in_f + out_f))
        return torch.randn(out_f, in_f, generator=generator) * std

    p: Params = {
        "ability_encoder.mlp.0.weight": xavier(H, in_dim),
        "ability_encoder.mlp.0.bias": torch.zeros(H),
        "ability_encoder.mlp.2.weight": xavier(H, H),
        "ability_encoder.mlp.2.bias": torch.zeros(H),
        "ability_encoder.mlp.4.weight": xavier(2 * D, H),
        "ability_encoder.mlp.4.bias": torch.zeros(2 * D),
        "item_encoder.mu_lookup.weight": torch.randn(num_item, Fw, generator=generator),
        "item_encoder.logvar_lookup.weight": torch.randn(num_item, Fw, generator=generator),
    }
    for k in range(n_flows):
        for name, d in (("ability_norm_flows", D), ("item_norm_flows", Fw)):
            p[f"{name}.flows.{k}.u"] = torch.randn(d, generator=generator)
            p[f"{name}.flows.{k}.w"] = torch.randn(d, generator=generator)
            p[f"{name}.flows.{k}.b"] = torch.ones(1)
    return p


def encoder_mlp(p: Params, x: Tensor) -> Tensor:
    """Linear -> ELU -> Linear -> ELU -> Linear, models.py:575-582."""
    h = F.elu(F.linear(x, p["ability_encoder.mlp.0.weight"], p["ability_encoder.mlp.0.bias"]))
    h = F.elu(F.linear(h, p["ability_encoder.mlp.2.weight"], p["ability_encoder.mlp.2.bias"]))
    return F.linear(h, p["ability_encoder.mlp.4.weight"], p["ability_encoder.mlp.4.bias"])


def product_of_experts(mu: Tensor, logvar: Tensor, eps: float = 1e-8):
    """src/utils.py:105-113; first dim is the expert dim."""
    var = torch.exp(logvar) + eps
    T = 1.0 / var
    pd_mu = torch.sum(mu * T, dim=0) / torch.sum(T, dim=0)
    pd_var = 1.0 / torch.sum(T, dim=0)
    return pd_mu, torch.log(pd_var)


def ability_posterior_mean(p: Params, response: Tensor, mask: Tensor, item_feat: Optional[Tensor],
                           ability_dim: int):
    """Per-cell encoder + mean merge: models.py:584-594 (``mlp1`` / ``mlp2``), :631-650
    (``_forward_mean`` incl. the per-person loop over observed cells when anything is missing)."""
    P, I, _ = response.shape
    flat = response.reshape(P * I, 1)
    if item_feat is not None:
        tiled = item_feat.unsqueeze(0).repeat(P, 1, 1).reshape(P * I, item_feat.shape[1])
        flat = torch.cat([flat, tiled], dim=1)
    h = F.elu(F.linear(flat, p["ability_encoder.mlp1.0.weight"], p["ability_encoder.mlp1.0.bias"]))
    h = F.linear(h, p["ability_encoder.mlp1.2.weight"], p["ability_encoder.mlp1.2.bias"])
    hid = F.elu(h).reshape(P, I, -1)
    H = hid.shape[2]
    has_missing = bool(torch.sum(1 - mask).item())
    if has_missing:
        rows = []
        for i in range(P):
            keep = mask[i].repeat(1, H).bool()
            n_i = int(mask[i].squeeze().sum().item())
            rows.append(hid[i][keep].view(n_i, H).mean(0))
        hid_mean = torch.stack(rows)
    else:
        hid_mean = hid.mean(1)
    o = F.elu(F.linear(hid_mean, p["ability_encoder.mlp2.0.weight"], p["ability_encoder.mlp2.0.bias"]))
    o = F.linear(o, p["ability_encoder.mlp2.2.weight"], p["ability_encoder.mlp2.2.bias"])
    mu, logvar = torch.chunk(o, 2, dim=1)
    return mu, logvar


def ability_posterior(p: Params, response: Tensor, mask: Tensor, item_feat: Optional[Tensor],
                      ability_dim: int, replace_missing_with_prior: bool = True):
    """Per-cell encoder + product merge.

    models.py:652-661 (unconditional input ``(P*I, 1)``), models.py:695-710
    (conditional input ``cat[r_ij, item_feat_j]``), models.py:596-629
    (``_forward_product`` incl. the per-person loop when anything is missing).
    """
    P, I, _ = response.shape
    D = ability_dim
    flat = response.reshape(P * I, 1)
    if item_feat is not None:
        tiled = item_feat.unsqueeze(0).repeat(P, 1, 1).reshape(P * I, item_feat.shape[1])
        flat = torch.cat([flat, tiled], dim=1)
    mu_flat, lv_flat = torch.chunk(encoder_mlp(p, flat), 2, dim=1)
    mu_set = mu_flat.reshape(P, I, D)
    lv_set = lv_flat.reshape(P, I, D)
    has_missing = bool(torch.sum(1 - mask).item())
    if not has_missing:
        return product_of_experts(mu_set.permute(1, 0, 2), lv_set.permute(1, 0, 2))
    mus, lvs = [], []
    for i in range(P):
        if mask[i].sum().item() != I:
            keep = mask[i].bool().repeat(1, D)
            mu_i = mu_set[i][keep].view(-1, D)
            lv_i = lv_set[i][keep].view(-1, D)
            if replace_missing_with_prior:
                n_missing = I - mu_i.shape[0]
                zeros = torch.zeros(n_missing, D, dtype=mu_i.dtype)
                mu_i = torch.cat([mu_i, zeros], dim=0)
                lv_i = torch.cat([lv_i, zeros], dim=0)
        else:
            mu_i, lv_i = mu_set[i], lv_set[i]
        m, lv = product_of_experts(mu_i, lv_i)
        mus.append(m)
        lvs.append(lv)
    return torch.stack(mus), torch.stack(lvs)


def reparameterize(mean: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """models.py:506-510 with the noise passed in instead of randn_like."""
    return eps.mul(torch.exp(0.5 * logvar)).add(mean)


def irt_response_mu(ability: Tensor, item_feat: Tensor, irt_model: int) -> Tensor:
    """models.py:729-766 -> (P, I, 1)."""
    D = ability.shape[1]
    if irt_model == 1:
        logit = (torch.sum(ability, dim=1, keepdim=True) + item_feat.T).unsqueeze(2)
        return torch.sigmoid(logit)
    disc = item_feat[:, :D]
    diff = item_feat[:, D:D + 1]
    logit = (torch.mm(ability, -disc.T) + diff.T).unsqueeze(2)
    if irt_model == 2:
        return torch.sigmoid(logit)
    guess = torch.sigmoid(item_feat[:, D + 1:D + 2]).unsqueeze(0)
    return guess + (1.0 - guess) * torch.sigmoid(logit)


def masked_bernoulli_log_pdf(x: Tensor, mask: Tensor, probs: Tensor) -> Tensor:
    """src/utils.py:46-49.  ``validate_args=False`` is what
    ``Distribution.set_default_validate_args(False)`` gives the reference on
    torch >= 1.8 (needed for the -1 of missing cells, SURVEY.md finding 5)."""
    dist = torch.distributions.bernoulli.Bernoulli(probs=probs, validate_args=False)
    return dist.log_prob(x) * mask.float()


def normal_log_pdf(x: Tensor, mu: Tensor, logvar: Tensor) -> Tensor:
    """src/utils.py:59-61."""
    return torch.distributions.normal.Normal(mu, torch.exp(0.5 * logvar),
                                             validate_args=False).log_prob(x)


def standard_normal_log_pdf(x: Tensor) -> Tensor:
    """src/utils.py:64-67."""
    return torch.distributions.normal.Normal(torch.zeros_like(x), torch.ones_like(x),
                                             validate_args=False).log_prob(x)


def kl_standard_normal(mu: Tensor, logvar: Tensor) -> Tensor:
    """src/utils.py:85-88."""
    return torch.sum(-0.5 * (1 + logvar - mu.pow(2) - logvar.exp()), dim=1)


def planar_flows(p: Params, prefix: str, n_flows: int, z: Tensor):
    """src/torch_core/flows.py:21-41 applied n_flows times (flows.py:58-66)."""
    total = 0
    for k in range(n_flows):
        u, w, b = (p[f"{prefix}.flows.{k}.{n}"] for n in ("u", "w", "b"))
        uw = torch.dot(u, w)
        uhat = u + ((-1 + F.softplus(uw)) - uw) * w / torch.sum(w ** 2)
        zwb = torch.mv(z, w) + b
        h = torch.tanh(zwb)
        z = z + uhat.view(1, -1) * h.view(-1, 1)
        psi_u = torch.mv((1 - h ** 2).view(-1, 1) * w.view(1, -1), uhat)
        total = total + torch.log(torch.abs(1 + psi_u) + 1e-8)
    return z, total


def forward(p: Params, response: Tensor, mask: Tensor, eps_item: Tensor, eps_ability: Tensor, *,
            irt_model: int, ability_dim: int, conditional: bool = False, n_flows: int = 0,
            replace_missing_with_prior: bool = True) -> Dict[str, Tensor]:
    """models.py:337-371: encode (items first, then abilities), optional
    flows, decode.  ``response`` (P, I, 1) float, ``mask`` (P, I, 1) long."""
    item_mu = p["item_encoder.mu_lookup.weight"]
    item_lv = p["item_encoder.logvar_lookup.weight"]
    item_feat = reparameterize(item_mu, item_lv, eps_item)
    if "ability_encoder.mlp1.0.weight" in p:   # --ability-merge mean
        a_mu, a_lv = ability_posterior_mean(p, response, mask, item_feat if conditional else None, ability_dim)
    else:
        a_mu, a_lv = ability_posterior(p, response, mask, item_feat if conditional else None,
                                       ability_dim, replace_missing_with_prior)
    ability = reparameterize(a_mu, a_lv, eps_ability)
    out = dict(ability=ability, ability_mu=a_mu, ability_logvar=a_lv,
               item_feat=item_feat, item_feat_mu=item_mu, item_feat_logvar=item_lv)
    if n_flows > 0:
        ability_k, a_ldj = planar_flows(p, "ability_norm_flows", n_flows, ability)
        item_k, i_ldj = planar_flows(p, "item_norm_flows", n_flows, item_feat)
        out.update(ability_k=ability_k, ability_logabsdetjac=a_ldj,
                   item_feat_k=item_k, item_feat_logabsdetjac=i_ldj)
        out["response_mu"] = irt_response_mu(ability_k, item_k, irt_model)
    else:
        out["response_mu"] = irt_response_mu(ability, item_feat, irt_model)
    return out


def negative_elbo(fw: Dict[str, Tensor], response: Tensor, mask: Tensor, *, n_flows: int = 0,
                  annealing_factor: float = 1.0, use_kl_divergence: bool = True) -> Tensor:
    """models.py:380-443."""
    ll = masked_bernoulli_log_pdf(response, mask, fw["response_mu"]).sum()
    if n_flows > 0:
        log_q_u0 = normal_log_pdf(fw["ability"], fw["ability_mu"], fw["ability_logvar"]).sum()
        log_q_d0 = normal_log_pdf(fw["item_feat"], fw["item_feat_mu"], fw["item_feat_logvar"]).sum()
        log_p_uk = standard_normal_log_pdf(fw["ability_k"]).sum()
        log_p_dk = standard_normal_log_pdf(fw["item_feat_k"]).sum()
        log_q_uk = log_q_u0 - fw["ability_logabsdetjac"].sum()
        log_q_dk = log_q_d0 - fw["item_feat_logabsdetjac"].sum()
        elbo = (ll + log_p_uk + log_p_dk) - (log_q_uk + log_q_dk)
    elif use_kl_divergence:
        kl_u = kl_standard_normal(fw["ability_mu"], fw["ability_logvar"]).sum()
        kl_d = kl_standard_normal(fw["item_feat_mu"], fw["item_feat_logvar"]).sum()
        elbo = ll - annealing_factor * kl_u - annealing_factor * kl_d
    else:
        log_p_u = standard_normal_log_pdf(fw["ability"]).sum()
        log_p_d = standard_normal_log_pdf(fw["item_feat"]).sum()
        log_q_u = normal_log_pdf(fw["ability"], fw["ability_mu"], fw["ability_logvar"]).sum()
        log_q_d = normal_log_pdf(fw["item_feat"], fw["item_feat_mu"], fw["item_feat_logvar"]).sum()
        elbo = (ll + log_p_u + log_p_d) - (log_q_u + log_q_d)
    return -elbo


def loss_and_grads(params: Params, response: Tensor, mask: Tensor, eps_item: Tensor,
                   eps_ability: Tensor, *, irt_model: int, ability_dim: int,
                   conditional: bool = False, n_flows: int = 0,
                   replace_missing_with_prior: bool = True, annealing_factor: float = 1.0,
                   use_kl_divergence: bool = True, want_grads: bool = True):
    """One reference training-step's worth of math without the optimiser:
    vibo.py:243-267 (forward, elbo, backward)."""
    leaves = {k: v.detach().clone().requires_grad_(want_grads) for k, v in params.items()}
    with torch.set_grad_enabled(want_grads):
        fw = forward(leaves, response, mask, eps_item, eps_ability, irt_model=irt_model,
                     ability_dim=ability_dim, conditional=conditional, n_flows=n_flows,
                     replace_missing_with_prior=replace_missing_with_prior)
        loss = negative_elbo(fw, response, mask, n_flows=n_flows,
                             annealing_factor=annealing_factor,
                             use_kl_divergence=use_kl_divergence)
    grads = {}
    if want_grads:
        loss.backward()
        grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v))
                 for k, v in leaves.items()}
    return loss.detach(), {k: v.detach() for k, v in fw.items()}, grads


def log_marginal(params: Params, response: Tensor, mask: Tensor, eps_items, eps_abilities, *,
                 irt_model: int, ability_dim: int, conditional: bool = False, n_flows: int = 0,
                 replace_missing_with_prior: bool = True) -> Tensor:
    """models.py:445-504: logsumexp over S batch-summed log-weights - log S."""
    logw = []
    for e_i, e_a in zip(eps_items, eps_abilities):
        loss, _, _ = loss_and_grads(params, response, mask, e_i, e_a, irt_model=irt_model,
                                    ability_dim=ability_dim, conditional=conditional,
                                    n_flows=n_flows,
                                    replace_missing_with_prior=replace_missing_with_prior,
                                    use_kl_divergence=False, want_grads=False)
        logw.append(-loss)
    logw = torch.stack(logw)
    return torch.logsumexp(logw, 0) - math.log(len(logw))


def adam_train_step(params: Params, opt_state: dict, response, mask, eps_item, eps_ability, *,
                    lr: float = 5e-3, **kw):
    """vibo.py:243-268 including ``optimizer.step()`` (Adam defaults,
    vibo.py:221).  ``opt_state`` is created on first use."""
    if "opt" not in opt_state:
        opt_state["leaves"] = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        opt_state["opt"] = torch.optim.Adam(list(opt_state["leaves"].values()), lr=lr)
    leaves, opt = opt_state["leaves"], opt_state["opt"]
    opt.zero_grad()
    fw = forward(leaves, response, mask, eps_item, eps_ability,
                 irt_model=kw["irt_model"], ability_dim=kw["ability_dim"],
                 conditional=kw.get("conditional", False), n_flows=kw.get("n_flows", 0),
                 replace_missing_with_prior=kw.get("replace_missing_with_prior", True))
    loss = negative_elbo(fw, response, mask, n_flows=kw.get("n_flows", 0),
                         annealing_factor=kw.get("annealing_factor", 1.0),
                         use_kl_divergence=kw.get("use_kl_divergence", True))
    loss.backward()
    opt.step()
    return loss.detach()
