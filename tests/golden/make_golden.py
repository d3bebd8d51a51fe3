"""Generate golden fixtures by running the LIVE reference implementation.

Run in the build container only (``/root/reference`` does not exist on the
GPU box):

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

It imports ``src.torch_core.models`` from the reference checkout (read-only),
builds ``VIBO_{1,2,3}PL`` with seeded initial weights, injects pre-drawn noise
by shadowing ``reparameterize_gaussian`` on the instance (draw order: items
``(I, F)`` first, then abilities ``(P, D)``; models.py:361, 368), runs
``forward`` + ``elbo`` + ``backward`` on CPU fp32 and stores inputs, the
state dict, the outputs and every parameter gradient.  Nothing from the
reference's source is copied; only its numerical outputs are recorded.

Shims needed on torch 2.11 (SURVEY.md finding 5): argument validation of
torch.distributions is switched off so that the -1 of missing cells does not
raise inside ``Bernoulli.log_prob``.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("VIBO_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    torch.distributions.Distribution.set_default_validate_args(False)
    from src.torch_core import models  # noqa: WPS433
    return models


def synth_responses(P, I, D, irt_model, seed, missing_frac, continuous=False):
    """Plain-torch restatement of the generative draw of
    src/pyro_core/models.py:68-110 (abilities, then item features, then
    Bernoulli responses) + src/datasets.py:46-78 style masking."""
    g = torch.Generator().manual_seed(seed)
    Fw = {1: 1, 2: D + 1, 3: D + 2}[irt_model]
    ability = torch.randn(P, D, generator=g)
    item = torch.randn(I, Fw, generator=g)
    if irt_model == 1:
        z = ability.sum(1, keepdim=True) + item.T
    else:
        z = ability @ (-item[:, :D].T) + item[:, D:D + 1].T
    p = torch.sigmoid(z)
    if irt_model == 3:
        gs = torch.sigmoid(item[:, D + 1]).unsqueeze(0)
        p = gs + (1 - gs) * p
    resp = torch.bernoulli(p, generator=g)
    if continuous:   # --response-dist gaussian: real-valued responses around the link probability
        resp = p + 0.1 * torch.randn(P, I, generator=g)
    mask = torch.ones(P, I, dtype=torch.bool)
    if missing_frac > 0:
        mask = torch.rand(P, I, generator=g) >= missing_frac
        mask[0] = True                      # one fully observed row (models.py:621-623 branch)
        mask[1, : I // 2] = False           # one heavily masked row
        resp[~mask] = -1.0
    return resp.unsqueeze(2), mask.unsqueeze(2)


CASES = []


def case(name, irt, D, cond, P, I, missing=0.0, drop=False, beta=1.0, use_kl=True,
         flows=0, trained_like=False, seed=0, merge="product", generative="irt", response_dist="bernoulli"):
    c = dict(name=name, irt_model=irt, ability_dim=D, conditional=cond, P=P, I=I,
             missing_frac=missing, drop_missing=drop, beta=beta, use_kl=use_kl,
             n_flows=flows, trained_like=trained_like, seed=seed, merge=merge)
    if generative != "irt" or response_dist != "bernoulli":   # keys only where they differ from the defaults
        c.update(generative=generative, response_dist=response_dist)
    CASES.append(c)


# --- the grid (small enough that the whole CPU suite stays in seconds) -----
s = 100
for irt in (1, 2, 3):
    for D in (1, 3):
        for cond in (False, True):
            s += 1
            case(f"m{irt}pl_d{D}_{'cond' if cond else 'unc'}_full", irt, D, cond, 24, 20,
                 trained_like=(irt == 3 and cond and D >= 3), seed=s)
            s += 1
            case(f"m{irt}pl_d{D}_{'cond' if cond else 'unc'}_miss", irt, D, cond, 19, 17,
                 missing=0.15, beta=0.7, trained_like=(irt == 3 and cond and D >= 3), seed=s)
case("m2pl_d1_unc_drop", 2, 1, False, 21, 16, missing=0.2, drop=True, beta=0.3, seed=201)
case("m3pl_d2_cond_drop", 3, 2, True, 18, 15, missing=0.2, drop=True, seed=202)
case("m2pl_d1_unc_sampleform", 2, 1, False, 24, 20, use_kl=False, seed=203)
case("m3pl_d2_cond_sampleform_miss", 3, 2, True, 20, 13, missing=0.1, use_kl=False, seed=204)
case("m2pl_d1_unc_flows2_miss", 2, 1, False, 22, 19, missing=0.1, flows=2, use_kl=False, seed=205)
case("m3pl_d2_cond_flows2", 3, 2, True, 20, 16, flows=2, use_kl=False, seed=206)
case("m1pl_d2_unc_flows1", 1, 2, False, 17, 12, flows=1, use_kl=False, seed=207)
case("m3pl_d5_cond_trained", 3, 5, True, 32, 24, trained_like=True, seed=208)
case("m2pl_d1_unc_wide", 2, 1, False, 9, 140, seed=209)
case("m2pl_d1_unc_saturating", 2, 1, False, 16, 12, seed=210)
# --ability-merge mean (the constructor default of the reference; models.py:584-594, 631-650).
# The mean-merge posterior at fresh init is wide (sd ~ 1), which with N(0,1) item features puts
# logits in the 12..16 band where the fp32 reference is its own noise ("knife-edge"):
# the cases use the trained-like item state.
case("m2pl_d1_unc_mean_full", 2, 1, False, 24, 20, merge="mean", trained_like=True, seed=301)
case("m2pl_d2_unc_mean_miss", 2, 2, False, 19, 17, missing=0.15, beta=0.7, merge="mean", trained_like=True, seed=302)
case("m3pl_d2_cond_mean_full", 3, 2, True, 22, 18, merge="mean", trained_like=True, seed=303)
case("m1pl_d3_cond_mean_miss", 1, 3, True, 20, 15, missing=0.2, use_kl=False, merge="mean", trained_like=True,
     seed=304)
case("m2pl_d1_unc_mean_flows2", 2, 1, False, 18, 14, flows=2, use_kl=False, merge="mean", trained_like=True,
     seed=305)
# nonlinear generative models (models.py:769-919) and Gaussian responses (models.py:400-402,
# utils.py:52-56): SURVEY.md 8 rows f3 / f4
case("m2pl_d1_unc_link_full", 2, 1, False, 20, 16, generative="link", trained_like=True, seed=401)
case("m3pl_d2_cond_link_miss", 3, 2, True, 17, 13, missing=0.15, beta=0.7, generative="link", trained_like=True,
     seed=402)
case("m1pl_d1_unc_deep_full", 1, 1, False, 18, 15, generative="deep", trained_like=True, seed=403)
case("m2pl_d2_unc_deep_miss", 2, 2, False, 16, 14, missing=0.2, use_kl=False, generative="deep", trained_like=True,
     seed=404)
case("m2pl_d1_unc_residual_full", 2, 1, False, 19, 12, generative="residual", trained_like=True, seed=405)
case("m3pl_d2_unc_residual_miss", 3, 2, False, 15, 16, missing=0.1, generative="residual", trained_like=True,
     seed=406)
case("m2pl_d1_unc_gauss_full", 2, 1, False, 18, 14, response_dist="gaussian", trained_like=True, seed=407)
case("m2pl_d2_cond_gauss_miss", 2, 2, True, 16, 12, missing=0.15, beta=0.7, response_dist="gaussian",
     trained_like=True, seed=408)
case("m1pl_d1_unc_gauss_mean_miss", 1, 1, False, 17, 11, missing=0.1, merge="mean", response_dist="gaussian",
     trained_like=True, seed=409)
case("m2pl_d1_unc_gauss_drop_link", 2, 1, False, 15, 13, missing=0.2, drop=True, generative="link",
     response_dist="gaussian", trained_like=True, seed=410)


def run_case(models, c):
    irt, D, cond = c["irt_model"], c["ability_dim"], c["conditional"]
    P, I = c["P"], c["I"]
    cls = {1: models.VIBO_1PL, 2: models.VIBO_2PL, 3: models.VIBO_3PL}[irt]
    torch.manual_seed(c["seed"])
    model = cls(D, I, hidden_dim=64, ability_merge=c.get("merge", "product"), conditional_posterior=cond,
                generative_model=c.get("generative", "irt"), response_dist=c.get("response_dist", "bernoulli"),
                replace_missing_with_prior=not c["drop_missing"], n_norm_flows=c["n_flows"])
    init_state = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    if c["trained_like"]:
        with torch.no_grad():
            model.item_encoder.logvar_lookup.weight.fill_(float(np.log(0.01)))
            model.item_encoder.mu_lookup.weight.mul_(0.5)
    if c["name"].endswith("saturating"):
        with torch.no_grad():  # push logits past the eps32 clamp on purpose
            model.item_encoder.mu_lookup.weight[:, 1] = torch.linspace(-30, 30, I)
            model.item_encoder.logvar_lookup.weight.fill_(-8.0)
    Fw = model.item_feat_dim
    response, mask_b = synth_responses(P, I, D, irt, c["seed"] + 1000, c["missing_frac"],
                                       continuous=c.get("response_dist") == "gaussian")
    mask = mask_b.long()
    g = torch.Generator().manual_seed(c["seed"] + 2000)
    eps_item = torch.randn(I, Fw, generator=g)
    eps_ability = torch.randn(P, D, generator=g)
    queue = [eps_item, eps_ability]

    def injected(mean, logvar):
        eps = queue.pop(0)
        assert eps.shape == mean.shape
        return eps.mul(torch.exp(0.5 * logvar)).add_(mean)

    model.reparameterize_gaussian = injected
    out = model(response, mask)
    nf = c["n_flows"]
    if nf > 0:
        (_, _, response_mu, ability_k, ability, ability_mu, ability_logvar, a_ldj,
         item_k, item_feat, item_mu, item_lv, i_ldj) = out
        loss = model.elbo(response, mask, response_mu, ability, ability_mu, ability_logvar,
                          item_feat, item_mu, item_lv, annealing_factor=c["beta"],
                          use_kl_divergence=False, ability_k=ability_k, item_feat_k=item_k,
                          ability_logabsdetjac=a_ldj, item_logabsdetjac=i_ldj)
    else:
        (_, _, response_mu, ability, ability_mu, ability_logvar,
         item_feat, item_mu, item_lv) = out
        loss = model.elbo(*out, annealing_factor=c["beta"], use_kl_divergence=c["use_kl"])
    loss.backward()
    rec = dict(
        response=response.numpy()[:, :, 0], mask=mask_b.numpy()[:, :, 0].astype(np.uint8),
        eps_item=eps_item.numpy(), eps_ability=eps_ability.numpy(),
        loss=np.float64(loss.item()), response_mu=response_mu.detach().numpy()[:, :, 0],
        ability=ability.detach().numpy(), ability_mu=ability_mu.detach().numpy(),
        ability_logvar=ability_logvar.detach().numpy(), item_feat=item_feat.detach().numpy(),
    )
    if nf > 0:
        rec.update(ability_k=ability_k.detach().numpy(), item_feat_k=item_k.detach().numpy(),
                   ability_logabsdetjac=a_ldj.detach().numpy(),
                   item_feat_logabsdetjac=i_ldj.detach().numpy())
    for k, v in model.state_dict().items():
        rec["param/" + k] = v.detach().numpy()
    for k, v in model.named_parameters():
        rec["grad/" + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    # Same step in fp64 with the clamp kept at eps32 (an unpatched fp64 run
    # would clamp at 2.2e-16 and is not the same function, SURVEY.md finding 4):
    # separates "differs from the reference" from "the reference's own fp32 noise".
    import torch.distributions.utils as tdu
    orig_clamp = tdu.clamp_probs
    eps32 = torch.finfo(torch.float32).eps
    tdu.clamp_probs = lambda p: p.clamp(min=eps32, max=1 - eps32)
    try:
        model.double()
        model.zero_grad()
        queue[:] = [eps_item.double(), eps_ability.double()]
        out64 = model(response.double(), mask)
        if nf > 0:
            loss64 = model.elbo(out64[0], out64[1], out64[2], out64[4], out64[5], out64[6],
                                out64[9], out64[10], out64[11], annealing_factor=c["beta"],
                                use_kl_divergence=False, ability_k=out64[3], item_feat_k=out64[8],
                                ability_logabsdetjac=out64[7], item_logabsdetjac=out64[12])
        else:
            loss64 = model.elbo(*out64, annealing_factor=c["beta"], use_kl_divergence=c["use_kl"])
        loss64.backward()
        rec["loss64"] = np.float64(loss64.item())
        for k, v in model.named_parameters():
            g64 = v.grad if v.grad is not None else torch.zeros_like(v)
            rec["grad64/" + k] = g64.numpy().astype(np.float32)
    finally:
        tdu.clamp_probs = orig_clamp
    for k, v in init_state.items():      # only where the state was edited after seeded init
        if not np.array_equal(v, rec["param/" + k]):
            rec["init/" + k] = v
    return rec


def run_log_marginal(models):
    """models.py:445-504 with S injected noise pairs."""
    torch.manual_seed(77)
    I, P, D, S = 14, 11, 2, 5
    model = models.VIBO_2PL(D, I, ability_merge="product")
    response, mask_b = synth_responses(P, I, D, 2, 1077, 0.1)
    g = torch.Generator().manual_seed(2077)
    eps_items = [torch.randn(I, D + 1, generator=g) for _ in range(S)]
    eps_abils = [torch.randn(P, D, generator=g) for _ in range(S)]
    queue = []
    for a, b in zip(eps_items, eps_abils):
        queue += [a, b]

    def injected(mean, logvar):
        return queue.pop(0).mul(torch.exp(0.5 * logvar)).add_(mean)

    model.reparameterize_gaussian = injected
    logp = model.log_marginal(response, mask_b.long(), num_samples=S)
    rec = dict(response=response.numpy()[:, :, 0], mask=mask_b.numpy()[:, :, 0].astype(np.uint8),
               eps_items=np.stack([e.numpy() for e in eps_items]),
               eps_abilities=np.stack([e.numpy() for e in eps_abils]),
               logp=np.float64(logp.item()))
    for k, v in model.state_dict().items():
        rec["param/" + k] = v.detach().numpy()
    return rec


def run_vi(models):
    """Un-amortized VI_2PL (models.py:89-243): forward(index, response, mask) + elbo + backward."""
    torch.manual_seed(91)
    P_all, I, D, P = 40, 13, 2, 12
    model = models.VI_2PL(D, P_all, I)
    response, mask_b = synth_responses(P, I, D, 2, 1091, 0.15)
    index = torch.tensor([3, 7, 39, 0, 11, 12, 20, 21, 22, 5, 6, 30])
    g = torch.Generator().manual_seed(2091)
    eps_item, eps_ability = torch.randn(I, D + 1, generator=g), torch.randn(P, D, generator=g)
    queue = [eps_item, eps_ability]
    model.reparameterize_gaussian = lambda mean, logvar: queue.pop(0).mul(torch.exp(0.5 * logvar)).add_(mean)
    out = model(index, response, mask_b.long())
    loss = model.elbo(*out, annealing_factor=0.7, use_kl_divergence=True)
    loss.backward()
    rec = dict(index=index.numpy(), response=response.numpy()[:, :, 0], mask=mask_b.numpy()[:, :, 0].astype(np.uint8),
               eps_item=eps_item.numpy(), eps_ability=eps_ability.numpy(), loss=np.float64(loss.item()),
               response_mu=out[2].detach().numpy()[:, :, 0], beta=np.float64(0.7))
    for k, v in model.state_dict().items():
        rec["param/" + k] = v.detach().numpy()
    for k, v in model.named_parameters():
        rec["grad/" + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    return rec


def run_mask_fixture():
    """artificially_mask_dataset of the reference (src/datasets.py:46-78) on a toy
    dataset object; nltk is a dead import there (datasets.py:8) and is stubbed."""
    import types
    sys.modules.setdefault("nltk", types.SimpleNamespace(word_tokenize=None))
    from src import datasets as ref_ds
    rng = np.random.RandomState(3)
    P, I = 23, 11
    response = (rng.rand(P, I, 1) < 0.5).astype(np.float32)
    mask = np.ones((P, I, 1), dtype=np.float32)
    holes = rng.rand(P, I) < 0.1
    response[holes] = -1
    mask[holes] = 0
    ds = types.SimpleNamespace(response=response.copy(), mask=mask.copy())
    out = ref_ds.artificially_mask_dataset(ds, 0.2)
    return dict(response_in=response, mask_in=mask, response_out=out.response, mask_out=out.mask,
                missing_indices=out.missing_indices, missing_labels=out.missing_labels)


def main():
    models = _import_reference()
    index = []
    regenerate_all = "--all" in sys.argv   # default: only cases whose fixture file is missing
    for c in CASES:
        index.append(c)
        path = os.path.join(HERE, c["name"] + ".npz")
        if os.path.exists(path) and not regenerate_all:
            continue
        rec = run_case(models, c)
        np.savez_compressed(path, **rec)
        print(f"{c['name']:40s} loss={float(rec['loss']):.6f}")
    if regenerate_all or not os.path.exists(os.path.join(HERE, "log_marginal_2pl_d2.npz")):
        np.savez_compressed(os.path.join(HERE, "log_marginal_2pl_d2.npz"), **run_log_marginal(models))
    if regenerate_all or not os.path.exists(os.path.join(HERE, "vi_2pl_d2.npz")):
        np.savez_compressed(os.path.join(HERE, "vi_2pl_d2.npz"), **run_vi(models))
    if regenerate_all or not os.path.exists(os.path.join(HERE, "artificial_mask.npz")):
        np.savez_compressed(os.path.join(HERE, "artificial_mask.npz"), **run_mask_fixture())
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, reference=REF, cases=index), f, indent=1)


if __name__ == "__main__":
    main()
