"""Host logic of the drop-in modules (autograd glue, expert table, flows, API
tuple layouts, seeded init) on CPU, with the kernel wrappers swapped for the
numpy oracle (tests/oracle_backend.py).  Compared with the live-reference
fixtures."""
import numpy as np
import pytest
import torch

import oracle_backend
from helpers import CASE_NAMES, build_model, case_inputs, load_case, max_rel, rel_l2


@pytest.fixture(autouse=True)
def _oracle_kernels(monkeypatch):
    oracle_backend.install(monkeypatch)


def _check_grads(model, rec, grads):
    for k, ref in grads.items():
        p = dict(model.named_parameters())[k]
        got = p.grad.numpy() if p.grad is not None else np.zeros_like(ref)
        ref64 = rec["grad64/" + k]
        noise = rel_l2(ref, ref64)
        assert rel_l2(got, ref64) < 1e-4, (k, rel_l2(got, ref64))  # fp32 host chain
        assert rel_l2(got, ref) < 1e-4 + noise, (k, rel_l2(got, ref), noise)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_fused_elbo_matches_reference(name):
    cfg, rec, params, grads = load_case(name)
    model = build_model(cfg, params)
    response, mask, eps_item, eps_ability = case_inputs(rec)
    loss, out = model.fused_elbo(response, mask, annealing_factor=cfg["beta"],
                                 use_kl_divergence=cfg["use_kl"], eps_item=eps_item,
                                 eps_ability=eps_ability, return_outputs=True)
    loss.backward()
    assert abs(loss.item() - rec["loss"]) <= 1e-5 * abs(rec["loss"])
    for k in ("ability_mu", "ability_logvar", "ability", "item_feat"):
        assert max_rel(out[k].detach().numpy(), rec[k]) < 1e-4, k
    _check_grads(model, rec, grads)


@pytest.mark.parametrize("name", ["m2pl_d1_unc_full", "m3pl_d3_cond_miss", "m1pl_d3_unc_miss",
                                  "m2pl_d1_unc_flows2_miss", "m3pl_d2_cond_flows2",
                                  "m2pl_d1_unc_sampleform"])
def test_forward_elbo_api_matches_reference(name):
    """The reference call pattern (vibo.py:264-266): model(response, mask) then
    model.elbo(*outputs), with noise injected the way the golden generator
    injects it into the reference."""
    cfg, rec, params, grads = load_case(name)
    model = build_model(cfg, params)
    response, mask, eps_item, eps_ability = case_inputs(rec)
    queue = [eps_item, eps_ability]
    model.reparameterize_gaussian = lambda mean, logvar: queue.pop(0) * torch.exp(0.5 * logvar) + mean
    out = model(response, mask.long())
    if cfg["n_flows"] > 0:
        assert len(out) == 13
        (_, _, response_mu, ability_k, ability, a_mu, a_lv, a_ldj, item_k, item_feat, i_mu, i_lv, i_ldj) = out
        loss = model.elbo(response, mask.long(), response_mu, ability, a_mu, a_lv, item_feat, i_mu, i_lv,
                          annealing_factor=cfg["beta"], use_kl_divergence=False, ability_k=ability_k,
                          item_feat_k=item_k, ability_logabsdetjac=a_ldj, item_logabsdetjac=i_ldj)
        assert max_rel(ability_k.detach().numpy(), rec["ability_k"]) < 1e-5
        assert max_rel(a_ldj.detach().numpy(), rec["ability_logabsdetjac"]) < 1e-4
        assert max_rel(i_ldj.detach().numpy(), rec["item_feat_logabsdetjac"]) < 1e-4
    else:
        assert len(out) == 9
        response_mu = out[2]
        loss = model.elbo(*out, annealing_factor=cfg["beta"], use_kl_divergence=cfg["use_kl"])
    assert response_mu.shape == (cfg["P"], cfg["I"], 1)
    assert max_rel(response_mu.detach().numpy()[:, :, 0], rec["response_mu"]) < 1e-5
    loss.backward()
    assert loss.dim() == 0
    assert abs(loss.item() - rec["loss"]) <= 1e-5 * abs(rec["loss"])
    _check_grads(model, rec, grads)


@pytest.mark.parametrize("name", ["m2pl_d1_unc_full", "m3pl_d3_cond_full", "m1pl_d2_unc_flows1",
                                  "m3pl_d2_cond_flows2"])
def test_seeded_init_matches_reference(name):
    """torch.manual_seed(s); VIBO_*PL(...) gives the reference's initial
    weights bit for bit (same containers built in the same order)."""
    import vibo_b200
    cfg, rec, params, _ = load_case(name)
    torch.manual_seed(cfg["seed"])
    cls = {1: vibo_b200.VIBO_1PL, 2: vibo_b200.VIBO_2PL, 3: vibo_b200.VIBO_3PL}[cfg["irt_model"]]
    model = cls(cfg["ability_dim"], cfg["I"], ability_merge="product",
                conditional_posterior=cfg["conditional"], n_norm_flows=cfg["n_flows"])
    sd = model.state_dict()
    assert list(sd) == list(params), "state_dict keys / order differ from the reference"
    for k, v in sd.items():
        ref = rec.get("init/" + k, rec["param/" + k])
        assert np.array_equal(v.numpy(), ref), k


def test_log_marginal_and_shapes():
    import os
    from helpers import GOLDEN
    import vibo_b200
    z = np.load(os.path.join(GOLDEN, "log_marginal_2pl_d2.npz"))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    model = vibo_b200.VIBO_2PL(2, z["response"].shape[1], ability_merge="product")
    model.load_state_dict(params)
    response = torch.from_numpy(z["response"]).unsqueeze(2)
    mask = torch.from_numpy(z["mask"]).bool().unsqueeze(2)
    S = z["eps_items"].shape[0]
    with torch.no_grad():
        logw = torch.stack([-model.fused_elbo(response, mask, use_kl_divergence=False,
                                              eps_item=torch.from_numpy(z["eps_items"][s]),
                                              eps_ability=torch.from_numpy(z["eps_abilities"][s]))
                            for s in range(S)])
    logp = torch.logsumexp(logw, 0) - np.log(S)
    assert logp.dim() == 0
    assert abs(logp.item() - float(z["logp"])) <= 1e-5 * abs(float(z["logp"]))
    assert model.log_marginal(response, mask, num_samples=3).dim() == 0


def test_item_kl_charged_in_full_per_batch():
    """Loss is a sum and the full item KL is added on every minibatch
    (reference models.py:399, :428-430)."""
    cfg, rec, params, _ = load_case("m2pl_d1_unc_full")
    model = build_model(cfg, params)
    response, mask, eps_item, eps_ability = case_inputs(rec)
    with torch.no_grad():
        full = model.fused_elbo(response, mask, eps_item=eps_item, eps_ability=eps_ability)
        a = model.fused_elbo(response[:10], mask[:10], eps_item=eps_item, eps_ability=eps_ability[:10])
        b = model.fused_elbo(response[10:], mask[10:], eps_item=eps_item, eps_ability=eps_ability[10:])
        from vibo_b200.models import kl_divergence_standard_normal_prior as kl
        kl_item = kl(*model.item_encoder()).sum()
    assert abs((a + b - kl_item).item() - full.item()) < 1e-3


def test_constructor_options():
    import vibo_b200
    m = vibo_b200.VIBO_2PL(1, 5)  # class default ability_merge='mean' (reference models.py:252)
    assert sorted(k for k in m.state_dict() if k.startswith("ability_encoder")) == sorted(
        f"ability_encoder.{n}.{i}.{w}" for n in ("mlp1", "mlp2") for i in (0, 2) for w in ("weight", "bias"))
    d = vibo_b200.VIBO_2PL(1, 5, ability_merge="product", generative_model="deep")
    assert any(k.startswith("decoder.mlp_concat.") for k in d.state_dict())
    with pytest.raises(AssertionError):
        vibo_b200.VIBO_2PL(1, 5, ability_merge="transformer")
    with pytest.raises(ValueError):
        vibo_b200.VIBO_2PL(9, 5)       # ability_dim beyond the kernels' limit: fail at construction
    with pytest.raises(ValueError):
        vibo_b200.VIBO_2PL(1, 5000)    # item bank wider than the kernels accept


@pytest.mark.parametrize("name", ["m2pl_d1_unc_link_full", "m1pl_d1_unc_deep_full", "m2pl_d1_unc_residual_full",
                                  "m3pl_d2_unc_residual_miss", "m2pl_d1_unc_gauss_full"])
def test_seeded_init_matches_reference_decoders(name):
    """Same as test_seeded_init_matches_reference for the nonlinear generative models: the decoder
    classes re-initialise themselves inside their constructors (models.py:786, :848, :887) before the
    outer apply(weights_init), and the RNG stream has to be consumed identically."""
    import vibo_b200
    cfg, rec, params, _ = load_case(name)
    torch.manual_seed(cfg["seed"])
    cls = {1: vibo_b200.VIBO_1PL, 2: vibo_b200.VIBO_2PL, 3: vibo_b200.VIBO_3PL}[cfg["irt_model"]]
    model = cls(cfg["ability_dim"], cfg["I"], ability_merge=cfg.get("merge", "product"),
                conditional_posterior=cfg["conditional"], generative_model=cfg.get("generative", "irt"),
                response_dist=cfg.get("response_dist", "bernoulli"))
    sd = model.state_dict()
    assert list(sd) == list(params), "state_dict keys / order differ from the reference"
    for k, v in sd.items():
        ref = rec.get("init/" + k, rec["param/" + k])
        assert np.array_equal(v.numpy(), ref), k


def test_vi_module_matches_reference():
    """Un-amortized VI_2PL (reference models.py:89-243) against the live-reference fixture."""
    import os
    from helpers import GOLDEN
    import vibo_b200
    z = np.load(os.path.join(GOLDEN, "vi_2pl_d2.npz"))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    model = vibo_b200.VI_2PL(2, 40, z["response"].shape[1])
    assert list(model.state_dict()) == list(params)
    model.load_state_dict(params)
    index = torch.from_numpy(z["index"])
    response = torch.from_numpy(z["response"]).unsqueeze(2)
    mask = torch.from_numpy(z["mask"]).bool().unsqueeze(2)
    for fused in (False, True):
        model.zero_grad()
        queue = [torch.from_numpy(z["eps_item"]), torch.from_numpy(z["eps_ability"])]
        model.reparameterize_gaussian = lambda mean, logvar: queue.pop(0) * torch.exp(0.5 * logvar) + mean
        if fused:
            loss = model.fused_elbo(index, response, mask, annealing_factor=float(z["beta"]))
        else:
            out = model(index, response, mask.long())
            assert max_rel(out[2].detach().numpy()[:, :, 0], z["response_mu"]) < 1e-5
            loss = model.elbo(*out, annealing_factor=float(z["beta"]))
        loss.backward()
        assert abs(loss.item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
        for k, p in model.named_parameters():
            assert rel_l2(p.grad.numpy(), z["grad/" + k]) < 1e-4, (fused, k)


@pytest.mark.parametrize("K,D", [(1, 1), (2, 2), (3, 5)])
def test_planar_params_function_host_logic(monkeypatch, K, D):
    """functional.PlanarParams (the K flows' separate u, w, b -> stacked uhat, w, b; one kernel each way on the
    GPU) on the oracle backend: outputs and every parameter gradient equal PyTorch autograd of the reference
    formulation (flows.py:26-29) -- argument order, the per-parameter gradient views and the b pass-through."""
    import oracle_backend
    import vibo_b200
    from vibo_b200 import functional as VF
    from vibo_b200.flows import NormalizingFlows
    oracle_backend.install(monkeypatch)
    torch.manual_seed(K + 10 * D)
    nf = NormalizingFlows(D, n_flows=K)
    ref_uhat, ref_w, ref_b = nf.stacked_parameters()          # CPU parameters: the autograd formulation
    g = [torch.randn_like(ref_uhat), torch.randn_like(ref_w), torch.randn_like(ref_b)]
    (ref_uhat * g[0]).sum().add((ref_w * g[1]).sum()).add((ref_b * g[2]).sum()).backward()
    ref_grads = [[f.u.grad.clone(), f.w.grad.clone(), f.b.grad.clone()] for f in nf.flows]
    nf.zero_grad()
    uhat, w, b = VF.PlanarParams.apply(K, *[f.u for f in nf.flows], *[f.w for f in nf.flows], *[f.b for f in nf.flows])
    assert torch.allclose(uhat, ref_uhat.detach(), atol=1e-6) and torch.equal(w, ref_w.detach())
    assert torch.equal(b, ref_b.detach())
    (uhat * g[0]).sum().add((w * g[1]).sum()).add((b * g[2]).sum()).backward()
    for f, (gu, gw, gb) in zip(nf.flows, ref_grads):
        assert torch.allclose(f.u.grad, gu, rtol=1e-5, atol=1e-6)
        assert torch.allclose(f.w.grad, gw, rtol=1e-5, atol=1e-6)
        assert torch.allclose(f.b.grad, gb)
