"""Pin the oracle (oracle/reference_port.py, oracle/kernel_spec.py) against
fixtures produced by the LIVE reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import kernel_spec as KS
from oracle import reference_port as RP
from helpers import PORT_CASE_NAMES as CASE_NAMES, GOLDEN, case_inputs, load_case, max_rel, rel_l2


@pytest.mark.parametrize("name", CASE_NAMES)
def test_reference_port_matches_reference(name):
    """Same op sequence as the reference -> agreement at fp32 rounding level."""
    cfg, rec, params, grads = load_case(name)
    response, mask, eps_item, eps_ability = case_inputs(rec)
    loss, fw, g = RP.loss_and_grads(
        params, response, mask.long(), eps_item, eps_ability, irt_model=cfg["irt_model"],
        ability_dim=cfg["ability_dim"], conditional=cfg["conditional"], n_flows=cfg["n_flows"],
        replace_missing_with_prior=not cfg["drop_missing"], annealing_factor=cfg["beta"],
        use_kl_divergence=cfg["use_kl"])
    assert abs(loss.item() - rec["loss"]) <= 2e-6 * abs(rec["loss"])
    for k in ("ability_mu", "ability_logvar", "ability", "item_feat"):
        assert max_rel(fw[k].numpy(), rec[k]) < 1e-5, k
    assert max_rel(fw["response_mu"].numpy()[:, :, 0], rec["response_mu"]) < 1e-5
    for k, ref in grads.items():
        assert rel_l2(g[k].numpy(), ref) < 2e-5, k


@pytest.mark.parametrize("name", [n for n in CASE_NAMES if "flows" not in n and "_mean_" not in n])
def test_kernel_spec_matches_reference(name):
    """Closed-form fp64 kernel spec, chained through torch autograd for the
    tiny parameter-side pieces exactly as the product does, reproduces the
    reference's loss and parameter gradients (1e-4 is the north-star bar; the
    observed error is the reference's own fp32 noise)."""
    cfg, rec, params, grads = load_case(name)
    irt, D, cond = cfg["irt_model"], cfg["ability_dim"], cfg["conditional"]
    P64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    eps_item = torch.from_numpy(rec["eps_item"]).double()
    item_mu, item_lv = P64["item_encoder.mu_lookup.weight"], P64["item_encoder.logvar_lookup.weight"]
    item_feat = eps_item * torch.exp(0.5 * item_lv) + item_mu
    I = cfg["I"]
    if cond:
        r = torch.zeros(2, I, 1, dtype=torch.float64)
        r[1] = 1
        rows = torch.cat([r, item_feat.unsqueeze(0).expand(2, I, -1)], 2).reshape(2 * I, -1)
        table = RP.encoder_mlp(P64, rows).reshape(2, I, 2 * D)
    else:
        table = RP.encoder_mlp(P64, torch.tensor([[0.0], [1.0]], dtype=torch.float64)).reshape(2, 1, 2 * D)
    form = KS.ELBO_KL if cfg["use_kl"] else KS.ELBO_SAMPLE
    out = KS.fused_elbo(rec["response"].astype(np.float64), rec["mask"], table.detach().numpy(),
                        item_feat.detach().numpy(), rec["eps_ability"].astype(np.float64),
                        irt_model=irt, beta=cfg["beta"], elbo_form=form,
                        missing_policy=KS.MISSING_DROP if cfg["drop_missing"] else KS.MISSING_PRIOR)
    if cfg["use_kl"]:
        item_term = cfg["beta"] * RP.kl_standard_normal(item_mu, item_lv).sum()
    else:
        item_term = -(RP.standard_normal_log_pdf(item_feat).sum()
                      - RP.normal_log_pdf(item_feat, item_mu, item_lv).sum())
    loss = out["loss_k"] + item_term.item()
    assert abs(loss - rec["loss"]) <= 1e-5 * abs(rec["loss"]), (loss, rec["loss"])
    assert abs(loss - rec["loss64"]) <= 1e-10 * abs(rec["loss64"]), (loss, rec["loss64"])
    for k in ("ability_mu", "ability_logvar", "ability"):
        assert max_rel(out[k], rec[k]) < 1e-4, k
    # parameter gradients: kernel grads -> autograd through the small chains
    surrogate = (table * torch.from_numpy(out["g_table"])).sum() \
        + (item_feat * torch.from_numpy(out["g_item"]).double()).sum() + item_term
    surrogate.backward()
    for k, ref in grads.items():
        got = P64[k].grad.numpy() if P64[k].grad is not None else np.zeros_like(ref)
        ref64 = rec["grad64/" + k]
        # exact math vs the reference evaluated in fp64 (fixture stored as f32)
        assert rel_l2(got, ref64) < 1e-6, (k, rel_l2(got, ref64))
        # vs the fp32 reference: 1e-4, widened only by the reference's own
        # fp32 noise on saturating 3PL states ("knife-edge")
        noise = rel_l2(ref, ref64)
        assert rel_l2(got, ref) < 1e-4 + noise, (k, rel_l2(got, ref), noise)


def test_log_marginal_port_matches_reference():
    z = np.load(os.path.join(GOLDEN, "log_marginal_2pl_d2.npz"))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    response = torch.from_numpy(z["response"]).unsqueeze(2)
    mask = torch.from_numpy(z["mask"]).long().unsqueeze(2)
    logp = RP.log_marginal(params, response, mask, list(torch.from_numpy(z["eps_items"])),
                           list(torch.from_numpy(z["eps_abilities"])), irt_model=2, ability_dim=2)
    assert abs(logp.item() - float(z["logp"])) <= 2e-6 * abs(float(z["logp"]))


def test_clamp_single_cells():
    """Value floor -15.9424 and zero gradient outside the eps32 clamp
    (SURVEY.md Appendix B)."""
    resp = np.array([[1.0, 0.0, 1.0, 0.0]])
    mask = np.ones_like(resp, dtype=np.uint8)
    theta = np.zeros((1, 1))
    item = np.array([[0.0, -20.0], [0.0, 20.0], [0.0, 20.0], [0.0, -20.0]])  # z = b
    r = KS.link_loglik(resp, mask, theta, item, 2)
    ll = KS.bernoulli_loglik(resp, mask, KS.decode(theta, item, 2))[0][0]
    assert np.allclose(ll[:2], np.log(KS.EPS32), rtol=1e-6)
    assert np.allclose(r["dz"][0, :2], 0.0)
    assert np.allclose(ll[2:], np.log1p(-KS.EPS32), rtol=1e-3)


def test_person_counts_equivalence():
    """Unconditional posterior depends on the row only through its counts."""
    rng = np.random.default_rng(0)
    P, I, D = 13, 37, 2
    resp = (rng.random((P, I)) < 0.4).astype(np.float64)
    mask = (rng.random((P, I)) < 0.8).astype(np.uint8)
    table = rng.normal(size=(2, 1, 2 * D))
    enc = KS.encode(resp, mask, table, D)
    n0, n1, nm = KS.person_counts(resp, mask)
    mu, _, tau = KS.expert_precision(table, D)
    S = n0[:, None] * tau[0, 0] + n1[:, None] * tau[1, 0] + nm[:, None] / (1 + 1e-8)
    assert np.allclose(S, enc["S"])
    # permutation invariance of the product of experts
    perm = rng.permutation(I)
    enc2 = KS.encode(resp[:, perm], mask[:, perm], table, D)
    assert np.allclose(enc2["ability_mu"], enc["ability_mu"])


@pytest.mark.parametrize("D,K", [(1, 2), (3, 1), (4, 5)])
def test_flow_person_spec_matches_reference_port(D, K):
    """oracle/kernel_spec.flow_person (the closed form the CUDA flow kernels implement) against
    the literal port's planar flows under float64 autograd."""
    torch.manual_seed(D * 10 + K)
    P = 37
    p = {}
    for k in range(K):
        p[f"ability_norm_flows.flows.{k}.u"] = torch.randn(D, dtype=torch.float64, requires_grad=True)
        p[f"ability_norm_flows.flows.{k}.w"] = torch.randn(D, dtype=torch.float64, requires_grad=True)
        p[f"ability_norm_flows.flows.{k}.b"] = torch.ones(1, dtype=torch.float64, requires_grad=True)
    mu = torch.randn(P, D, dtype=torch.float64, requires_grad=True)
    lv = (0.3 * torch.randn(P, D, dtype=torch.float64) - 1.0).requires_grad_()
    eps = torch.randn(P, D, dtype=torch.float64)
    w_ll = torch.randn(P, D, dtype=torch.float64)
    t0 = RP.reparameterize(mu, lv, eps)
    tk, ldj = RP.planar_flows(p, "ability_norm_flows", K, t0)
    term = RP.standard_normal_log_pdf(tk).sum() - (RP.normal_log_pdf(t0, mu, lv).sum() - ldj.sum())
    loss = -((tk * w_ll).sum() + 0.7 * term)
    loss.backward()
    # uhat as the product forms it (flows.py:26-29), with autograd to map g_uhat back to (u, w)
    us = [p[f"ability_norm_flows.flows.{k}.u"].detach().clone().requires_grad_() for k in range(K)]
    ws = [p[f"ability_norm_flows.flows.{k}.w"].detach().clone().requires_grad_() for k in range(K)]
    uhat = torch.stack([u + (torch.nn.functional.softplus((u * w).sum()) - 1.0 - (u * w).sum()) * w / (w * w).sum()
                        for u, w in zip(us, ws)])
    out = KS.flow_person(mu.detach().numpy(), lv.detach().numpy(), eps.numpy(), uhat.detach().numpy(),
                         torch.stack(ws).detach().numpy(), np.ones(K), g_ability_k=-w_ll.numpy(), g_term=-0.7)
    assert np.allclose(out["ability_k"], tk.detach().numpy(), rtol=1e-12, atol=1e-12)
    assert abs(out["term"] - term.item()) < 1e-9 * max(1.0, abs(term.item()))
    assert np.allclose(out["g_mu"], mu.grad.numpy(), rtol=1e-9, atol=1e-11)
    assert np.allclose(out["g_logvar"], lv.grad.numpy(), rtol=1e-9, atol=1e-11)
    (uhat * torch.from_numpy(out["g_uhat"])).sum().backward()
    for k in range(K):
        g_u = us[k].grad.numpy()
        g_w = ws[k].grad.numpy() + out["g_w"][k]
        assert np.allclose(g_u, p[f"ability_norm_flows.flows.{k}.u"].grad.numpy(), rtol=1e-8, atol=1e-10)
        assert np.allclose(g_w, p[f"ability_norm_flows.flows.{k}.w"].grad.numpy(), rtol=1e-8, atol=1e-10)
        assert abs(out["g_b"][k] - p[f"ability_norm_flows.flows.{k}.b"].grad.item()) < 1e-8 * max(1, abs(out["g_b"][k]))


@pytest.mark.parametrize("name", [n for n in CASE_NAMES if "_mean_" in n])
def test_mean_merge_spec_matches_reference(name):
    """Table-collapsed masked mean (oracle/kernel_spec.mean_merge_hidden) + mlp2 reproduce the
    live reference's ability posterior for --ability-merge mean."""
    cfg, rec, params, _ = load_case(name)
    P64 = {k: v.double() for k, v in params.items()}
    I, cond = cfg["I"], cfg["conditional"]
    F = torch.nn.functional
    if cond:
        item_feat = torch.from_numpy(rec["item_feat"]).double()
        r = torch.zeros(2, I, 1, dtype=torch.float64)
        r[1] = 1
        rows = torch.cat([r, item_feat.unsqueeze(0).expand(2, I, -1)], 2).reshape(2 * I, -1)
    else:
        rows = torch.tensor([[0.0], [1.0]], dtype=torch.float64)
    h = F.elu(F.linear(rows, P64["ability_encoder.mlp1.0.weight"], P64["ability_encoder.mlp1.0.bias"]))
    h = F.elu(F.linear(h, P64["ability_encoder.mlp1.2.weight"], P64["ability_encoder.mlp1.2.bias"]))
    table = h.reshape(2, I if cond else 1, -1).numpy()
    hid_mean = torch.from_numpy(KS.mean_merge_hidden(rec["response"], rec["mask"], table))
    o = F.elu(F.linear(hid_mean, P64["ability_encoder.mlp2.0.weight"], P64["ability_encoder.mlp2.0.bias"]))
    o = F.linear(o, P64["ability_encoder.mlp2.2.weight"], P64["ability_encoder.mlp2.2.bias"])
    mu, lv = torch.chunk(o, 2, dim=1)
    assert max_rel(mu.numpy(), rec["ability_mu"]) < 1e-5
    assert max_rel(lv.numpy(), rec["ability_logvar"]) < 1e-5


@pytest.mark.parametrize("D,policy,missing", [(1, 0, 0.0), (1, 0, 0.1), (2, 1, 0.3), (5, 0, 0.2)])
def test_count_based_unconditional_encode_backward_spec(D, policy, missing):
    """oracle.kernel_spec.encode_backward_counts (the spec of vibo_encode_backward_counts: table gradient of the
    unconditional posterior from per-person counts) equals the row-level encode_backward it replaces."""
    rng = np.random.default_rng(11 + D)
    P, I = 257, 95
    resp = (rng.random((P, I)) < 0.55).astype(np.float64)
    mask = (rng.random((P, I)) >= missing).astype(np.uint8)
    resp[mask == 0] = -1.0
    table = 0.4 * rng.normal(size=(2, 1, 2 * D))
    enc = KS.encode(resp, mask, table, D, policy)
    g_mu, g_lv = rng.normal(size=(P, D)), rng.normal(size=(P, D))
    ref = KS.encode_backward(resp, mask, table, D, enc["S"], enc["ability_mu"], g_mu, g_lv)
    n0, n1, _ = KS.person_counts(resp, mask)
    got = KS.encode_backward_counts(np.stack([n1, n0 + n1], 1), table, D, enc["S"], enc["ability_mu"], g_mu, g_lv)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("K,D", [(1, 1), (2, 1), (2, 2), (8, 7)])
def test_planar_params_spec_matches_reference_flow_correction(K, D):
    """oracle.kernel_spec.planar_params (spec of vibo_planar_params_forward / _backward) against float64 autograd
    of the reference's invertibility correction (flows.py:26-29), including a flow past the softplus threshold."""
    g = torch.Generator().manual_seed(K * 10 + D)
    u = torch.randn(K, D, generator=g, dtype=torch.float64)
    w = torch.randn(K, D, generator=g, dtype=torch.float64)
    if K > 1:
        u[0], w[0] = 6.0, 4.0   # w.u = 24 D > 20: softplus is the identity there
    u.requires_grad_(True)
    w.requires_grad_(True)
    uw = (u * w).sum(1, keepdim=True)
    uhat = u + (torch.nn.functional.softplus(uw) - 1.0 - uw) * w / (w * w).sum(1, keepdim=True)
    g_uhat = torch.randn(K, D, generator=g, dtype=torch.float64)
    g_wout = torch.randn(K, D, generator=g, dtype=torch.float64)
    (uhat * g_uhat).sum().add((w * g_wout).sum()).backward()
    got_uhat, got_gu, got_gw = KS.planar_params(u.detach().numpy(), w.detach().numpy(), g_uhat.numpy(), g_wout.numpy())
    assert np.allclose(got_uhat, uhat.detach().numpy(), rtol=1e-12, atol=1e-12)
    assert np.allclose(got_gu, u.grad.numpy(), rtol=1e-10, atol=1e-12)
    assert np.allclose(got_gw, w.grad.numpy(), rtol=1e-10, atol=1e-12)
