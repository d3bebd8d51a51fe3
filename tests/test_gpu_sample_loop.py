"""Sample-loop kernels: the in-kernel IWAE ``log_marginal`` against the
live-reference fixture and against S separate fused passes, and the posterior-predictive mean
against S decodes.  Tolerance 1e-4 relative (BASELINE.json north star); observed ~1e-6."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _model(irt, D, I, dev, seed=3, drop=False):
    import vibo_b200
    torch.manual_seed(seed)
    cls = {1: vibo_b200.VIBO_1PL, 2: vibo_b200.VIBO_2PL, 3: vibo_b200.VIBO_3PL}[irt]
    m = cls(D, I, ability_merge="product", replace_missing_with_prior=not drop).to(dev)
    with torch.no_grad():   # keep the state off the eps32 clamp (see test_gpu_trainer)
        m.item_encoder.mu_lookup.weight.mul_(0.5)
        m.item_encoder.logvar_lookup.weight.mul_(0.2).sub_(2.0)
    return m


def test_log_marginal_kernel_matches_reference_fixture():
    import vibo_b200
    dev = torch.device("cuda:0")
    z = np.load(os.path.join(GOLDEN, "log_marginal_2pl_d2.npz"))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    model = vibo_b200.VIBO_2PL(2, z["response"].shape[1], ability_merge="product")
    model.load_state_dict(params)
    model = model.to(dev)
    response = torch.from_numpy(z["response"]).unsqueeze(2).to(dev)
    mask = torch.from_numpy(z["mask"]).bool().unsqueeze(2).to(dev)
    got = model.log_marginal(response, mask, num_samples=z["eps_items"].shape[0],
                             eps_item=torch.from_numpy(z["eps_items"]).to(dev),
                             eps_ability=torch.from_numpy(z["eps_abilities"]).to(dev))
    assert got.dim() == 0
    assert abs(float(got) - float(z["logp"])) <= 1e-5 * abs(float(z["logp"])), (float(got), float(z["logp"]))


@pytest.mark.parametrize("irt,D,I,P,S,miss,drop", [
    (2, 1, 100, 16, 40, 0.0, False),      # the CLI's shape: 16-person batch, samples spread over SMs
    (2, 2, 95, 77, 9, 0.15, False),
    (1, 1, 64, 300, 5, 0.1, True),
    (3, 3, 333, 130, 6, 0.2, False),
    (3, 1, 1000, 2500, 4, 0.0, False),    # several tiles per CTA
    (2, 8, 40, 33, 3, 0.0, False),
])
def test_log_marginal_kernel_equals_per_sample_passes(irt, D, I, P, S, miss, drop):
    dev = torch.device("cuda:0")
    model = _model(irt, D, I, dev, drop=drop)
    g = torch.Generator().manual_seed(P + I)
    resp = (torch.rand(P, I, 1, generator=g) < 0.5).float()
    mask = torch.rand(P, I, 1, generator=g) >= miss
    mask[:, 0] = True   # --drop-missing: no all-missing row
    resp[~mask] = -1.0
    F = model.item_feat_dim
    e_i = torch.randn(S, I, F, generator=g).to(dev)
    e_a = torch.randn(S, P, D, generator=g).to(dev)
    resp, mask = resp.to(dev), mask.to(dev)
    with torch.no_grad():
        logw = torch.stack([-model.fused_elbo(resp, mask, use_kl_divergence=False, eps_item=e_i[s],
                                              eps_ability=e_a[s]).double() for s in range(S)])
    want = torch.logsumexp(logw, 0) - np.log(S)
    from vibo_b200 import functional as VF
    r2, m2 = VF.prepare_rows(resp, mask)
    item_mu, item_lv = model.item_encoder()
    with torch.no_grad():
        logp, lw = VF.K.log_marginal(r2, m2, model.ability_encoder.expert_table(), item_mu, item_lv, S,
                                     irt_model=irt, missing_policy=model.ability_encoder.missing_policy,
                                     eps_item=e_i, eps_ability=e_a)
    assert torch.allclose(lw, logw, rtol=TOL, atol=1e-3), (lw, logw)
    assert abs(float(logp) - float(want)) <= TOL * abs(float(want))
    assert abs(float(model.log_marginal(resp, mask, S, eps_item=e_i, eps_ability=e_a)) - float(want)) \
        <= TOL * abs(float(want))


def test_log_marginal_philox_is_reproducible_and_sane():
    dev = torch.device("cuda:0")
    model = _model(2, 1, 100, dev)
    g = torch.Generator().manual_seed(1)
    resp = (torch.rand(64, 100, 1, generator=g) < 0.5).float().to(dev)
    mask = torch.ones(64, 100, 1, dtype=torch.bool, device=dev)
    a = float(model.log_marginal(resp, mask, 50, seed=7))
    b = float(model.log_marginal(resp, mask, 50, seed=7))
    c = float(model.log_marginal(resp, mask, 50, seed=8))
    assert a == b and a != c
    # IWAE bound >= ELBO in expectation: compare with the mean sample-form ELBO of fresh draws
    with torch.no_grad():
        elbo = np.mean([-float(model.fused_elbo(resp, mask, use_kl_divergence=False)) for _ in range(50)])
    assert a >= elbo - 0.02 * abs(elbo)
    assert abs(a - c) <= 0.05 * abs(a)


@pytest.mark.parametrize("irt,D,I,P", [(2, 1, 100, 50), (3, 2, 333, 41), (1, 1, 1000, 9)])
def test_predictive_mean_kernel(irt, D, I, P):
    """mean_s decode(theta_s, d_s): against torch on the kernel's own noise stream statistics --
    S large, so the Monte-Carlo mean is compared with a second, independent torch estimate."""
    dev = torch.device("cuda:0")
    model = _model(irt, D, I, dev)
    g = torch.Generator().manual_seed(5)
    resp = (torch.rand(P, I, 1, generator=g) < 0.5).float().to(dev)
    mask = torch.ones(P, I, 1, dtype=torch.bool, device=dev)
    S = 4000
    got = model.posterior_predictive_mean(resp, mask, S, seed=11)
    assert got.shape == (P, I, 1)
    with torch.no_grad():
        _, a_mu, a_lv, _, i_mu, i_lv = model.encode(resp, mask)
        acc = torch.zeros(P, I, 1, device=dev)
        for _ in range(S):
            ab = a_mu + torch.exp(0.5 * a_lv) * torch.randn_like(a_mu)
            it = i_mu + torch.exp(0.5 * i_lv) * torch.randn_like(i_mu)
            acc += model.decode(ab, it)
        want = acc / S
    # two independent S-sample means of values in [0, 1]: std error <= 0.5 / sqrt(S) each
    assert float((got - want).abs().max()) < 6 * 0.5 * np.sqrt(2.0 / S)
    assert float((got - want).abs().mean()) < 0.5 * np.sqrt(2.0 / S)
    # zero-variance posterior: the mean is the decode of the posterior means exactly
    from vibo_b200 import kernels as K
    zero = torch.full_like(a_lv, -80.0)
    out = K.predictive_mean(a_mu, zero, i_mu, torch.full_like(i_lv, -80.0), 3, irt_model=irt, seed=1)
    with torch.no_grad():
        exact = model.decode(a_mu, i_mu)[:, :, 0]
    assert torch.allclose(out, exact, atol=2e-6)
