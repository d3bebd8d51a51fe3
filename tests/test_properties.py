"""hypothesis property tests of the closed-form kernel spec (CPU): the
invariants the CUDA kernels rely on (SURVEY.md 4 iii)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import kernel_spec as KS


def _case(rng, P, I, D, irt, cond, missing):
    F = KS.item_feat_width(irt, D)
    resp = (rng.random((P, I)) < 0.5).astype(np.float64)
    mask = (rng.random((P, I)) >= missing).astype(np.uint8)
    resp[mask == 0] = -1.0
    table = 0.5 * rng.normal(size=(2, I if cond else 1, 2 * D))
    item = 0.8 * rng.normal(size=(I, F))
    eps = rng.normal(size=(P, D))
    return resp, mask, table, item, eps


shape = st.tuples(st.integers(2, 9), st.integers(1, 12), st.integers(1, 3), st.sampled_from([1, 2, 3]),
                  st.booleans(), st.sampled_from([0.0, 0.3]), st.integers(0, 10 ** 6))


@settings(max_examples=40, deadline=None)
@given(shape)
def test_shard_sums_equal_whole(s):
    P, I, D, irt, cond, missing, seed = s
    rng = np.random.default_rng(seed)
    resp, mask, table, item, eps = _case(rng, P, I, D, irt, cond, missing)
    whole = KS.fused_elbo(resp, mask, table, item, eps, irt_model=irt, beta=0.6)
    cut = P // 2
    a = KS.fused_elbo(resp[:cut], mask[:cut], table, item, eps[:cut], irt_model=irt, beta=0.6)
    b = KS.fused_elbo(resp[cut:], mask[cut:], table, item, eps[cut:], irt_model=irt, beta=0.6)
    assert np.isclose(a["ll"] + b["ll"], whole["ll"], rtol=1e-10)
    assert np.allclose(a["g_item"] + b["g_item"], whole["g_item"], rtol=1e-9, atol=1e-12)
    assert np.allclose(a["g_table"] + b["g_table"], whole["g_table"], rtol=1e-9, atol=1e-12)


@settings(max_examples=40, deadline=None)
@given(shape)
def test_item_permutation_invariance(s):
    """PoE and the log-likelihood sum do not depend on item order (with the
    per-item table / parameters permuted consistently)."""
    P, I, D, irt, cond, missing, seed = s
    rng = np.random.default_rng(seed)
    resp, mask, table, item, eps = _case(rng, P, I, D, irt, cond, missing)
    perm = rng.permutation(I)
    t2 = table[:, perm] if cond else table
    x = KS.fused_elbo(resp, mask, table, item, eps, irt_model=irt)
    y = KS.fused_elbo(resp[:, perm], mask[:, perm], t2, item[perm], eps, irt_model=irt)
    assert np.isclose(x["ll"], y["ll"], rtol=1e-10)
    assert np.allclose(x["ability_mu"], y["ability_mu"], rtol=1e-9, atol=1e-12)
    assert np.allclose(x["g_item"][perm], y["g_item"], rtol=1e-9, atol=1e-12)


@settings(max_examples=25, deadline=None)
@given(shape)
def test_gradients_match_finite_differences(s):
    P, I, D, irt, cond, missing, seed = s
    rng = np.random.default_rng(seed)
    resp, mask, table, item, eps = _case(rng, P, I, D, irt, cond, missing)
    base = KS.fused_elbo(resp, mask, table, item, eps, irt_model=irt, beta=0.8)
    h = 1e-6
    j, f = rng.integers(I), rng.integers(item.shape[1])
    it2 = item.copy(); it2[j, f] += h
    fd = (KS.fused_elbo(resp, mask, table, it2, eps, irt_model=irt, beta=0.8, want_grads=False)["loss_k"]
          - base["loss_k"]) / h
    assert np.isclose(fd, base["g_item"][j, f], rtol=2e-3, atol=2e-4)
    r, jt, k = rng.integers(2), rng.integers(table.shape[1]), rng.integers(2 * D)
    t2 = table.copy(); t2[r, jt, k] += h
    fd = (KS.fused_elbo(resp, mask, t2, item, eps, irt_model=irt, beta=0.8, want_grads=False)["loss_k"]
          - base["loss_k"]) / h
    assert np.isclose(fd, base["g_table"][r, jt, k], rtol=2e-3, atol=2e-4)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(1, 10), st.integers(1, 3))
def test_masked_cells_do_not_matter(seed, I, D):
    """whatever value a missing cell holds, nothing changes"""
    rng = np.random.default_rng(seed)
    resp, mask, table, item, eps = _case(rng, 5, I, D, 2, False, 0.4)
    a = KS.fused_elbo(resp, mask, table, item, eps, irt_model=2)
    resp2 = resp.copy(); resp2[mask == 0] = rng.choice([0.0, 1.0, -1.0, 7.0], size=int((mask == 0).sum()))
    b = KS.fused_elbo(resp2, mask, table, item, eps, irt_model=2)
    assert a["ll"] == b["ll"] and np.array_equal(a["g_item"], b["g_item"])


@settings(max_examples=25, deadline=None)
@given(st.tuples(st.integers(1, 4), st.integers(1, 4), st.integers(0, 10 ** 6)))
def test_flow_person_gradients_match_finite_differences(s):
    """Closed-form backward of the fused planar-flow person kernels (oracle/kernel_spec.flow_person)
    against central differences of its own forward, for the loss  w.theta_K + c * term."""
    D, K, seed = s
    rng = np.random.default_rng(seed)
    P = 5
    mu, lv, eps = rng.normal(size=(P, D)), 0.3 * rng.normal(size=(P, D)) - 1.0, rng.normal(size=(P, D))
    uhat, w, b = 0.6 * rng.normal(size=(K, D)), 0.6 * rng.normal(size=(K, D)), rng.normal(size=K)
    wll, c = rng.normal(size=(P, D)), 0.7

    def loss(mu_, lv_, uhat_, w_, b_):
        o = KS.flow_person(mu_, lv_, eps, uhat_, w_, b_)
        return float((o["ability_k"] * wll).sum() + c * o["term"])

    out = KS.flow_person(mu, lv, eps, uhat, w, b, g_ability_k=wll, g_term=c)
    h = 1e-6
    for name, arr, g in (("mu", mu, out["g_mu"]), ("lv", lv, out["g_logvar"]), ("uhat", uhat, out["g_uhat"]),
                         ("w", w, out["g_w"]), ("b", b, out["g_b"])):
        idx = tuple(rng.integers(0, n) for n in arr.shape)
        args = dict(mu_=mu, lv_=lv, uhat_=uhat, w_=w, b_=b)
        key = {"mu": "mu_", "lv": "lv_", "uhat": "uhat_", "w": "w_", "b": "b_"}[name]
        up, dn = arr.copy(), arr.copy()
        up[idx] += h
        dn[idx] -= h
        fd = (loss(**{**args, key: up}) - loss(**{**args, key: dn})) / (2 * h)
        assert np.isclose(fd, g[idx], rtol=2e-5, atol=2e-6), (name, fd, g[idx])
