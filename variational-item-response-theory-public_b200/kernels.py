"""Tensor-level wrappers of the C ABI (one Python function per entry point).

Every function takes/returns CUDA tensors, enqueues on the current torch
stream and never synchronises.  ``tests/`` swaps this module's functions for a
numpy oracle to exercise the host logic on machines without a GPU; the product
itself has no such fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import Desc, ELBO_KL, ELBO_SAMPLE, MISSING_DROP, MISSING_PRIOR  # noqa: F401

_workspaces = {}
ERR_UNSUPPORTED = -2   # VIBO_ERR_UNSUPPORTED (include/vibo_b200.h)


def item_feat_width(irt_model: int, ability_dim: int) -> int:
    return {1: 1, 2: ability_dim + 1, 3: ability_dim + 2}[irt_model]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_rows(response: torch.Tensor, mask: torch.Tensor):
    if not response.is_cuda:
        raise _lib.ViboError("VIBO kernels need CUDA tensors (no CPU fallback exists)")
    assert response.dtype == torch.float32 and response.dim() == 2 and response.is_contiguous()
    assert mask.dtype == torch.uint8 and mask.shape == response.shape and mask.is_contiguous()


def make_desc(P, I, D, irt_model, conditional, missing_policy=MISSING_PRIOR, elbo_form=ELBO_KL,
              person_offset=0) -> Desc:
    return Desc(int(P), int(I), int(D), int(irt_model), int(bool(conditional)), int(missing_policy),
                int(elbo_form), int(person_offset))


def workspace(desc: Desc, device) -> torch.Tensor:
    """Per-(device, stream) scratch, grown on demand and reused across calls."""
    need = int(_lib.load().vibo_workspace_bytes(C.byref(desc)))
    key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def fused_elbo(response, mask, table, item_feat, eps_ability, *, irt_model, conditional,
               missing_policy=MISSING_PRIOR, elbo_form=ELBO_KL, beta=1.0, seed=0, person_offset=0,
               want_grads=True, want_person_outputs=False):
    """vibo_fused_elbo.  Returns dict(scalars (2,) f64 [LL, person_term],
    g_table, g_item (or None), ability_mu/ability_logvar/ability (or None)).
    ``seed`` may be a device int64 tensor {seed, step} (vibo_fused_elbo_graph): the Philox key is
    then read on the device when the kernel runs, so CUDA-graph replays draw fresh noise."""
    _check_rows(response, mask)
    lib = _lib.load()
    P, I = response.shape
    D = table.shape[-1] // 2
    dev = response.device
    desc = make_desc(P, I, D, irt_model, conditional, missing_policy, elbo_form, person_offset)
    F = item_feat_width(irt_model, D)
    assert table.shape == (2, I if conditional else 1, 2 * D) and table.is_contiguous()
    assert item_feat.shape == (I, F) and item_feat.is_contiguous()
    assert table.dtype == torch.float32 and item_feat.dtype == torch.float32
    if eps_ability is not None:
        assert eps_ability.shape == (P, D) and eps_ability.is_contiguous()
        assert eps_ability.dtype == torch.float32
    scalars = torch.empty(2, dtype=torch.float64, device=dev)
    g_table = torch.empty_like(table) if want_grads else None
    g_item = torch.empty_like(item_feat) if want_grads else None
    amu = alv = th = None
    if want_person_outputs:
        amu = torch.empty(P, D, dtype=torch.float32, device=dev)
        alv = torch.empty_like(amu)
        th = torch.empty_like(amu)
    ws = workspace(desc, dev)
    if isinstance(seed, torch.Tensor):
        assert eps_ability is None and seed.is_cuda and seed.dtype == torch.int64 and seed.numel() == 2
        rc = lib.vibo_fused_elbo_graph(C.byref(desc), _ptr(response), _ptr(mask), _ptr(table), _ptr(item_feat),
                                       _ptr(seed), C.c_float(beta), _ptr(scalars), _ptr(amu), _ptr(alv),
                                       _ptr(th), _ptr(g_table), _ptr(g_item), _ptr(ws), ws.numel(),
                                       _stream(dev))
        _lib.check(rc, "vibo_fused_elbo_graph")
    else:
        rc = lib.vibo_fused_elbo(C.byref(desc), _ptr(response), _ptr(mask), _ptr(table), _ptr(item_feat),
                                 _ptr(eps_ability), C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_float(beta),
                                 _ptr(scalars), _ptr(amu), _ptr(alv), _ptr(th), _ptr(g_table), _ptr(g_item),
                                 _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(rc, "vibo_fused_elbo")
    return dict(scalars=scalars, g_table=g_table, g_item=g_item, ability_mu=amu,
                ability_logvar=alv, ability=th)


def fused_elbo_host(response_host, mask_host, table, item_feat, eps_ability, *, irt_model,
                    conditional, missing_policy=MISSING_PRIOR, elbo_form=ELBO_KL, beta=1.0, seed=0,
                    person_offset=0, want_grads=True, chunk_person=65536, staging=None):
    """vibo_fused_elbo_host: response/mask are HOST tensors (pinned for full
    copy bandwidth); returns the same dict as fused_elbo plus ``scalars_host``."""
    lib = _lib.load()
    packed = response_host.dtype == torch.int8   # one byte per cell (-1 missing / 0 / 1), mask_host unused
    assert not response_host.is_cuda and response_host.is_contiguous()
    if not packed:
        assert not mask_host.is_cuda
        assert response_host.dtype == torch.float32
        assert mask_host.dtype == torch.uint8 and mask_host.is_contiguous()
    P, I = response_host.shape
    D = table.shape[-1] // 2
    dev = table.device
    chunk_person = int(min(chunk_person, max(P, 1)))
    desc = make_desc(P, I, D, irt_model, conditional, missing_policy, elbo_form, person_offset)
    need = int(lib.vibo_host_staging_bytes(C.byref(desc), chunk_person))
    if staging is None or staging.numel() < need:
        staging = torch.empty(need, dtype=torch.uint8, device=dev)
    scalars = torch.empty(2, dtype=torch.float64, device=dev)
    scalars_host = torch.empty(2, dtype=torch.float64).pin_memory()
    g_table = torch.empty_like(table) if want_grads else None
    g_item = torch.empty_like(item_feat) if want_grads else None
    ws = workspace(desc, dev)
    if packed:
        rc = lib.vibo_fused_elbo_host_packed(C.byref(desc), _ptr(response_host), _ptr(table), _ptr(item_feat),
                                             _ptr(eps_ability), C.c_uint64(int(seed) & (2 ** 64 - 1)),
                                             C.c_float(beta), _ptr(scalars), _ptr(scalars_host), _ptr(g_table),
                                             _ptr(g_item), C.c_int64(chunk_person), _ptr(staging),
                                             staging.numel(), _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(rc, "vibo_fused_elbo_host_packed")
    else:
        rc = lib.vibo_fused_elbo_host(C.byref(desc), _ptr(response_host), _ptr(mask_host), _ptr(table),
                                      _ptr(item_feat), _ptr(eps_ability),
                                      C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_float(beta), _ptr(scalars),
                                      _ptr(scalars_host), _ptr(g_table), _ptr(g_item),
                                      C.c_int64(chunk_person), _ptr(staging), staging.numel(), _ptr(ws),
                                      ws.numel(), _stream(dev))
        _lib.check(rc, "vibo_fused_elbo_host")
    return dict(scalars=scalars, scalars_host=scalars_host, g_table=g_table, g_item=g_item,
                staging=staging)


def encode(response, mask, table, *, conditional, missing_policy=MISSING_PRIOR):
    """vibo_encode -> (ability_mu, ability_logvar, precision_sum), each (P, D)."""
    _check_rows(response, mask)
    P, I = response.shape
    D = table.shape[-1] // 2
    desc = make_desc(P, I, D, 1, conditional, missing_policy)
    mu = torch.empty(P, D, dtype=torch.float32, device=response.device)
    lv = torch.empty_like(mu)
    S = torch.empty_like(mu)
    rc = _lib.load().vibo_encode(C.byref(desc), _ptr(response), _ptr(mask), _ptr(table.contiguous()),
                                 _ptr(mu), _ptr(lv), _ptr(S), _stream(response.device))
    _lib.check(rc, "vibo_encode")
    return mu, lv, S


def encode_counts(response, mask, table, *, missing_policy=MISSING_PRIOR):
    """vibo_encode_counts (unconditional posterior) -> (ability_mu, ability_logvar, precision_sum, counts) with
    counts (P, 2) = (observed ones, observed cells), or None when the rows are not covered (unaligned views):
    the caller then uses encode / encode_backward."""
    _check_rows(response, mask)
    P, I = response.shape
    D = table.shape[-1] // 2
    desc = make_desc(P, I, D, 1, False, missing_policy)
    mu = torch.empty(P, D, dtype=torch.float32, device=response.device)
    lv = torch.empty_like(mu)
    S = torch.empty_like(mu)
    counts = torch.empty(P, 2, dtype=torch.float32, device=response.device)
    rc = _lib.load().vibo_encode_counts(C.byref(desc), _ptr(response), _ptr(mask), _ptr(table.contiguous()),
                                        _ptr(mu), _ptr(lv), _ptr(S), _ptr(counts), _stream(response.device))
    if rc == ERR_UNSUPPORTED:
        return None
    _lib.check(rc, "vibo_encode_counts")
    return mu, lv, S, counts


def encode_backward_counts(counts, table, ability_mu, precision_sum, g_mu, g_logvar, *, num_item,
                           missing_policy=MISSING_PRIOR):
    """vibo_encode_backward_counts -> g_table (2, 1, 2D), from the counts of encode_counts (no pass over the
    rows)."""
    P, D = ability_mu.shape
    desc = make_desc(P, num_item, D, 1, False, missing_policy)
    g_table = torch.empty_like(table)
    ws = workspace(desc, ability_mu.device)
    rc = _lib.load().vibo_encode_backward_counts(
        C.byref(desc), _ptr(counts), _ptr(table), _ptr(ability_mu), _ptr(precision_sum),
        _ptr(g_mu.contiguous()), _ptr(g_logvar.contiguous()), _ptr(g_table), _ptr(ws), ws.numel(),
        _stream(ability_mu.device))
    _lib.check(rc, "vibo_encode_backward_counts")
    return g_table


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def planar_params_forward(us, ws, bs):
    """vibo_planar_params_forward: the K flows' separate (u (D), w (D), b (1)) parameters -> stacked
    (uhat (K, D), w (K, D), b (K)) with the invertibility correction of flows.py:26-29."""
    Kf, D = len(us), us[0].numel()
    dev = us[0].device
    for t in list(us) + list(ws) + list(bs):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    uhat = torch.empty(Kf, D, dtype=torch.float32, device=dev)
    w_out = torch.empty_like(uhat)
    b_out = torch.empty(Kf, dtype=torch.float32, device=dev)
    pu, pw, pb = _ptr_array(us), _ptr_array(ws), _ptr_array(bs)
    rc = _lib.load().vibo_planar_params_forward(Kf, D, pu, pw, pb, _ptr(uhat), _ptr(w_out), _ptr(b_out),
                                                _stream(dev))
    _lib.check(rc, "vibo_planar_params_forward")
    return uhat, w_out, b_out


def planar_params_backward(us, ws, g_uhat, g_w_out, g_b_out):
    """vibo_planar_params_backward -> (g_u (K, D), g_w (K, D), g_b (K)) for the separate parameters."""
    Kf, D = len(us), us[0].numel()
    dev = us[0].device
    g_u = torch.empty(Kf, D, dtype=torch.float32, device=dev)
    g_w = torch.empty_like(g_u)
    g_b = torch.empty(Kf, dtype=torch.float32, device=dev)
    pu, pw = _ptr_array(us), _ptr_array(ws)
    rc = _lib.load().vibo_planar_params_backward(Kf, D, pu, pw, _ptr(g_uhat.contiguous()),
                                                 _ptr(g_w_out.contiguous()), _ptr(g_b_out.contiguous()),
                                                 _ptr(g_u), _ptr(g_w), _ptr(g_b), _stream(dev))
    _lib.check(rc, "vibo_planar_params_backward")
    return g_u, g_w, g_b


def encode_backward(response, mask, table, ability_mu, precision_sum, g_mu, g_logvar, *,
                    conditional, missing_policy=MISSING_PRIOR):
    """vibo_encode_backward -> g_table (2, It, 2D)."""
    _check_rows(response, mask)
    P, I = response.shape
    D = table.shape[-1] // 2
    desc = make_desc(P, I, D, 1, conditional, missing_policy)
    g_table = torch.empty_like(table)
    ws = workspace(desc, response.device)
    rc = _lib.load().vibo_encode_backward(
        C.byref(desc), _ptr(response), _ptr(mask), _ptr(table), _ptr(ability_mu), _ptr(precision_sum),
        _ptr(g_mu.contiguous()), _ptr(g_logvar.contiguous()), _ptr(g_table), _ptr(ws), ws.numel(),
        _stream(response.device))
    _lib.check(rc, "vibo_encode_backward")
    return g_table


def link_loglik(response, mask, ability, item_feat, *, irt_model, want_grads=True):
    """vibo_link_loglik -> (ll (1,) f64, dLL/d ability or None, dLL/d item_feat or None)."""
    _check_rows(response, mask)
    P, I = response.shape
    D = ability.shape[1]
    desc = make_desc(P, I, D, irt_model, 0)
    ll = torch.empty(1, dtype=torch.float64, device=response.device)
    g_ab = torch.empty_like(ability) if want_grads else None
    g_it = torch.empty_like(item_feat) if want_grads else None
    ws = workspace(desc, response.device)
    rc = _lib.load().vibo_link_loglik(C.byref(desc), _ptr(response), _ptr(mask), _ptr(ability),
                                      _ptr(item_feat), _ptr(ll), _ptr(g_ab), _ptr(g_it), _ptr(ws),
                                      ws.numel(), _stream(response.device))
    _lib.check(rc, "vibo_link_loglik")
    return ll, g_ab, g_it


def decode(ability, item_feat, *, irt_model):
    """vibo_decode -> response_mu (P, I)."""
    if not ability.is_cuda:
        raise _lib.ViboError("VIBO kernels need CUDA tensors (no CPU fallback exists)")
    P, D = ability.shape
    I = item_feat.shape[0]
    desc = make_desc(P, I, D, irt_model, 0)
    out = torch.empty(P, I, dtype=torch.float32, device=ability.device)
    rc = _lib.load().vibo_decode(C.byref(desc), _ptr(ability), _ptr(item_feat), _ptr(out),
                                 _stream(ability.device))
    _lib.check(rc, "vibo_decode")
    return out


def bernoulli_loglik(response, mask, response_mu, *, want_grad=True):
    """vibo_bernoulli_loglik -> (ll (1,) f64, dLL/d response_mu (P, I) or None)."""
    _check_rows(response, mask)
    P, I = response.shape
    desc = make_desc(P, I, 1, 1, 0)
    ll = torch.empty(1, dtype=torch.float64, device=response.device)
    g = torch.empty_like(response_mu) if want_grad else None
    ws = workspace(desc, response.device)
    rc = _lib.load().vibo_bernoulli_loglik(C.byref(desc), _ptr(response), _ptr(mask), _ptr(response_mu),
                                           _ptr(ll), _ptr(g), _ptr(ws), ws.numel(),
                                           _stream(response.device))
    _lib.check(rc, "vibo_bernoulli_loglik")
    return ll, g


def param_forward(mu, lv, eps_item, w0, b0, w2, b2, w4, b4, *, irt_model, elbo_form=ELBO_KL):
    """vibo_param_forward -> (item_feat (I,F), table (2,1,2D), hidden (2,2,H), item_term (1,) f64)."""
    if not mu.is_cuda:
        raise _lib.ViboError("VIBO kernels need CUDA tensors (no CPU fallback exists)")
    I, F = mu.shape
    H = w2.shape[0]
    D = w4.shape[0] // 2
    desc = make_desc(0, I, D, irt_model, 0, MISSING_PRIOR, elbo_form)
    item_feat = torch.empty_like(mu)
    table = torch.empty(2, 1, 2 * D, dtype=torch.float32, device=mu.device)
    hidden = torch.empty(2, 2, H, dtype=torch.float32, device=mu.device)
    term = torch.empty(1, dtype=torch.float64, device=mu.device)
    rc = _lib.load().vibo_param_forward(C.byref(desc), H, _ptr(mu), _ptr(lv), _ptr(eps_item), _ptr(w0), _ptr(b0),
                                        _ptr(w2), _ptr(b2), _ptr(w4), _ptr(b4), _ptr(item_feat), _ptr(table),
                                        _ptr(hidden), _ptr(term), _stream(mu.device))
    _lib.check(rc, "vibo_param_forward")
    return item_feat, table, hidden, term


def param_backward(mu, lv, eps_item, w2, w4, hidden, g_table, g_item, g_term, *, irt_model, elbo_form=ELBO_KL):
    """vibo_param_backward -> gradients (g_mu, g_lv, g_w0, g_b0, g_w2, g_b2, g_w4, g_b4)."""
    I, F = mu.shape
    H = w2.shape[0]
    D = w4.shape[0] // 2
    desc = make_desc(0, I, D, irt_model, 0, MISSING_PRIOR, elbo_form)
    dev = mu.device
    outs = [torch.empty_like(mu), torch.empty_like(lv), torch.empty(H, 1, device=dev), torch.empty(H, device=dev),
            torch.empty(H, H, device=dev), torch.empty(H, device=dev), torch.empty(2 * D, H, device=dev),
            torch.empty(2 * D, device=dev)]
    rc = _lib.load().vibo_param_backward(C.byref(desc), H, _ptr(mu), _ptr(lv), _ptr(eps_item), _ptr(w2), _ptr(w4),
                                         _ptr(hidden), _ptr(g_table), _ptr(g_item), _ptr(g_term),
                                         *[_ptr(o) for o in outs], _stream(dev))
    _lib.check(rc, "vibo_param_backward")
    return outs


def flow_person_forward(ability_mu, ability_logvar, eps, uhat, w, b):
    """vibo_flow_person_forward -> (ability_0, ability_k (P, D), term f64 0-d)."""
    lib = _lib.load()
    if not ability_mu.is_cuda:
        raise _lib.ViboError("VIBO kernels need CUDA tensors (no CPU fallback exists)")
    P, D = ability_mu.shape
    K = uhat.shape[0]
    dev = ability_mu.device
    desc = make_desc(P, 1, D, 2, False)
    th0 = torch.empty_like(ability_mu)
    thk = torch.empty_like(ability_mu)
    term = torch.empty(1, dtype=torch.float64, device=dev)
    ws = workspace(desc, dev)
    rc = lib.vibo_flow_person_forward(C.byref(desc), C.c_int(K), _ptr(ability_mu.contiguous()),
                                      _ptr(ability_logvar.contiguous()), _ptr(eps.contiguous()),
                                      _ptr(uhat.contiguous()), _ptr(w.contiguous()), _ptr(b.contiguous()),
                                      _ptr(th0), _ptr(thk), _ptr(term), _ptr(ws), ws.numel(), _stream(dev))
    _lib.check(rc, "vibo_flow_person_forward")
    return th0, thk, term


def flow_person_backward(ability_mu, ability_logvar, eps, uhat, w, b, g_ability_k, g_term):
    """vibo_flow_person_backward -> (g_mu, g_logvar, g_uhat, g_w, g_b)."""
    lib = _lib.load()
    P, D = ability_mu.shape
    K = uhat.shape[0]
    dev = ability_mu.device
    desc = make_desc(P, 1, D, 2, False)
    g_mu = torch.empty_like(ability_mu)
    g_lv = torch.empty_like(ability_mu)
    g_uhat = torch.empty(K, D, dtype=torch.float32, device=dev)
    g_w = torch.empty(K, D, dtype=torch.float32, device=dev)
    g_b = torch.empty(K, dtype=torch.float32, device=dev)
    ws = workspace(desc, dev)
    rc = lib.vibo_flow_person_backward(C.byref(desc), C.c_int(K), _ptr(ability_mu.contiguous()),
                                       _ptr(ability_logvar.contiguous()), _ptr(eps.contiguous()),
                                       _ptr(uhat.contiguous()), _ptr(w.contiguous()), _ptr(b.contiguous()),
                                       _ptr(g_ability_k.contiguous()), _ptr(g_term.contiguous()), _ptr(g_mu),
                                       _ptr(g_lv), _ptr(g_uhat), _ptr(g_w), _ptr(g_b), _ptr(ws), ws.numel(),
                                       _stream(dev))
    _lib.check(rc, "vibo_flow_person_backward")
    return g_mu, g_lv, g_uhat, g_w, g_b


def person_counts(response, mask):
    """vibo_person_counts -> (P, 2) float32: (observed ones, observed cells) per person."""
    _check_rows(response, mask)
    P, I = response.shape
    desc = make_desc(P, I, 1, 2, False)
    counts = torch.empty(P, 2, dtype=torch.float32, device=response.device)
    rc = _lib.load().vibo_person_counts(C.byref(desc), _ptr(response), _ptr(mask), _ptr(counts),
                                        _stream(response.device))
    _lib.check(rc, "vibo_person_counts")
    return counts


def philox_normal(P, D, seed, person_offset, device):
    """vibo_philox_normal -> (P, D) float32: the noise the fused kernels draw in-kernel for
    (seed, person_offset).  ``seed`` is an int or a device int64 {seed, step} tensor."""
    if torch.device(device).type != "cuda":
        raise _lib.ViboError("VIBO kernels need CUDA tensors (no CPU fallback exists)")
    desc = make_desc(P, 1, D, 2, False, person_offset=person_offset)
    out = torch.empty(P, D, dtype=torch.float32, device=device)
    if isinstance(seed, torch.Tensor):
        rc = _lib.load().vibo_philox_normal(C.byref(desc), C.c_uint64(0), _ptr(seed), _ptr(out), _stream(device))
    else:
        rc = _lib.load().vibo_philox_normal(C.byref(desc), C.c_uint64(int(seed) & (2 ** 64 - 1)), None, _ptr(out),
                                            _stream(device))
    _lib.check(rc, "vibo_philox_normal")
    return out


def log_marginal(response, mask, table, item_mu, item_logvar, num_samples, *, irt_model,
                 missing_policy=MISSING_PRIOR, eps_item=None, eps_ability=None, seed=0, person_offset=0):
    """vibo_log_marginal -> (logp 0-d f64, log_weights (S,) f64): IWAE bound with the sample loop in
    the kernel (unconditional posterior)."""
    _check_rows(response, mask)
    lib = _lib.load()
    P, I = response.shape
    D = table.shape[-1] // 2
    dev = response.device
    desc = make_desc(P, I, D, irt_model, False, missing_policy, ELBO_SAMPLE, person_offset)
    S = int(num_samples)
    if eps_item is not None:
        assert eps_item.shape == (S, I, item_feat_width(irt_model, D)) and eps_item.is_contiguous()
    if eps_ability is not None:
        assert eps_ability.shape == (S, P, D) and eps_ability.is_contiguous()
    need = int(lib.vibo_log_marginal_workspace_bytes(S))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    log_w = torch.empty(S, dtype=torch.float64, device=dev)
    logp = torch.empty(1, dtype=torch.float64, device=dev)
    seed_t = seed if isinstance(seed, torch.Tensor) else None
    rc = lib.vibo_log_marginal(C.byref(desc), _ptr(response), _ptr(mask), _ptr(table.contiguous()),
                               _ptr(item_mu.contiguous()), _ptr(item_logvar.contiguous()), S, _ptr(eps_item),
                               _ptr(eps_ability), C.c_uint64(0 if seed_t is not None else int(seed) & (2 ** 64 - 1)),
                               _ptr(seed_t), _ptr(log_w), _ptr(logp), _ptr(ws), ws.numel(), _stream(dev))
    _lib.check(rc, "vibo_log_marginal")
    return logp[0], log_w


def predictive_mean(ability_mu, ability_logvar, item_mu, item_logvar, num_samples, *, irt_model, seed=0,
                    person_offset=0):
    """vibo_predictive_mean -> (P, I) float32: mean over S posterior draws of the decoded response."""
    if not ability_mu.is_cuda:
        raise _lib.ViboError("VIBO kernels need CUDA tensors (no CPU fallback exists)")
    P, D = ability_mu.shape
    I = item_mu.shape[0]
    dev = ability_mu.device
    desc = make_desc(P, I, D, irt_model, False, person_offset=person_offset)
    out = torch.empty(P, I, dtype=torch.float32, device=dev)
    seed_t = seed if isinstance(seed, torch.Tensor) else None
    rc = _lib.load().vibo_predictive_mean(C.byref(desc), _ptr(ability_mu.contiguous()),
                                          _ptr(ability_logvar.contiguous()), _ptr(item_mu.contiguous()),
                                          _ptr(item_logvar.contiguous()), int(num_samples),
                                          C.c_uint64(0 if seed_t is not None else int(seed) & (2 ** 64 - 1)),
                                          _ptr(seed_t), _ptr(out), _stream(dev))
    _lib.check(rc, "vibo_predictive_mean")
    return out


def pack_rows(response, mask):
    """vibo_pack: (P, I) float32 response + uint8 mask on the GPU -> (P, I) int8 (-1 missing / 0 / 1)."""
    _check_rows(response, mask)
    P, I = response.shape
    desc = make_desc(P, I, 1, 1, False)
    out = torch.empty(P, I, dtype=torch.int8, device=response.device)
    rc = _lib.load().vibo_pack(C.byref(desc), _ptr(response), _ptr(mask), _ptr(out), _stream(response.device))
    _lib.check(rc, "vibo_pack")
    return out


def unpack_rows(packed):
    """vibo_unpack: (P, I) int8 on the GPU -> ((P, I) float32 response, (P, I) uint8 mask)."""
    if not packed.is_cuda:
        raise _lib.ViboError("VIBO kernels need CUDA tensors (no CPU fallback exists)")
    assert packed.dtype == torch.int8 and packed.dim() == 2 and packed.is_contiguous()
    P, I = packed.shape
    desc = make_desc(P, I, 1, 1, False)
    resp = torch.empty(P, I, dtype=torch.float32, device=packed.device)
    mask = torch.empty(P, I, dtype=torch.uint8, device=packed.device)
    rc = _lib.load().vibo_unpack(C.byref(desc), _ptr(packed), _ptr(resp), _ptr(mask), _stream(packed.device))
    _lib.check(rc, "vibo_unpack")
    return resp, mask


def pack_rows_host(response, mask):
    """Host-side packing of dataset arrays (done once at load): (P, I[, 1]) response + mask ->
    pinned (P, I) int8 in the packed row format."""
    r = response.reshape(response.shape[0], response.shape[1]).contiguous().float()
    m = mask.reshape(mask.shape[0], mask.shape[1])
    m = (m if m.dtype == torch.uint8 else (m != 0).to(torch.uint8)).contiguous()
    out = torch.empty(r.shape, dtype=torch.int8)
    if torch.cuda.is_available():
        out = out.pin_memory()
    desc = make_desc(r.shape[0], r.shape[1], 1, 2, False)
    # the library's host thread pool (vibo_pack_host): ~100 GB/s of rows on 16 cores
    _lib.check(_lib.load().vibo_pack_host(C.byref(desc), _ptr(r), _ptr(m), _ptr(out)), "vibo_pack_host")
    return out


def percell_mlp(u, v, z, w0, w2, c2, w4, c4=0.0):
    """vibo_percell_mlp -> (P, I) float32: w4 . ELU(W2 ELU(u_j + v_i + z_ij w0) + c2) + c4 on the tcgen05
    tensor cores.  u (I or 1, 64), v (P or 1, 64), z (P, I) with w0 (64) or both None; P and I are taken
    from z, else from v and u."""
    if not w2.is_cuda:
        raise _lib.ViboError("VIBO kernels need CUDA tensors (no CPU fallback exists)")
    H = w2.shape[0]
    u, v = u.contiguous().float(), v.contiguous().float()
    if z is not None:
        z = z.contiguous().float()
        P, I = z.shape
    else:
        P, I = v.shape[0], u.shape[0]
    out = torch.empty(P, I, dtype=torch.float32, device=w2.device)
    rc = _lib.load().vibo_percell_mlp(P, I, H, u.shape[0], v.shape[0], _ptr(u), _ptr(v), _ptr(z),
                                      _ptr(None if w0 is None else w0.contiguous().float()),
                                      _ptr(w2.contiguous().float()), _ptr(c2.contiguous().float()),
                                      _ptr(w4.contiguous().float()), C.c_float(float(c4)), _ptr(out),
                                      _stream(w2.device))
    _lib.check(rc, "vibo_percell_mlp")
    return out
