"""numpy-oracle implementation of the ``kernels`` module interface.

TEST-ONLY.  ``install(monkeypatch)`` swaps the tensor-level kernel wrappers of
the product package for oracle/kernel_spec.py so that the HOST logic (drop-in
modules, autograd glue, flows, sharding) can be exercised on a machine without
a GPU.  The product itself never imports this.
"""
import numpy as np
import torch

from oracle import kernel_spec as KS


def _np(t, dt=np.float64):
    return None if t is None else t.detach().cpu().numpy().astype(dt)


def _t(a, like, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(like.device)


def philox_normals(seed, person_ids, D):
    """Host restatement of philox_normal4 (csrc/vibo_common.cuh)."""
    out = np.zeros((len(person_ids), D), dtype=np.float32)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask32 = np.uint64(0xFFFFFFFF)
    for row, pid in enumerate(person_ids):
        for blk in range((D + 3) // 4):
            c = [np.uint64(pid & 0xFFFFFFFF), np.uint64((pid >> 32) & 0xFFFFFFFF), np.uint64(blk), np.uint64(0)]
            k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
            for _ in range(10):
                p0 = M0 * c[0]
                p1 = M1 * c[2]
                c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask32, p1 & mask32,
                     ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask32, p0 & mask32]
                k0 = (k0 + np.uint64(0x9E3779B9)) & mask32
                k1 = (k1 + np.uint64(0xBB67AE85)) & mask32
            vals = []
            for h in range(2):
                u1 = (np.float32(int(c[2 * h])) + np.float32(1.0)) * np.float32(2.3283064365386963e-10)
                u2 = np.float32(int(c[2 * h + 1])) * np.float32(2.3283064365386963e-10)
                rad = np.sqrt(np.float32(-2.0) * np.log(u1))
                vals += [rad * np.cos(np.pi * 2.0 * u2), rad * np.sin(np.pi * 2.0 * u2)]
            for k in range(4):
                d = blk * 4 + k
                if d < D:
                    out[row, d] = vals[k]
    return out


def _seed_int(seed):
    if isinstance(seed, torch.Tensor):
        return int(seed[0].item()) + int(seed[1].item())
    return int(seed)


def philox_normal(P, D, seed, person_offset, device):
    out = philox_normals(_seed_int(seed), [person_offset + i for i in range(P)], D)
    return torch.from_numpy(out).to(device)


def fused_elbo(response, mask, table, item_feat, eps_ability, *, irt_model, conditional,
               missing_policy=0, elbo_form=0, beta=1.0, seed=0, person_offset=0, want_grads=True,
               want_person_outputs=False):
    P = response.shape[0]
    D = table.shape[-1] // 2
    if eps_ability is None:
        eps = philox_normals(_seed_int(seed), [person_offset + i for i in range(P)], D).astype(np.float64)
    else:
        eps = _np(eps_ability)
    r = KS.fused_elbo(_np(response), _np(mask, np.uint8), _np(table), _np(item_feat), eps,
                      irt_model=irt_model, beta=beta, missing_policy=missing_policy,
                      elbo_form=elbo_form, want_grads=want_grads)
    out = dict(scalars=torch.tensor([r["ll"], r["person_term"]], dtype=torch.float64, device=response.device),
               g_table=None, g_item=None, ability_mu=None, ability_logvar=None, ability=None)
    if want_grads:
        out["g_table"] = _t(r["g_table"], response)
        out["g_item"] = _t(r["g_item"], response)
    if want_person_outputs:
        for k in ("ability_mu", "ability_logvar", "ability"):
            out[k] = _t(r[k], response)
    return out


def encode(response, mask, table, *, conditional, missing_policy=0):
    D = table.shape[-1] // 2
    r = KS.encode(_np(response), _np(mask, np.uint8), _np(table), D, missing_policy)
    return _t(r["ability_mu"], response), _t(r["ability_logvar"], response), _t(r["S"], response)


def encode_backward(response, mask, table, ability_mu, precision_sum, g_mu, g_logvar, *,
                    conditional, missing_policy=0):
    D = table.shape[-1] // 2
    g = KS.encode_backward(_np(response), _np(mask, np.uint8), _np(table), D, _np(precision_sum),
                           _np(ability_mu), _np(g_mu), _np(g_logvar))
    return _t(g, response)


def encode_counts(response, mask, table, *, missing_policy=0):
    mu, lv, S = encode(response, mask, table, conditional=False, missing_policy=missing_policy)
    n0, n1, _ = KS.person_counts(_np(response), _np(mask, np.uint8))
    return mu, lv, S, _t(np.stack([n1, n0 + n1], 1), response)   # (observed ones, observed cells)


def encode_backward_counts(counts, table, ability_mu, precision_sum, g_mu, g_logvar, *, num_item,
                           missing_policy=0):
    D = table.shape[-1] // 2
    g = KS.encode_backward_counts(_np(counts), _np(table), D, _np(precision_sum), _np(ability_mu), _np(g_mu),
                                  _np(g_logvar))
    return _t(g, table)


def planar_params_forward(us, ws, bs):
    u, w = np.stack([_np(t) for t in us]), np.stack([_np(t) for t in ws])
    uhat = KS.planar_params(u, w)
    like = us[0]
    return _t(uhat, like), _t(w, like), _t(np.concatenate([_np(t) for t in bs]), like)


def planar_params_backward(us, ws, g_uhat, g_w_out, g_b_out):
    u, w = np.stack([_np(t) for t in us]), np.stack([_np(t) for t in ws])
    _, g_u, g_w = KS.planar_params(u, w, _np(g_uhat), _np(g_w_out))
    like = us[0]
    return _t(g_u, like), _t(g_w, like), g_b_out.clone()


def link_loglik(response, mask, ability, item_feat, *, irt_model, want_grads=True):
    r = KS.link_loglik(_np(response), _np(mask, np.uint8), _np(ability), _np(item_feat), irt_model,
                       want_grads)
    ll = torch.tensor([r["ll"]], dtype=torch.float64, device=response.device)
    if not want_grads:
        return ll, None, None
    return ll, _t(r["g_ability"], response), _t(r["g_item"], response)


def decode(ability, item_feat, *, irt_model):
    return _t(KS.decode(_np(ability), _np(item_feat), irt_model), ability)


def bernoulli_loglik(response, mask, response_mu, *, want_grad=True):
    ll, g = KS.bernoulli_loglik(_np(response), _np(mask, np.uint8), _np(response_mu))
    out = torch.tensor([ll.sum()], dtype=torch.float64, device=response.device)
    return out, (_t(g, response) if want_grad else None)


def _check_rows(response, mask):
    assert response.dtype == torch.float32 and response.dim() == 2
    assert mask.dtype == torch.uint8 and mask.shape == response.shape


def install(monkeypatch):
    import vibo_b200
    K = vibo_b200.kernels
    for name in ("fused_elbo", "encode", "encode_backward", "encode_counts", "encode_backward_counts",
                 "planar_params_forward", "planar_params_backward", "link_loglik", "decode",
                 "bernoulli_loglik", "_check_rows", "philox_normal"):
        monkeypatch.setattr(K, name, globals()[name])
