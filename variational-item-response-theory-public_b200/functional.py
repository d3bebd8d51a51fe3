"""Autograd bindings of the VIBO kernels.

Each Function wraps one C-ABI entry point (through ``kernels``) so that the
small parameter-side chains -- encoder MLP on the 2 / 2*I expert rows, item
reparameterisation, planar flows, item KL -- stay ordinary PyTorch autograd on
tiny tensors, while everything that touches the (P, I) response matrix runs in
the CUDA kernels.
"""
from __future__ import annotations

import torch

from . import kernels as K

ELBO_KL, ELBO_SAMPLE = K.ELBO_KL, K.ELBO_SAMPLE
MISSING_PRIOR, MISSING_DROP = K.MISSING_PRIOR, K.MISSING_DROP


def prepare_rows(response: torch.Tensor, mask: torch.Tensor, binary: bool = True):
    """(P, I, 1) float response + any-dtype mask -> (P, I) float32 and (P, I)
    uint8 views the kernels read.  bool/uint8 masks are reinterpreted without a
    copy; the int64 mask the reference CLI builds (vibo.py:240) costs one
    conversion pass."""
    if response.dim() == 3:
        response = response.view(response.shape[0], response.shape[1]) if response.is_contiguous() \
            else response.reshape(response.shape[0], response.shape[1])
    if response.dtype == torch.int8:
        # packed rows (-1 missing / 0 / 1, one byte per cell): host tensors stay packed (the host
        # entry transfers them as they are), device tensors are expanded by vibo_unpack
        response = response.contiguous()
        if not response.is_cuda:
            return response, None
        return K.unpack_rows(response)
    if mask.dim() == 3:
        mask = mask.reshape(mask.shape[0], mask.shape[1])
    if response.dtype != torch.float32:
        response = response.float()
    response = response.contiguous()
    if mask.dtype == torch.bool:
        mask = mask.contiguous().view(torch.uint8)
    elif mask.dtype != torch.uint8:
        mask = (mask != 0).to(torch.uint8)
    return response, mask.contiguous()


class FusedElbo(torch.autograd.Function):
    """loss_k = -LL + beta*KL_theta (KL form) or -LL - sum(log p - log q)(theta)
    (sample form), one pass over the rows (vibo_fused_elbo)."""

    @staticmethod
    def forward(ctx, response, mask, table, item_feat, eps_ability, cfg):
        want = bool(ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        out = K.fused_elbo(response, mask, table.detach().contiguous(),
                           item_feat.detach().contiguous(), eps_ability,
                           irt_model=cfg["irt_model"], conditional=cfg["conditional"],
                           missing_policy=cfg["missing_policy"], elbo_form=cfg["elbo_form"],
                           beta=cfg["beta"], seed=cfg.get("seed", 0),
                           person_offset=cfg.get("person_offset", 0), want_grads=want,
                           want_person_outputs=cfg.get("want_person_outputs", False))
        ll, term = out["scalars"][0], out["scalars"][1]
        if cfg["elbo_form"] == ELBO_KL:
            loss_k = -ll + cfg["beta"] * term
        else:
            loss_k = -ll - term
        if want:
            ctx.save_for_backward(out["g_table"], out["g_item"])
        extras = [out["scalars"]]
        for name in ("ability_mu", "ability_logvar", "ability"):
            extras.append(out[name] if out[name] is not None else torch.empty(0, device=response.device))
        ctx.mark_non_differentiable(*extras)
        return (loss_k.to(torch.float32), *extras)

    @staticmethod
    def backward(ctx, g_loss, *_):
        g_table, g_item = ctx.saved_tensors
        return None, None, g_loss * g_table, g_loss * g_item, None, None


class FusedElboHost(torch.autograd.Function):
    """FusedElbo with the response / mask rows in HOST memory
    (vibo_fused_elbo_host): person chunks are streamed host->device inside the
    call, overlapped with the kernels; the loss comes back in pinned memory."""

    @staticmethod
    def forward(ctx, response_host, mask_host, table, item_feat, eps_ability, cfg):
        want = bool(ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        out = K.fused_elbo_host(response_host, mask_host, table.detach().contiguous(),
                                item_feat.detach().contiguous(), eps_ability,
                                irt_model=cfg["irt_model"], conditional=cfg["conditional"],
                                missing_policy=cfg["missing_policy"], elbo_form=cfg["elbo_form"],
                                beta=cfg["beta"], seed=cfg.get("seed", 0),
                                person_offset=cfg.get("person_offset", 0), want_grads=want,
                                chunk_person=cfg.get("chunk_person", 65536),
                                staging=cfg.get("staging"))
        cfg["staging"] = out["staging"]  # reused by the next call
        ll, term = out["scalars"][0], out["scalars"][1]
        loss_k = (-ll + cfg["beta"] * term) if cfg["elbo_form"] == ELBO_KL else (-ll - term)
        if want:
            ctx.save_for_backward(out["g_table"], out["g_item"])
        ctx.mark_non_differentiable(out["scalars"], out["scalars_host"])
        return loss_k.to(torch.float32), out["scalars"], out["scalars_host"]

    @staticmethod
    def backward(ctx, g_loss, *_):
        g_table, g_item = ctx.saved_tensors
        return None, None, g_loss * g_table, g_loss * g_item, None, None


class ParamChain(torch.autograd.Function):
    """Parameter-side chain of the unconditional model in two kernels
    (vibo_param_forward / vibo_param_backward): item reparameterisation,
    expert table = encoder MLP on the cell inputs {0, 1}, and the item-side
    prior term.  Replaces ~80 tiny PyTorch kernels per training step."""

    @staticmethod
    def forward(ctx, mu, lv, w0, b0, w2, b2, w4, b4, eps_item, irt_model, elbo_form):
        args = [t.detach().contiguous() for t in (mu, lv, eps_item, w0, b0, w2, b2, w4, b4)]
        item_feat, table, hidden, term = K.param_forward(*args, irt_model=irt_model, elbo_form=elbo_form)
        ctx.save_for_backward(args[0], args[1], args[2], args[5], args[7], hidden)
        ctx.cfg = (irt_model, elbo_form)
        return item_feat, table, term[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g_item_feat, g_table, g_term):
        mu, lv, eps_item, w2, w4, hidden = ctx.saved_tensors
        irt_model, elbo_form = ctx.cfg
        if g_item_feat is None:
            g_item_feat = torch.zeros_like(mu)
        if g_table is None:
            g_table = torch.zeros(2, 1, w4.shape[0], device=mu.device)
        if g_term is None:
            g_term = torch.zeros((), device=mu.device)
        grads = K.param_backward(mu, lv, eps_item, w2, w4, hidden, g_table.contiguous().float(),
                                 g_item_feat.contiguous().float(), g_term.reshape(1).float().contiguous(),
                                 irt_model=irt_model, elbo_form=elbo_form)
        g_mu, g_lv, g_w0, g_b0, g_w2, g_b2, g_w4, g_b4 = grads
        return g_mu, g_lv, g_w0, g_b0, g_w2, g_b2, g_w4, g_b4, None, None, None


class EncodePosterior(torch.autograd.Function):
    """(ability_mu, ability_logvar) = PoE over each person's experts
    (vibo_encode / vibo_encode_backward)."""

    @staticmethod
    def forward(ctx, response, mask, table, conditional, missing_policy):
        tbl = table.detach().contiguous()
        counts = None
        if not conditional and table.requires_grad:
            # unconditional table: the forward pass also leaves the per-person counts, and the backward works
            # from them alone (one pass over the rows instead of two)
            out = K.encode_counts(response, mask, tbl, missing_policy=missing_policy)
            if out is not None:
                mu, lv, S, counts = out
        if counts is None:
            mu, lv, S = K.encode(response, mask, tbl, conditional=conditional, missing_policy=missing_policy)
            ctx.save_for_backward(response, mask, tbl, mu, S)
        else:
            ctx.save_for_backward(counts, tbl, mu, S)
        ctx.cfg = (conditional, missing_policy, counts is not None, response.shape[1])
        return mu, lv

    @staticmethod
    def backward(ctx, g_mu, g_lv):
        conditional, missing_policy, by_counts, num_item = ctx.cfg
        if by_counts:
            counts, tbl, mu, S = ctx.saved_tensors
            g_table = K.encode_backward_counts(counts, tbl, mu, S, g_mu.contiguous(), g_lv.contiguous(),
                                               num_item=num_item, missing_policy=missing_policy)
            return None, None, g_table, None, None
        response, mask, tbl, mu, S = ctx.saved_tensors
        g_table = K.encode_backward(response, mask, tbl, mu, S, g_mu.contiguous(), g_lv.contiguous(),
                                    conditional=conditional, missing_policy=missing_policy)
        return None, None, g_table, None, None


class LinkLogLik(torch.autograd.Function):
    """LL = sum_ij o_ij log Bernoulli(x_ij; irt(ability, item_feat)) without
    materialising response_mu (vibo_link_loglik).  Gradients are produced in
    the same pass as the value."""

    @staticmethod
    def forward(ctx, response, mask, ability, item_feat, irt_model):
        want = bool(ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        ll, g_ab, g_it = K.link_loglik(response, mask, ability.detach().contiguous(),
                                       item_feat.detach().contiguous(), irt_model=irt_model,
                                       want_grads=want)
        if want:
            ctx.save_for_backward(g_ab, g_it)
        return ll[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        g_ab, g_it = ctx.saved_tensors
        return None, None, g * g_ab, g * g_it, None


class PlanarParams(torch.autograd.Function):
    """The K planar flows' separate parameters (u_k, w_k, b_k) -> stacked (uhat (K, D), w (K, D), b (K)) with the
    invertibility correction of reference flows.py:26-29, one kernel each way
    (vibo_planar_params_forward / _backward)."""

    @staticmethod
    def forward(ctx, n_flows, *params):
        us, ws, bs = params[:n_flows], params[n_flows:2 * n_flows], params[2 * n_flows:]
        us = [t.detach() for t in us]
        ws = [t.detach() for t in ws]
        uhat, w_out, b_out = K.planar_params_forward(us, ws, [t.detach() for t in bs])
        ctx.save_for_backward(*us, *ws)
        ctx.n_flows = n_flows
        return uhat, w_out, b_out

    @staticmethod
    def backward(ctx, g_uhat, g_w, g_b):
        n = ctx.n_flows
        saved = ctx.saved_tensors
        us, ws = saved[:n], saved[n:]
        D = us[0].numel()
        dev = us[0].device
        if g_uhat is None:
            g_uhat = torch.zeros(n, D, device=dev)
        if g_w is None:
            g_w = torch.zeros(n, D, device=dev)
        if g_b is None:
            g_b = torch.zeros(n, device=dev)
        gu, gw, gb = K.planar_params_backward(us, ws, g_uhat, g_w, g_b)
        return (None, *[gu[k] for k in range(n)], *[gw[k] for k in range(n)], *[gb[k:k + 1] for k in range(n)])


class FlowPerson(torch.autograd.Function):
    """Reparameterised ability draw + K planar flows + the person-side terms of the
    flow-form ELBO, one kernel each way (vibo_flow_person_forward / _backward).
    Returns (ability_0, ability_K, term) with
    term = sum_i [log N(theta_K; 0, 1) - log N(theta_0; mu, exp lv) + sum_k ldj_k]
    (reference models.py:412-424, flows.py:21-41)."""

    @staticmethod
    def forward(ctx, ability_mu, ability_logvar, eps, uhat, w, b):
        mu, lv, e = ability_mu.detach().contiguous(), ability_logvar.detach().contiguous(), eps.contiguous()
        uh, ww, bb = uhat.detach().contiguous(), w.detach().contiguous(), b.detach().contiguous()
        th0, thk, term = K.flow_person_forward(mu, lv, e, uh, ww, bb)
        ctx.save_for_backward(mu, lv, e, uh, ww, bb)
        ctx.mark_non_differentiable(th0)
        return th0, thk, term[0].to(torch.float32)

    @staticmethod
    def backward(ctx, _g_th0, g_thk, g_term):
        mu, lv, e, uh, ww, bb = ctx.saved_tensors
        if g_thk is None:
            g_thk = torch.zeros_like(mu)
        if g_term is None:
            g_term = torch.zeros((), device=mu.device)
        g_mu, g_lv, g_uh, g_w, g_b = K.flow_person_backward(
            mu, lv, e, uh, ww, bb, g_thk.contiguous().float(), g_term.reshape(1).float().contiguous())
        return g_mu, g_lv, None, g_uh, g_w, g_b


class Decode(torch.autograd.Function):
    """response_mu (P, I, 1) = irt_model_{1,2,3}pl(ability, item_feat)
    (vibo_decode).  API-parity path: the backward re-derives the link with
    dense torch ops since a materialised (P, I) gradient arrives anyway."""

    @staticmethod
    def forward(ctx, ability, item_feat, irt_model):
        out = K.decode(ability.detach().contiguous(), item_feat.detach().contiguous(), irt_model=irt_model)
        ctx.save_for_backward(ability, item_feat)
        ctx.irt_model = irt_model
        return out.unsqueeze(2)

    @staticmethod
    def backward(ctx, g):
        ability, item_feat = ctx.saved_tensors
        m = ctx.irt_model
        D = ability.shape[1]
        g = g.reshape(g.shape[0], g.shape[1])
        if m == 1:
            z = ability.sum(1, keepdim=True) + item_feat[:, 0][None, :]
        else:
            z = ability @ (-item_feat[:, :D].T) + item_feat[:, D][None, :]
        s = torch.sigmoid(z)
        g_item = torch.zeros_like(item_feat)
        if m == 3:
            guess = torch.sigmoid(item_feat[:, D + 1])
            gz = g * (1 - guess)[None, :] * s * (1 - s)
            g_item[:, D + 1] = (g * (1 - s)).sum(0) * guess * (1 - guess)
        else:
            gz = g * s * (1 - s)
        if m == 1:
            g_ability = gz.sum(1, keepdim=True).expand(-1, D).contiguous()
            g_item[:, 0] = gz.sum(0)
        else:
            g_ability = -(gz @ item_feat[:, :D])
            g_item[:, :D] = -(gz.T @ ability)
            g_item[:, D] = gz.sum(0)
        return g_ability, g_item, None


class BernoulliLogLik(torch.autograd.Function):
    """masked_bernoulli_log_pdf(response, mask, response_mu).sum() on a
    materialised response_mu (vibo_bernoulli_loglik)."""

    @staticmethod
    def forward(ctx, response, mask, response_mu):
        want = bool(ctx.needs_input_grad[2])
        shape = response_mu.shape
        mu2 = response_mu.detach().reshape(response.shape).contiguous()
        ll, g = K.bernoulli_loglik(response, mask, mu2, want_grad=want)
        if want:
            ctx.save_for_backward(g)
        ctx.shape = shape
        return ll[0].to(torch.float32)

    @staticmethod
    def backward(ctx, gl):
        (g,) = ctx.saved_tensors
        return None, None, (gl * g).reshape(ctx.shape)
