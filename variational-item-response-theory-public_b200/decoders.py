"""Nonlinear generative models of the reference (``--generative-model link | deep | residual``,
src/torch_core/models.py:769-919) as drop-in modules: same parameter containers and
``state_dict`` keys (``decoder.link.{0,2,4}``, ``decoder.mlp_item_feat.*``,
``decoder.mlp_ability.*``, ``decoder.mlp_concat.*``), same construction / initialisation order
(so seeded initial weights match the reference bit for bit), evaluated WITHOUT materialising the
``(P, I, 2H)`` concatenation: the first layer of ``mlp_concat`` on ``[h_item_j, h_ability_i]`` is
separable, ``W_item h_item_j + W_ability h_ability_i + c``, so a cell's hidden pre-activation is
a per-item vector plus a per-person vector.

Every decoder is then a PER-CELL MLP with a rank-1-structured first layer,

    a1_ij = u_j + v_i            (deep / residual)        a1_ij = w0 z_ij + c0   (link)
    out_ij = w4 . ELU(W2 ELU(a1_ij) + c2) + c4

-- the one GEMM-shaped hot op of this path (M = cells, N = K = hidden): ``percell_tail`` runs it
on the tcgen05 kernel (``vibo_percell_mlp``, forward) where available and differentiates through
cuBLAS otherwise; persons are processed in chunks so activations stay bounded.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import init


def _weights_init(m):
    # reference models.py:512-518 (same in every decoder class)
    if isinstance(m, (nn.Linear, nn.Conv2d)):
        init.xavier_normal_(m.weight.data, gain=init.calculate_gain('relu'))
        init.constant_(m.bias.data, 0)


def irt_logit(ability, item_feat, irt_model):
    """irt_model_{1,2,3}pl(..., return_logit=True) of the reference (models.py:729-766):
    (logit (P, I, 1), guess (I,) or None)."""
    D = ability.shape[1]
    if irt_model == 1:
        return (ability.sum(1, keepdim=True) + item_feat[:, 0][None, :]).unsqueeze(2), None
    logit = (ability @ (-item_feat[:, :D].T) + item_feat[:, D][None, :]).unsqueeze(2)
    if irt_model == 2:
        return logit, None
    return logit, torch.sigmoid(item_feat[:, D + 1])


def percell_tail(pre1, lin2, lin3):
    """(…, H) first-layer pre-activations -> (…, 1): ELU -> Linear(H, H) -> ELU -> Linear(H, 1)."""
    h = F.elu(pre1)
    h = F.elu(F.linear(h, lin2.weight, lin2.bias))
    return F.linear(h, lin3.weight, lin3.bias)


def use_tensor_core_kernel(lin2, *tensors):
    """The tcgen05 kernel serves the no-gradient paths (evaluation, log-marginal, predictive
    sampling) on the GPU at hidden width 64; with autograd active the same math runs in PyTorch."""
    import os
    if os.environ.get("VIBO_DISABLE_TCGEN05") == "1":
        return False
    if not lin2.weight.is_cuda or lin2.weight.shape != (64, 64) or lin2.weight.dtype != torch.float32:
        return False
    if torch.is_grad_enabled() and (lin2.weight.requires_grad or any(t is not None and t.requires_grad
                                                                     for t in tensors)):
        return False
    return True


def percell(u, v, z, lin1_w0, lin2, lin3):
    """out (P, I, 1) = lin3(ELU(lin2(ELU(u_j + v_i + z_ij w0)))): the per-cell MLP every nonlinear
    decoder reduces to.  u (I or 1, H), v (P or 1, H), z (P, I, 1) / w0 (H) or None."""
    if use_tensor_core_kernel(lin2, u, v, z):
        from . import kernels as K
        out = K.percell_mlp(u, v, None if z is None else z[:, :, 0], lin1_w0, lin2.weight, lin2.bias,
                            lin3.weight[0])
        return (out + lin3.bias).unsqueeze(2)
    pre1 = v[:, None, :] + u[None, :, :]
    if z is not None:
        pre1 = pre1 + z * lin1_w0
    return percell_tail(pre1, lin2, lin3)


class LinkedIRT(nn.Module):
    """sigmoid(MLP(irt logit)) (reference models.py:769-808)."""

    def __init__(self, irt_model='1pl', hidden_dim=64):
        super().__init__()
        assert irt_model in ['1pl', '2pl', '3pl']
        self.irt_model = irt_model
        self.irt_num = int(irt_model[0])
        self.hidden_dim = hidden_dim
        self.link = nn.Sequential(
            nn.Linear(1, hidden_dim), nn.ELU(inplace=True),
            nn.Linear(hidden_dim, hidden_dim), nn.ELU(inplace=True),
            nn.Linear(hidden_dim, 1), nn.Sigmoid(),
        )
        self.apply(_weights_init)

    def forward(self, ability, item_feat):
        logit, guess = irt_logit(ability, item_feat, self.irt_num)
        lin1, lin2, lin3 = self.link[0], self.link[2], self.link[4]
        zero = torch.zeros(1, self.hidden_dim, dtype=logit.dtype, device=logit.device)
        # a1 = w0 z + c0: u = the bias row (broadcast over items), no per-person term
        prob = torch.sigmoid(percell(lin1.bias[None, :], zero, logit, lin1.weight[:, 0], lin2, lin3))
        if guess is not None:
            g = guess[None, :, None]
            return g + (1. - g) * prob
        return prob


class DeepIRT(nn.Module):
    """sigmoid(mlp_concat([mlp_item_feat(item), mlp_ability(ability)])) (reference models.py:811-877)."""

    def __init__(self, latent_dim, irt_model='1pl', hidden_dim=64):
        super().__init__()
        assert irt_model in ['1pl', '2pl', '3pl']
        self.latent_dim = latent_dim
        self.ability_dim = latent_dim
        self.irt_model = irt_model
        self.irt_num = int(irt_model[0])
        self.hidden_dim = hidden_dim
        self.item_feat_dim = {1: 1, 2: latent_dim + 1, 3: latent_dim + 2}[self.irt_num]
        H = hidden_dim

        def mlp(i, o):
            return nn.Sequential(nn.Linear(i, H), nn.ELU(inplace=True), nn.Linear(H, H), nn.ELU(inplace=True),
                                 nn.Linear(H, o))
        self.mlp_item_feat = mlp(self.item_feat_dim, H)
        self.mlp_ability = mlp(self.ability_dim, H)
        self.mlp_concat = mlp(2 * H, 1)
        self.apply(_weights_init)

    def residual_forward(self, ability, item_feat):
        H = self.hidden_dim
        hid_ability = self.mlp_ability(ability)          # (P, H)
        hid_item = self.mlp_item_feat(item_feat)         # (I, H)
        lin1, lin2, lin3 = self.mlp_concat[0], self.mlp_concat[2], self.mlp_concat[4]
        # cat([hid_item_j, hid_ability_i]) @ W1.T + c1, without the (P, I, 2H) tensor
        u = F.linear(hid_item, lin1.weight[:, :H], lin1.bias)      # (I, H)
        v = F.linear(hid_ability, lin1.weight[:, H:])              # (P, H)
        return percell(u, v, None, None, lin2, lin3)   # (P, I, 1)

    def forward(self, ability, item_feat):
        return torch.sigmoid(self.residual_forward(ability, item_feat))


class ResidualIRT(DeepIRT):
    """IRT logit + a deep residual (reference models.py:880-919)."""

    def __init__(self, latent_dim, irt_model='1pl', hidden_dim=64):
        super().__init__(latent_dim, irt_model=irt_model, hidden_dim=hidden_dim)
        self.apply(_weights_init)   # the reference's zero_init re-draws xavier-normal (models.py:912-918)

    def forward(self, ability, item_feat):
        res = self.residual_forward(ability, item_feat)
        logit, guess = irt_logit(ability, item_feat, self.irt_num)
        prob = torch.sigmoid(res + logit)
        if guess is not None:
            g = guess[None, :, None]
            return g + (1. - g) * prob
        return prob
