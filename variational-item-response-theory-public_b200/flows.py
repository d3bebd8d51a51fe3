"""Planar normalizing flows on the (P, D) abilities and (I, F) item features
(``--n-norm-flows``).  Parameter containers and state_dict keys match the
reference (src/torch_core/flows.py:6-66: ``flows.{k}.{u,w,b}``, u, w ~ N(0, 1),
b = 1) so reference checkpoints load.  On the GPU the K planar steps and the flow-form
prior / entropy terms run fused per row (``vibo_flow_person_forward / _backward``) for the
abilities and for the item features alike; these modules are the autograd fallback.
"""
import torch
import torch.nn.functional as F
from torch import nn


class PlanarFlow(nn.Module):
    """f(z) = z + u_hat tanh(w.z + b) with u_hat made invertible
    (Rezende & Mohamed 2015); reference flows.py:21-41."""

    def __init__(self, in_features):
        super().__init__()
        self.u = nn.Parameter(torch.randn(in_features))
        self.w = nn.Parameter(torch.randn(in_features))
        self.b = nn.Parameter(torch.ones(1))

    def forward(self, z):
        w, u = self.w, self.u
        uw = (u * w).sum()
        u_hat = u + (F.softplus(uw) - 1.0 - uw) * w / (w * w).sum()
        h = torch.tanh(z @ w + self.b)
        f_z = z + h.unsqueeze(1) * u_hat.unsqueeze(0)
        slope = (1.0 - h * h) * (w * u_hat).sum()
        return f_z, torch.log(torch.abs(1.0 + slope) + 1e-8)


def planar_uhat(u, w):
    """The invertibility-corrected u of a planar flow (reference flows.py:26-29)."""
    uw = (u * w).sum()
    return u + (F.softplus(uw) - 1.0 - uw) * w / (w * w).sum()


class NormalizingFlows(nn.Module):
    """K planar flows in sequence; returns (z_K, sum_k log|det J_k|),
    reference flows.py:58-66."""

    def __init__(self, in_features, flow_type=PlanarFlow, n_flows=1):
        super().__init__()
        self.flows = nn.ModuleList([flow_type(in_features) for _ in range(n_flows)])

    def forward(self, z):
        total = 0
        for flow in self.flows:
            z, ldj = flow(z)
            total = total + ldj
        return z, total

    def stacked_parameters(self):
        """(uhat (K, D), w (K, D), b (K,)) for the fused planar-flow kernels; the invertibility
        correction of all K flows in one batch of elementwise ops (reference flows.py:26-29)."""
        first = self.flows[0]
        if first.u.is_cuda and len(self.flows) <= 8 and first.u.numel() <= 8 and first.u.dtype == torch.float32:
            # one kernel each way instead of ~13 forward / ~25 backward elementwise launches
            from . import functional as VF
            return VF.PlanarParams.apply(len(self.flows), *[f.u for f in self.flows],
                                         *[f.w for f in self.flows], *[f.b for f in self.flows])
        u = torch.stack([f.u for f in self.flows])
        w = torch.stack([f.w for f in self.flows])
        b = torch.cat([f.b for f in self.flows])
        uw = (u * w).sum(1, keepdim=True)
        uhat = u + (F.softplus(uw) - 1.0 - uw) * w / (w * w).sum(1, keepdim=True)
        return uhat, w, b
