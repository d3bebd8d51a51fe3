"""Workload for compute-sanitizer (racecheck / memcheck) that makes every CTA of the streaming
kernels walk its stage ring several times (the parity cases under the sanitizers are too small
for that):  compute-sanitizer --tool racecheck python profiles/sanitize_multistage.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vibo_b200  # noqa: E402

K = vibo_b200.kernels
os.environ["VIBO_DISABLE_FUSED"] = "1"   # exercise the multi-pass kernels


def rows(P, I, missing, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    resp = (torch.rand(P, I, generator=g, device="cuda") < 0.45).float()
    mask = (torch.rand(P, I, generator=g, device="cuda") >= missing).to(torch.uint8) if missing > 0 \
        else torch.ones(P, I, dtype=torch.uint8, device="cuda")
    resp[mask == 0] = -1.0
    return resp, mask


def run(P, I, D, irt, cond, missing):
    g = torch.Generator(device="cuda").manual_seed(P + I)
    F = {1: 1, 2: D + 1, 3: D + 2}[irt]
    resp, mask = rows(P, I, missing, 1)
    table = 0.4 * torch.randn(2, I if cond else 1, 2 * D, generator=g, device="cuda")
    item = 0.5 * torch.randn(I, F, generator=g, device="cuda")
    eps = torch.randn(P, D, generator=g, device="cuda")
    out = K.fused_elbo(resp, mask, table, item, eps, irt_model=irt, conditional=cond, beta=0.7)
    torch.cuda.synchronize()
    print(f"P={P} I={I} D={D} {irt}PL cond={cond} missing={missing}: LL={out['scalars'][0].item():.3f}", flush=True)


if __name__ == "__main__":
    # grid = 148 CTAs: >= 4 chunks per CTA of 16 rows
    run(148 * 16 * 4 + 13, 1000, 5, 3, True, 0.0)     # tensor-core encode / encode-backward, link (4 items per lane)
    run(148 * 16 * 4 + 5, 640, 3, 2, True, 0.1)       # the same with missing cells (correction path, observed MMAs)
    run(148 * 8 * 48 * 2 + 7, 95, 1, 2, False, 0.1)   # slab-stream kernels, unaligned item count
    counts = K.person_counts(*rows(148 * 8 * 48 + 3, 95, 0.2, 3))
    torch.cuda.synchronize()
    print("counts", counts.sum().item())
