"""Workload for compute-sanitizer (memcheck / racecheck), round 2:

* the headline single-pass kernels (fused2_kernel for evaluation and narrow rows, the item-owner
  fused3_kernel for training with wide rows; in-kernel noise prologue) on shapes where every team
  of every CTA walks its stage ring >= 3 laps, plus a ragged last chunk and rows with missing cells;
* the round-2 kernels: step tail / flat Adam / param forward with in-kernel item noise (through
  ShardedElboTrainer's five-launch step), the sample-loop kernels, pack / unpack, the tcgen05
  per-cell MLP, the composed conditional path.

    compute-sanitizer --tool racecheck python profiles/sanitize_r02.py [fused|rest]

`fused` runs only the single-pass kernels, `rest` everything else (default: both).  With
VIBO_FUSED_DEBUG=3 fused2_kernel / fused_uncond_kernel put a team barrier in front of the stage
hand-off (see csrc/vibo_fused2_kernel.cuh): racecheck does not credit their production hand-off
(mbarrier arrive by the reading warps, wait by the refilling warp) as ordering the bulk copy after the
reads and reports it; with the barrier it reports nothing, i.e. there is no other hazard in those
kernels.  fused3_kernel hands a stage back through the team barrier itself and is reported clean as is.
With VIBO_E5_DEBUG=8 the single-pass conditional evaluation (tc5_eval_kernel) runs its epilogue / packer
warps and its link warps in lock-step through two 768-thread barriers per tile: the production kernel
orders their shared-memory hand-offs (theta, flags, bit tile) with mbarrier arrive / wait pairs, which
racecheck credits for bulk copies but not for ordinary loads and stores; in lock-step it reports nothing.
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vibo_b200  # noqa: E402
from vibo_b200 import kernels as K  # noqa: E402
from vibo_b200.distributed import ShardedElboTrainer  # noqa: E402

dev = torch.device("cuda:0")


def rows(P, I, missing, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    resp = (torch.rand(P, I, 1, generator=g, device="cuda") < 0.45).float()
    mask = torch.rand(P, I, 1, generator=g, device="cuda") >= missing
    resp[~mask] = -1.0
    return resp, mask


def fused(P, I, irt, D, missing):
    torch.manual_seed(1)
    cls = {1: vibo_b200.VIBO_1PL, 2: vibo_b200.VIBO_2PL, 3: vibo_b200.VIBO_3PL}[irt]
    model = cls(D, I, ability_merge="product").to(dev)
    resp, mask = rows(P, I, missing, 2)
    tr = ShardedElboTrainer(model, cuda_graph=False, seed=3)
    assert tr.uses_fused
    a = float(tr.train_step(resp, mask).item())
    b = float(tr.eval_step(resp, mask).item())
    torch.cuda.synchronize()
    print(f"fused P={P} I={I} {irt}PL D={D} missing={missing}: train {a:.3f} eval {b:.3f}", flush=True)


def main_fused():
    # I = 1000: 4 rows per stage, 5 teams, 148 CTAs, ring depth 2-3 -> 148 * 5 * 4 * 3 laps * 3 stages + ragged tail
    fused(148 * 5 * 4 * 10 + 3, 1000, 2, 1, 0.0)
    fused(148 * 5 * 16 * 10 + 7, 100, 2, 1, 0.1)     # narrow rows (C1's shape), missing cells
    fused(148 * 4 * 8 * 8 + 5, 500, 3, 1, 0.05)      # register-accumulator 3PL kernel
    fused(9000, 96, 2, 2, 0.0)                       # D = 2
    # item-owner training kernel (fused3_kernel): rows with missing cells (row-by-row path), one item group
    # per thread (I <= 512), 1PL
    fused(148 * 5 * 4 * 6 + 2, 1000, 2, 1, 0.08)
    fused(148 * 5 * 8 * 6 + 5, 500, 1, 1, 0.0)


def main_rest():
    # sample loops
    torch.manual_seed(0)
    model = vibo_b200.VIBO_3PL(2, 95, ability_merge="product").to(dev)
    resp, mask = rows(700, 95, 0.1, 5)
    print("log_marginal", float(model.log_marginal(resp, mask, 12, seed=1)))
    print("predictive", float(model.posterior_predictive_mean(resp, mask, 12, seed=1).sum()))
    # pack / unpack + packed host entry
    r2, m2 = resp[:, :, 0].contiguous(), mask[:, :, 0].contiguous().view(torch.uint8)
    pk = K.pack_rows(r2, m2)
    a, b = K.unpack_rows(pk)
    with torch.no_grad():
        print("packed host", float(model.fused_elbo(pk.cpu(), None, seed=4)))
    # host-buffer entry with chunks large enough for the host-compressed route (host thread pool packs a share
    # of every chunk into pinned buffers while the rest crosses PCIe as is)
    big_r, big_m = rows(3 * 8192 + 77, 200, 0.1, 11)
    hm = vibo_b200.VIBO_2PL(1, 200, ability_merge="product").to(dev)
    hm.host_chunk_person = 8192
    with torch.no_grad():
        print("host-compressed route", float(hm.fused_elbo(big_r.cpu().pin_memory(), big_m.cpu().pin_memory(), seed=4)),
              "share", K._lib.load().vibo_host_pack_share(C.byref(K.make_desc(8192, 200, 1, 2, False)), 8192))
    # tcgen05 per-cell MLP (ragged tile edges)
    u, v = torch.randn(95, 64, device=dev), torch.randn(701, 64, device=dev)
    out = K.percell_mlp(u, v, None, None, torch.randn(64, 64, device=dev) * 0.2, torch.randn(64, device=dev),
                        torch.randn(64, device=dev))
    print("percell", float(out.sum()))
    # conditional composition (tensor-core encode / encode-backward + link) through the generic trainer step
    cm = vibo_b200.VIBO_3PL(5, 1000, ability_merge="product", conditional_posterior=True).to(dev)
    cr, ck = rows(148 * 16 * 3 + 9, 1000, 0.0, 7)
    ct = ShardedElboTrainer(cm, cuda_graph=False)
    print("conditional", float(ct.train_step(cr, ck).item()))
    # single-pass conditional evaluation (tcgen05 encode + on-chip bit tile + link warps), ragged last tile
    print("conditional eval", float(ct.eval_step(cr, ck).item()))
    torch.cuda.synchronize()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "fused"):
        main_fused()
    if which in ("all", "rest"):
        main_rest()
