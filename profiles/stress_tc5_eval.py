#!/usr/bin/env python
"""A/B stress of the tc5_eval stage hand-off: thousands of back-to-back launches of the
single-pass conditional evaluation on a shape with three tiles per CTA, comparing every launch's (LL, person term)
bit for bit.  VIBO_E5_DEBUG=128 restores the round-2 hand-off (arrive issued right behind the packers' loads),
0 is the shipped one, 144 releases a stage after all ballots.  Log: profiles/r02_tc5_eval_handoff_ab.log."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vibo_b200 as vb
vb._lib.load()
from test_gpu_parity import _synth
dev = "cuda"
P, I, D, irt = 148 * 128 * 3 + 77, 1000, 5, 3
resp, mask, table, item, eps = _synth(P, I, D, irt, True, 0.0, seed=5)
args = [torch.from_numpy(a).to(dev) for a in (resp, mask, table, item, eps)]
big = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
ref = None
def trial(tag, dbg, n, flush=False):
    global ref
    os.environ["VIBO_E5_DEBUG"] = str(dbg)
    outs = []
    for it in range(n):
        if flush: big.fill_(float(it))
        o = vb.kernels.fused_elbo(*args, irt_model=irt, conditional=True, beta=0.9, want_grads=False)
        outs.append(o["scalars"])
    torch.cuda.synchronize()
    v = torch.stack(outs).cpu().numpy()
    if ref is None:
        vals, counts = np.unique(v[:, 0], return_counts=True); ref = vals[np.argmax(counts)]
    bad = v[:, 0] != ref
    print(tag, "dbg", dbg, "n", n, "LL mismatches", int(bad.sum()), "max |dev|", float(np.abs(v[:, 0] - ref).max()),
          "term mismatches", int((v[:, 1] != v[0, 1]).sum()), flush=True)
for rep in range(3):
    trial("plain", 128, 5000)
    trial("plain", 0, 5000)
    trial("flush", 128, 300, flush=True)
    trial("flush", 0, 300, flush=True)
trial("plain", 144, 5000)   # late stage hand-off only, no fences
