#!/usr/bin/env python
"""Timings of the round-2 kernels that are not on the bench.py line (CUDA events, warm, median):

* vibo_percell_mlp (tcgen05 / TMEM) against the same per-cell MLP in PyTorch (cuBLAS + elementwise);
* vibo_log_marginal (sample loop in the kernel) against S separate fused passes (round 1);
* vibo_predictive_mean against S decodes.

Prints one JSON object.
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vibo_b200  # noqa: E402
from vibo_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def percell():
    dev = torch.device("cuda:0")
    P, I, H = 100000, 500, 64
    g = torch.Generator(device=dev).manual_seed(0)
    u = torch.randn(I, H, device=dev, generator=g)
    v = torch.randn(P, H, device=dev, generator=g)
    W2 = torch.randn(H, H, device=dev, generator=g) * 0.2
    c2 = torch.randn(H, device=dev, generator=g) * 0.1
    w4 = torch.randn(H, device=dev, generator=g) * 0.3
    ms_k = timeit(lambda: K.percell_mlp(u, v, None, None, W2, c2, w4))

    def torch_path(chunk=2048):
        outs = []
        for a in range(0, P, chunk):
            pre = v[a:a + chunk, None, :] + u[None, :, :]
            h = F.elu(F.linear(F.elu(pre), W2, c2))
            outs.append(h @ w4)
        return torch.cat(outs)
    ms_t = timeit(torch_path, iters=3, warm=1)
    cells = P * I
    return {"shape": f"{P} persons x {I} items, hidden 64 (deep / residual form)",
            "tcgen05_kernel_ms": ms_k, "tcgen05_cells_per_s": cells / (ms_k * 1e-3),
            "useful_tflops": cells * (2 * H * H + 4 * H) / (ms_k * 1e-3) / 1e12,
            "tensor_tflops_issued_bf16": cells * 3 * 2 * H * H / (ms_k * 1e-3) / 1e12,
            "torch_cublas_ms": ms_t, "torch_cells_per_s": cells / (ms_t * 1e-3), "speedup_vs_torch": ms_t / ms_k}


def sample_loops():
    dev = torch.device("cuda:0")
    out = {}
    for name, P, I, S in (("cli_batch_16x100_S400", 16, 100, 400), ("c1_train_split_8000x100_S400", 8000, 100, 400),
                          ("100000x500_S32", 100000, 500, 32)):
        torch.manual_seed(0)
        model = vibo_b200.VIBO_2PL(1, I, ability_merge="product").to(dev)
        resp = (torch.rand(P, I, 1, device=dev) < 0.5).float()
        mask = torch.ones(P, I, 1, dtype=torch.bool, device=dev)
        ms_k = timeit(lambda: model.log_marginal(resp, mask, S, seed=1), iters=5, warm=2)

        def loop():
            with torch.no_grad():
                lw = torch.stack([-model.fused_elbo(resp, mask, use_kl_divergence=False) for _ in range(S)])
                return torch.logsumexp(lw, 0)
        ms_l = timeit(loop, iters=2, warm=1)
        ms_p = timeit(lambda: model.posterior_predictive_mean(resp, mask, S, seed=1), iters=5, warm=2)

        def ploop():
            with torch.no_grad():
                _, a_mu, a_lv, _, i_mu, i_lv = model.encode(resp, mask)
                acc = torch.zeros(P, I, 1, device=dev)
                for _ in range(S):
                    acc += model.decode(a_mu + torch.exp(0.5 * a_lv) * torch.randn_like(a_mu),
                                        i_mu + torch.exp(0.5 * i_lv) * torch.randn_like(i_mu))
                return acc / S
        ms_pl = timeit(ploop, iters=2, warm=1)
        out[name] = {"log_marginal_kernel_ms": ms_k, "log_marginal_S_passes_ms": ms_l, "speedup": ms_l / ms_k,
                     "cell_samples_per_s": P * I * S / (ms_k * 1e-3),
                     "predictive_kernel_ms": ms_p, "predictive_S_decodes_ms": ms_pl,
                     "predictive_speedup": ms_pl / ms_p}
    return out


if __name__ == "__main__":
    print(json.dumps({"percell_mlp": percell(), "sample_loops": sample_loops()}))
