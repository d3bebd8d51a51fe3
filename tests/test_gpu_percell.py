"""Per-cell MLP of the nonlinear generative models on tcgen05 / TMEM (vibo_percell_mlp) against the
same math in PyTorch fp64.  Tolerance: the kernel carries the hidden activations and W2 as bf16
hi + lo pairs (3 products) -> ~1e-5 relative; asserted at 1e-4 of the output scale."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(u, v, z, w0, W2, c2, w4):
    pre = v.double()[:, None, :] + u.double()[None, :, :]
    if z is not None:
        pre = pre + z.double()[:, :, None] * w0.double()
    h = F.elu(pre)
    h = F.elu(h @ W2.double().T + c2.double())
    return h @ w4.double()


@pytest.mark.parametrize("P,I,link", [(8, 16, False), (64, 100, False), (37, 95, True), (1, 1, True),
                                      (300, 1000, False), (129, 17, True), (2048, 500, True)])
def test_percell_mlp_kernel_vs_torch(P, I, link):
    from vibo_b200 import kernels as K
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(P * 31 + I)
    H = 64
    W2 = (torch.randn(H, H, generator=g) * 0.25).to(dev)
    c2 = (torch.randn(H, generator=g) * 0.1).to(dev)
    w4 = (torch.randn(H, generator=g) * 0.3).to(dev)
    if link:
        u = (torch.randn(1, H, generator=g) * 0.1).to(dev)
        v = torch.zeros(1, H, device=dev)
        z = (torch.randn(P, I, generator=g) * 2.0).to(dev)
        w0 = torch.randn(H, generator=g).to(dev)
    else:
        u = torch.randn(I, H, generator=g).to(dev)
        v = torch.randn(P, H, generator=g).to(dev)
        z = w0 = None
    out = K.percell_mlp(u, v, z, w0, W2, c2, w4)
    torch.cuda.synchronize()
    want = _ref(u, v, z, w0, W2, c2, w4)
    assert out.shape == (P, I)
    scale = float(want.abs().max()) + 1e-6
    err = float((out.double() - want).abs().max())
    assert err <= 1e-4 * scale, (err, scale)


@pytest.mark.parametrize("gen,irt,D", [("link", 2, 1), ("link", 3, 2), ("deep", 1, 1), ("residual", 3, 2)])
def test_decoders_use_tensor_cores_without_grad(monkeypatch, gen, irt, D):
    import vibo_b200
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    cls = {1: vibo_b200.VIBO_1PL, 2: vibo_b200.VIBO_2PL, 3: vibo_b200.VIBO_3PL}[irt]
    P, I = 203, 77
    model = cls(D, I, ability_merge="product", generative_model=gen).to(dev)
    ability = torch.randn(P, D, device=dev)
    item = torch.randn(I, model.item_feat_dim, device=dev)
    lib = vibo_b200._lib.load()
    n0 = lib.vibo_launch_count()
    with torch.no_grad():
        got = model.decode(ability, item)
    assert lib.vibo_launch_count() > n0, "the tcgen05 kernel did not run"
    monkeypatch.setenv("VIBO_DISABLE_TCGEN05", "1")
    with torch.no_grad():
        want = model.decode(ability, item)
    assert got.shape == want.shape == (P, I, 1)
    assert float((got - want).abs().max()) <= 2e-5
    # with autograd active the PyTorch path runs (gradients flow)
    monkeypatch.delenv("VIBO_DISABLE_TCGEN05")
    n1 = lib.vibo_launch_count()
    model.decode(ability, item).sum().backward()
    assert lib.vibo_launch_count() == n1
    assert all(p.grad is not None for p in model.decoder.parameters())
