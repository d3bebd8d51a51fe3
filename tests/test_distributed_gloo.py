"""Person-sharded data parallelism on CPU: world_size 2 over gloo, kernels
swapped for the numpy oracle.  The two-rank step must reproduce the one-rank
step on the concatenated rows (loss, every parameter gradient, parameters
after Adam): persons are independent given the item sample, the item-side
prior term is split 1/world_size, and ability noise is keyed by the global
person index."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Patch:
    """minimal stand-in for pytest's monkeypatch inside worker processes"""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def _make(irt, D, I, cond, seed):
    import vibo_b200
    torch.manual_seed(seed)
    cls = {1: vibo_b200.VIBO_1PL, 2: vibo_b200.VIBO_2PL, 3: vibo_b200.VIBO_3PL}[irt]
    return cls(D, I, ability_merge="product", conditional_posterior=cond)


def _data(P, I, seed):
    g = torch.Generator().manual_seed(seed)
    resp = (torch.rand(P, I, 1, generator=g) < 0.5).float()
    mask = torch.rand(P, I, 1, generator=g) >= 0.1
    resp[~mask] = -1.0
    return resp, mask


def _worker(rank, world, port, irt, D, I, cond, P, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_backend
    import vibo_b200
    from vibo_b200.distributed import ShardedElboTrainer, shard_bounds
    oracle_backend.install(_Patch)
    torch.set_num_threads(1)
    model = _make(irt, D, I, cond, seed=5)
    resp, mask = _data(P, I, seed=6)
    a, b = shard_bounds(P, rank, world)
    tr = ShardedElboTrainer(model, lr=1e-2, world_size=world, rank=rank, person_offset=a, seed=99)
    losses = []
    for step in range(2):
        torch.manual_seed(1000 + step)  # identical item noise on every rank
        losses.append(float(tr.train_step(resp[a:b], mask[a:b], step_index=step).item()))
    if rank == 0:
        torch.save({"loss": losses, "grad": tr.flat.clone(),
                    "state": {k: v.clone() for k, v in model.state_dict().items()}},
                   os.path.join(out_dir, "dist.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("irt,D,I,cond", [(2, 1, 24, False), (3, 2, 18, True)])
def test_two_rank_step_matches_single_rank(tmp_path, irt, D, I, cond):
    P, world = 37, 2
    port = 29500 + (os.getpid() + irt * 7) % 2000
    mp.spawn(_worker, args=(world, port, irt, D, I, cond, P, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dist.pt"))

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_backend
    import vibo_b200
    from vibo_b200.distributed import ShardedElboTrainer
    saved = {n: getattr(vibo_b200.kernels, n) for n in ("fused_elbo", "encode", "encode_backward",
                                                        "link_loglik", "decode", "bernoulli_loglik",
                                                        "_check_rows")}
    try:
        oracle_backend.install(_Patch)
        model = _make(irt, D, I, cond, seed=5)
        resp, mask = _data(P, I, seed=6)
        tr = ShardedElboTrainer(model, lr=1e-2, world_size=1, rank=0, person_offset=0, seed=99)
        losses = []
        for step in range(2):
            torch.manual_seed(1000 + step)
            losses.append(float(tr.train_step(resp, mask, step_index=step).item()))
    finally:
        for n, f in saved.items():
            setattr(vibo_b200.kernels, n, f)
    assert np.allclose(got["loss"], losses, rtol=1e-5), (got["loss"], losses)
    assert torch.allclose(got["grad"][1:], tr.flat[1:], rtol=1e-4, atol=1e-5)
    for k, v in model.state_dict().items():
        assert torch.allclose(got["state"][k], v, rtol=1e-4, atol=1e-6), k


def test_shard_bounds_cover_all_persons():
    from vibo_b200.distributed import shard_bounds
    for P in (0, 1, 7, 100, 1000003):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(P, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
