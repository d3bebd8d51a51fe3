"""The BENCHMARKED path is a tested path: ``ShardedElboTrainer`` under CUDA-graph replay
(what bench.py times) against ``oracle.reference_port.adam_train_step`` (reference
vibo.py:243-268 incl. ``optimizer.step()``) on identical injected noise; graph replays against
eager steps; person-keyed noise across shardings; and a 2-process run over the peer-memory
all-reduce (and NCCL) against the single-process run.

Tolerance: 1e-4 relative on the loss and on every parameter after Adam (BASELINE.json north
star); gradients per tensor 1e-4 rel-L2 plus the fp32 reference's own noise floor."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4


def _model(irt, D, I, cond, dev, seed=7, flows=0):
    import vibo_b200
    torch.manual_seed(seed)
    cls = {1: vibo_b200.VIBO_1PL, 2: vibo_b200.VIBO_2PL, 3: vibo_b200.VIBO_3PL}[irt]
    return cls(D, I, hidden_dim=64, ability_merge="product", conditional_posterior=cond,
               n_norm_flows=flows).to(dev)


def _rows(P, I, miss, seed):
    g = torch.Generator().manual_seed(seed)
    resp = (torch.rand(P, I, 1, generator=g) < 0.55).float()
    mask = torch.rand(P, I, 1, generator=g) >= miss
    resp[~mask] = -1.0
    return resp, mask


# (name, irt, D, I, cond, P, missing): C2's shape (2PL 500 items, single-pass kernel) and a slice of
# C3's shape (3PL D=5 conditional 1000 items: tensor-core encode + slab-stream link composition)
SHAPES = [("c2_shape", 2, 1, 500, False, 384, 0.0),
          ("c2_shape_missing", 2, 1, 500, False, 200, 0.1),
          ("c3_shape_slice", 3, 5, 1000, True, 160, 0.0)]


@pytest.mark.parametrize("name,irt,D,I,cond,P,miss", SHAPES, ids=[s[0] for s in SHAPES])
@pytest.mark.parametrize("graph", [True, False], ids=["graph", "eager"])
def test_trainer_steps_match_oracle_adam(name, irt, D, I, cond, P, miss, graph):
    from oracle import reference_port as RP
    from vibo_b200.distributed import ShardedElboTrainer
    dev = torch.device("cuda:0")
    model = _model(irt, D, I, cond, dev)
    if irt == 3:
        # A fresh N(0,1) 3PL D=5 item table is a SATURATING state: thousands of cells sit on the
        # eps32 clamp (zero gradient outside), where the fp32 reference itself is a knife edge
        # (fp32 vs fp64 reference: 2e-3 on the loss).  Multi-step Adam parity is only
        # meaningful off the clamp, so the item posterior is pulled in (reference fp32 vs fp64
        # then agree to 1e-7 over these 3 steps); the clamp itself is pinned by
        # test_saturated_cells and the m2pl_d1_unc_saturating fixture.
        with torch.no_grad():
            model.item_encoder.mu_lookup.weight.mul_(0.3)
            model.item_encoder.logvar_lookup.weight.mul_(0.1).sub_(3.0)
    params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    resp, mask = _rows(P, I, miss, seed=11)
    resp_d, mask_d = resp.to(dev), mask.to(dev)
    tr = ShardedElboTrainer(model, lr=5e-3, cuda_graph=graph)
    F = RP.item_feat_width(irt, D)
    g = torch.Generator().manual_seed(3)
    state = {}
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for step in range(3):
        e_i, e_a = torch.randn(I, F, generator=g), torch.randn(P, D, generator=g)
        ref = RP.adam_train_step(params, state, resp, mask.long(), e_i, e_a, lr=5e-3, irt_model=irt,
                                 ability_dim=D, conditional=cond)
        got = tr.train_step(resp_d, mask_d, eps_item=e_i.to(dev), eps_ability=e_a.to(dev))
        assert abs(float(got.item()) - float(ref)) <= TOL * abs(float(ref)), (step, float(got), float(ref))
    if graph:
        assert tr.graph_replays == 3, "the CUDA-graph path did not run"
    # gradients of the last step and parameters after three Adam updates
    leaves = state["leaves"]
    off = 1
    sd = dict(model.named_parameters())
    for k, p in sd.items():
        gk = tr.flat[off:off + p.numel()].view_as(p).cpu().numpy()
        off += p.numel()
        ref_g = leaves[k].grad.numpy()
        assert rel_l2(gk, ref_g) <= 3e-4, (k, rel_l2(gk, ref_g))
        # Adam normalises every coordinate's update to ~lr, which amplifies the relative error of
        # near-zero gradient coordinates: compare the three-step UPDATE per tensor (an extra or a
        # missing Adam application would be an O(1) error here), and the values at 1e-4 of scale
        upd = (p.detach().cpu() - params[k]).numpy()
        ref_upd = (leaves[k].detach() - params[k]).numpy()
        assert rel_l2(upd, ref_upd) <= 5e-3, (k, rel_l2(upd, ref_upd))
        ref_p = leaves[k].detach().numpy()
        assert np.abs(p.detach().cpu().numpy() - ref_p).max() <= TOL * max(np.abs(ref_p).max(), 0.1), k


@pytest.mark.parametrize("irt,D,I,cond,flows", [(2, 1, 500, False, 0), (3, 2, 96, True, 0), (2, 1, 95, False, 2)],
                         ids=["fused", "conditional", "flows"])
def test_graph_replays_equal_eager_steps(irt, D, I, cond, flows):
    """Same seed, same rows: N graph-replayed steps == N eager steps, bit for bit on the loss
    and to float rounding on the parameters (the warm-up runs before capture must leave no trace:
    no extra Adam updates, no advanced step counters)."""
    from vibo_b200.distributed import ShardedElboTrainer
    dev = torch.device("cuda:0")
    resp, mask = _rows(777, I, 0.05, seed=5)
    resp, mask = resp.to(dev), mask.to(dev)
    out = {}
    for graph in (False, True):
        model = _model(irt, D, I, cond, dev, flows=flows)
        tr = ShardedElboTrainer(model, lr=5e-3, cuda_graph=graph, seed=99, use_kl_divergence=(flows == 0))
        losses = [float(tr.train_step(resp, mask).item()) for _ in range(4)]
        losses.append(float(tr.eval_step(resp, mask).item()))
        out[graph] = (losses, {k: v.detach().clone() for k, v in model.state_dict().items()},
                      int(tr.seed_state[1].item()))
        if graph:
            assert tr.graph_replays == 5
    l0, s0, c0 = out[False]
    l1, s1, c1 = out[True]
    assert c0 == c1 == 5
    # same kernels on the same inputs (item and ability noise are both Philox(seed + step)):
    assert l0 == l1, (l0, l1)
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k


def test_fused_step_equals_generic_step():
    """The five-launch C-ABI step (FusedStep: own Adam kernel, in-kernel item noise) against the
    autograd + torch.optim.Adam step on the same Philox noise."""
    from vibo_b200.distributed import ShardedElboTrainer
    dev = torch.device("cuda:0")
    resp, mask = _rows(1500, 500, 0.05, seed=5)
    resp, mask = resp.to(dev), mask.to(dev)
    out = {}
    for fused in (True, False):
        model = _model(2, 1, 500, False, dev)
        tr = ShardedElboTrainer(model, lr=5e-3, cuda_graph=True, seed=3, fused_step=fused)
        losses = [float(tr.train_step(resp, mask).item()) for _ in range(5)]
        losses.append(float(tr.eval_step(resp, mask).item()))
        assert (tr.fused is not None) == fused
        out[fused] = (losses, {k: v.detach().clone() for k, v in model.state_dict().items()})
    assert np.allclose(out[True][0], out[False][0], rtol=2e-6), (out[True][0], out[False][0])
    for k in out[True][1]:
        assert torch.allclose(out[True][1][k], out[False][1][k], rtol=1e-4, atol=2e-6), k


def test_adam_state_after_graph_capture_is_one_step():
    """the eager warm-up before capture used to apply Adam 4 times on step 1."""
    from vibo_b200.distributed import ShardedElboTrainer
    dev = torch.device("cuda:0")
    model = _model(2, 1, 500, False, dev)
    resp, mask = _rows(512, 500, 0.0, seed=1)
    tr = ShardedElboTrainer(model, cuda_graph=True)
    tr.train_step(resp.to(dev), mask.to(dev))
    assert tr.adam_steps() == 1
    tr.train_step(resp.to(dev), mask.to(dev))
    assert tr.adam_steps() == 2
    assert int(tr.seed_state[1].item()) == 2
    # the same through the generic (autograd + torch.optim) step
    model = _model(2, 1, 500, False, dev)
    tr = ShardedElboTrainer(model, cuda_graph=True, fused_step=False)
    tr.train_step(resp.to(dev), mask.to(dev))
    assert tr.adam_steps() == 1 and tr.fused is None


def test_noise_is_keyed_by_global_person_not_by_shard():
    """The loss of a step over P persons == the sum over two shards with their person offsets,
    in graph mode (device-side Philox key) -- the property the N-rank run relies on."""
    dev = torch.device("cuda:0")
    P, I = 1200, 500
    resp, mask = _rows(P, I, 0.0, seed=2)
    resp, mask = resp.to(dev), mask.to(dev)
    e_i = torch.randn(I, 2, generator=torch.Generator().manual_seed(8)).to(dev)

    seed_state = torch.tensor([5, 2], dtype=torch.int64, device=dev)

    def loss_of(rows, offset, scale):
        model = _model(2, 1, I, False, dev)
        a, b = rows
        r, m = resp[a:b].contiguous(), mask[a:b].contiguous()
        # eps_item injected, ability noise drawn in-kernel from the device-side key
        with torch.no_grad():
            return float(model.fused_elbo(r, m, eps_item=e_i, seed=seed_state, person_offset=offset,
                                          item_term_scale=scale).item())

    whole = loss_of((0, P), 0, 1.0)
    parts = loss_of((0, 500), 0, 0.5) + loss_of((500, P), 500, 0.5)
    assert abs(whole - parts) <= 2e-6 * abs(whole), (whole, parts)


def test_in_kernel_philox_matches_fill_kernel():
    """fused kernel's in-kernel draw == vibo_philox_normal stream (same key, same person index)."""
    import vibo_b200
    from vibo_b200 import kernels as K
    dev = torch.device("cuda:0")
    P, I, D = 2309, 500, 2
    model = _model(2, D, I, False, dev)
    resp, mask = _rows(P, I, 0.0, seed=4)
    resp, mask = resp.to(dev), mask.to(dev)
    e_i = torch.randn(I, D + 1, device=dev)
    seed_state = torch.tensor([41, 3], dtype=torch.int64, device=dev)
    with torch.no_grad():
        a = model.fused_elbo(resp, mask, eps_item=e_i, seed=seed_state, person_offset=1000)
        eps = K.philox_normal(P, D, 44, 1000, dev)
        b = model.fused_elbo(resp, mask, eps_item=e_i, eps_ability=eps)
        c = model.fused_elbo(resp, mask, eps_item=e_i, seed=44, person_offset=1000)
    assert float(a) == float(c)
    assert abs(float(a) - float(b)) <= 1e-6 * abs(float(b))


# ------------------------------------------------------------------ 2 ranks
def _worker(rank, world, port, kind, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from vibo_b200.distributed import ShardedElboTrainer, shard_bounds
    P, I = 4001, 500
    resp, mask = _rows(P, I, 0.03, seed=6)
    a, b = shard_bounds(P, rank, world)
    res = {}
    for graph in (True, False):
        model = _model(2, 1, I, False, dev)
        tr = ShardedElboTrainer(model, lr=5e-3, world_size=world, rank=rank, person_offset=a, seed=17,
                                cuda_graph=graph, allreduce=kind)
        assert tr.allreduce_kind.startswith("peer" if kind == "peer" else "torch")
        torch.manual_seed(100)
        r, m = resp[a:b].to(dev), mask[a:b].to(dev)
        losses = [float(tr.train_step(r, m).item()) for _ in range(3)]
        losses.append(float(tr.eval_step(r, m).item()))
        if tr.peer is not None:
            tr.peer.status()
        res[graph] = (losses, {k: v.detach().cpu() for k, v in model.state_dict().items()})
        tr.close()
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["peer", "dist"])
def test_two_rank_equals_one_rank(tmp_path, kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from vibo_b200.distributed import ShardedElboTrainer
    world = 2
    port = 29600 + os.getpid() % 1000 + (7 if kind == "peer" else 0)
    mp.spawn(_worker, args=(world, port, kind, str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    # single process, same seeds
    dev = torch.device("cuda:0")
    P, I = 4001, 500
    resp, mask = _rows(P, I, 0.03, seed=6)
    model = _model(2, 1, I, False, dev)
    tr = ShardedElboTrainer(model, lr=5e-3, seed=17, cuda_graph=True)
    torch.manual_seed(100)
    r, m = resp.to(dev), mask.to(dev)
    losses = [float(tr.train_step(r, m).item()) for _ in range(3)]
    losses.append(float(tr.eval_step(r, m).item()))
    for graph in (True, False):
        # every rank holds the same reduced loss and the same parameters, bit for bit
        assert got[0][graph][0] == got[1][graph][0]
        for k in got[0][graph][1]:
            assert torch.equal(got[0][graph][1][k], got[1][graph][1][k]), k
        assert np.allclose(got[0][graph][0], losses, rtol=1e-5), (got[0][graph][0], losses)
        for k, v in model.state_dict().items():
            assert torch.allclose(got[0][graph][1][k], v.cpu(), rtol=1e-4, atol=1e-6), k
