"""Person-sharded data parallelism for the VIBO ELBO step.

Persons are conditionally independent given the item sample, so rank r holds
a contiguous block of rows resident on its GPU and the only exchange per step
is ONE all-reduce (NCCL on GPUs, gloo in the CPU tests) of a flat float32
buffer ``[loss, every parameter gradient]``.  The item-side
prior term is weighted 1/world_size on each rank so that the reduced sum
counts it once; item noise must be identical on every rank (same torch seed),
ability noise is keyed by the global person index.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(num_person: int, rank: int, world_size: int):
    """Contiguous block [start, stop) of persons owned by `rank`."""
    base, rem = divmod(num_person, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class ShardedElboTrainer:
    """One training / evaluation step of a drop-in VIBO module over this
    rank's rows: fused ELBO forward+backward, one all-reduce, Adam
    (reference vibo.py:243-268 with the whole resident shard as the batch)."""

    def __init__(self, model, lr=5e-3, world_size=1, rank=0, person_offset=0, beta=1.0,
                 use_kl_divergence=True, group=None, cuda_graph=False, seed=1234):
        self.model = model
        self.world_size = world_size
        self.rank = rank
        self.person_offset = int(person_offset)
        self.beta = beta
        self.use_kl = use_kl_divergence
        self.group = group
        self.seed = seed
        self.params = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        # flat [loss | grads]; each p.grad is a view, so the all-reduce is one call
        self.flat = torch.zeros(1 + n, dtype=torch.float32, device=dev)
        off = 1
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        on_gpu = dev.type == "cuda"
        # one fused multi-tensor Adam kernel on the GPU (graph-capturable); plain Adam on CPU (tests)
        self.opt = torch.optim.Adam(self.params, lr=lr, capturable=True, fused=True) if on_gpu \
            else torch.optim.Adam(self.params, lr=lr, foreach=True)
        self.cuda_graph = bool(cuda_graph) and on_gpu
        self._graphs = {}
        self.graph_replays = 0
        self.kernels_per_step = 0

    @property
    def uses_fused(self):
        from . import _lib
        m = self.model
        d = _lib.Desc(1 << 20, m.num_item, m.ability_dim, m.irt_num, int(m.conditional_posterior), 0, 0, 0)
        import ctypes
        return (m.n_norm_flows == 0 and m.ability_merge == 'product'
                and bool(_lib.load().vibo_single_pass(ctypes.byref(d))))

    # ---------------------------------------------------------------- pieces
    def _loss(self, response, mask, seed):
        return self.model.fused_elbo(response, mask, annealing_factor=self.beta,
                                     use_kl_divergence=self.use_kl, seed=seed,
                                     person_offset=self.person_offset,
                                     item_term_scale=1.0 / self.world_size)

    def _train_pre(self, response, mask, seed):
        self.flat.zero_()
        loss = self._loss(response, mask, seed)
        loss.backward()
        self.flat[0:1].copy_(loss.detach().reshape(1))

    def _eval_pre(self, response, mask, seed):
        with torch.no_grad():
            loss = self._loss(response, mask, seed)
            self.flat[0:1].copy_(loss.reshape(1))

    def _train_body(self, response, mask, seed):
        self._train_pre(response, mask, seed)
        if self.world_size > 1:
            dist.all_reduce(self.flat, group=self.group)
        self.opt.step()
        return self.flat[0]

    def _eval_body(self, response, mask, seed):
        self._eval_pre(response, mask, seed)
        if self.world_size > 1:
            dist.all_reduce(self.flat[0:1], group=self.group)
        return self.flat[0]

    def _replay(self, key, fn, count_kernels=False):
        """Capture `fn()` in a CUDA graph on first use (after eager warm-up on a
        side stream), then replay it.  The NCCL all-reduce is deliberately NOT
        captured (capturing it hung on this stack): a training step is
        graph(pre) -> all_reduce -> graph(Adam).  Returns False if capture was
        refused, after which the caller runs eagerly."""
        from . import _lib
        g = self._graphs.get(key)
        if g is False:
            return False
        if g is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            lib = _lib.load()
            n0 = lib.vibo_launch_count()
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    fn()
            except Exception as exc:
                import warnings
                warnings.warn(f"CUDA graph capture of {key[0]} failed ({exc}); running eagerly")
                torch.cuda.synchronize()
                self._graphs[key] = False
                return False
            if count_kernels:
                self.kernels_per_step = int(lib.vibo_launch_count() - n0)
            self._graphs[key] = g
        g.replay()
        return True

    # ------------------------------------------------------------------ steps
    def train_step(self, response, mask, step_index=0, force_eager=False):
        """Returns the (all-reduced) loss as a 0-d tensor view (read it before
        the next step).  Graph replays draw noise from torch's graph-safe
        generator (seed=None); eager steps key the in-kernel Philox by
        ``seed + step_index``."""
        self.model.train()
        if self.cuda_graph and response.is_cuda and not force_eager:
            key = (response.data_ptr(), mask.data_ptr(), tuple(response.shape))
            if self._replay(("train_pre",) + key, lambda: self._train_pre(response, mask, None), True):
                if self.world_size > 1:
                    dist.all_reduce(self.flat, group=self.group)
                if not self._replay(("train_post",), self.opt.step):
                    self.opt.step()
                self.graph_replays += 1
                return self.flat[0]
        return self._train_body(response, mask, self.seed + step_index)

    def eval_step(self, response, mask, step_index=0, force_eager=False):
        self.model.eval()
        if self.cuda_graph and response.is_cuda and not force_eager:
            key = (response.data_ptr(), mask.data_ptr(), tuple(response.shape))
            if self._replay(("eval_pre",) + key, lambda: self._eval_pre(response, mask, None), True):
                if self.world_size > 1:
                    dist.all_reduce(self.flat[0:1], group=self.group)
                self.graph_replays += 1
                return self.flat[0]
        return self._eval_body(response, mask, self.seed + step_index)
