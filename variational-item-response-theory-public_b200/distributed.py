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
        self.opt = torch.optim.Adam(self.params, lr=lr, capturable=on_gpu, foreach=True)
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
        return m.n_norm_flows == 0 and bool(_lib.load().vibo_single_pass(ctypes.byref(d)))

    # ---------------------------------------------------------------- pieces
    def _loss(self, response, mask, seed):
        return self.model.fused_elbo(response, mask, annealing_factor=self.beta,
                                     use_kl_divergence=self.use_kl, seed=seed,
                                     person_offset=self.person_offset,
                                     item_term_scale=1.0 / self.world_size)

    def _train_body(self, response, mask, seed):
        self.flat.zero_()
        loss = self._loss(response, mask, seed)
        loss.backward()
        self.flat[0:1].copy_(loss.detach().reshape(1))
        if self.world_size > 1:
            dist.all_reduce(self.flat, group=self.group)
        self.opt.step()
        return self.flat[0]

    def _eval_body(self, response, mask, seed):
        with torch.no_grad():
            loss = self._loss(response, mask, seed)
            self.flat[0:1].copy_(loss.reshape(1))
            if self.world_size > 1:
                dist.all_reduce(self.flat[0:1], group=self.group)
        return self.flat[0]

    def _graphed(self, kind, body, response, mask):
        """Capture the whole single-GPU step (small PyTorch ops, the fused kernel
        and Adam) in one CUDA graph keyed by the row buffers; noise
        comes from torch's graph-safe Philox generator so every replay draws
        fresh eps.  Falls back to eager launches if capture is refused."""
        from . import _lib
        key = (kind, response.data_ptr(), mask.data_ptr(), tuple(response.shape))
        g = self._graphs.get(key)
        if g is False:
            return None
        if g is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    body(response, mask, None)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            lib = _lib.load()
            n0 = lib.vibo_launch_count()
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    body(response, mask, None)   # seed=None: noise from torch's graph-safe generator
            except Exception as exc:  # e.g. a collective that cannot be captured
                import warnings
                warnings.warn(f"CUDA graph capture of the {kind} step failed ({exc}); running eagerly")
                torch.cuda.synchronize()
                self._graphs[key] = False
                return None
            self.kernels_per_step = int(lib.vibo_launch_count() - n0)
            self._graphs[key] = g
        g.replay()
        self.graph_replays += 1
        return self.flat[0]

    # ------------------------------------------------------------------ steps
    def train_step(self, response, mask, step_index=0, force_eager=False):
        """Returns the (all-reduced) loss as a 0-d tensor view (read it before
        the next step)."""
        self.model.train()
        # (NCCL collectives are kept out of graph capture: capturing them hung on this stack)
        if self.cuda_graph and response.is_cuda and not force_eager and self.world_size == 1:
            out = self._graphed("train", self._train_body, response, mask)
            if out is not None:
                return out
        return self._train_body(response, mask, self.seed + step_index)

    def eval_step(self, response, mask, step_index=0, force_eager=False):
        self.model.eval()
        if self.cuda_graph and response.is_cuda and not force_eager and self.world_size == 1:
            out = self._graphed("eval", self._eval_body, response, mask)
            if out is not None:
                return out
        return self._eval_body(response, mask, self.seed + step_index)
