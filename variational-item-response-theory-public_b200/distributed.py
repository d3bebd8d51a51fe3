"""Person-sharded data parallelism for the VIBO ELBO step.

Persons are conditionally independent given the item sample, so rank r holds
a contiguous block of rows resident on its GPU and the only exchange per step
is ONE all-reduce of a flat float32 buffer ``[loss, every parameter gradient]``
(SURVEY.md 8e).  The item-side prior term is weighted 1/world_size on each
rank so that the reduced sum counts it once; item noise is identical on every
rank (same torch seed), ability noise is Philox keyed by (seed + step, GLOBAL
person index) -- in eager mode AND in CUDA-graph replays, where the key is
read from device memory (``seed_state``) when the kernels run -- so results do
not depend on how persons are sharded.

On GPUs the whole step -- zero_grad, fused forward/backward, all-reduce, Adam,
step counter -- is ONE CUDA graph: the exchange is a single kernel over NVLink
peer memory (``comm.PeerAllReduce``), not an eagerly launched NCCL call between
two graphs.  ``allreduce='dist'`` (torch.distributed: NCCL, or gloo in the CPU
tests) keeps the split form graph(pre) -> all_reduce -> graph(Adam).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


# "person index" range used for the item noise (persons are numbered from 0)
ITEM_NOISE_STREAM = 1 << 62


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
                 use_kl_divergence=True, group=None, cuda_graph=False, seed=1234, allreduce="auto"):
        self.model = model
        self.world_size = world_size
        self.rank = rank
        self.person_offset = int(person_offset)
        self.beta = beta
        self.use_kl = use_kl_divergence
        self.group = group
        self.seed = int(seed)
        self.step = 0  # steps taken so far; the default step_index of the next one
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
        self.on_gpu = on_gpu
        # one fused multi-tensor Adam kernel on the GPU (graph-capturable); plain Adam on CPU (tests)
        self.opt = torch.optim.Adam(self.params, lr=lr, capturable=True, fused=True) if on_gpu \
            else torch.optim.Adam(self.params, lr=lr, foreach=True)
        self.cuda_graph = bool(cuda_graph) and on_gpu
        # {seed, step}: the kernels read the Philox key seed + step from here at run time
        self.seed_state = torch.tensor([self.seed, 0], dtype=torch.int64, device=dev) if on_gpu else None
        self.peer = None
        if world_size > 1 and on_gpu and allreduce in ("auto", "peer"):
            from .comm import PeerAllReduce
            try:
                self.peer = PeerAllReduce(self.flat.numel(), dev, group=group)
            except Exception:
                if allreduce == "peer":
                    raise
                import warnings
                warnings.warn("peer-memory all-reduce unavailable; using torch.distributed all_reduce")
        self._graphs = {}
        self._inject = {}
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

    @property
    def allreduce_kind(self):
        if self.world_size == 1:
            return "none"
        return "peer-memory kernel (in graph)" if self.peer is not None else "torch.distributed"

    # ---------------------------------------------------------------- pieces
    def _loss(self, response, mask, seed, eps_item=None, eps_ability=None):
        if eps_item is None and self.on_gpu and seed is not None:
            # item noise: ONE global draw per step (models.py:361), identical on every rank and in
            # eager / graph mode: Philox(seed + step) on a counter range disjoint from the persons'
            from . import kernels as K
            m = self.model
            eps_item = K.philox_normal(m.num_item * m.item_feat_dim, 1, seed, ITEM_NOISE_STREAM,
                                       self.flat.device).view(m.num_item, m.item_feat_dim)
        return self.model.fused_elbo(response, mask, annealing_factor=self.beta,
                                     use_kl_divergence=self.use_kl, seed=seed,
                                     eps_item=eps_item, eps_ability=eps_ability,
                                     person_offset=self.person_offset,
                                     item_term_scale=1.0 / self.world_size)

    def _reduce(self, buf):
        if self.world_size == 1:
            return
        if self.peer is not None:
            self.peer.all_reduce_(buf)
        else:
            dist.all_reduce(buf, group=self.group)

    def _train_pre(self, response, mask, seed, eps_item=None, eps_ability=None):
        self.flat.zero_()
        loss = self._loss(response, mask, seed, eps_item, eps_ability)
        loss.backward()
        self.flat[0:1].copy_(loss.detach().reshape(1))

    def _train_post(self):
        self.opt.step()
        if self.seed_state is not None:
            self.seed_state[1] += 1

    def _train_body(self, response, mask, seed, eps_item=None, eps_ability=None):
        self._train_pre(response, mask, seed, eps_item, eps_ability)
        self._reduce(self.flat)
        self._train_post()

    def _eval_pre(self, response, mask, seed, eps_item=None, eps_ability=None):
        with torch.no_grad():
            loss = self._loss(response, mask, seed, eps_item, eps_ability)
            self.flat[0:1].copy_(loss.reshape(1))

    def _eval_body(self, response, mask, seed, eps_item=None, eps_ability=None):
        self._eval_pre(response, mask, seed, eps_item, eps_ability)
        self._reduce(self.flat[0:1])
        if self.seed_state is not None:
            self.seed_state[1] += 1

    # ------------------------------------------------------------ CUDA graphs
    def _snapshot(self):
        state = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                 for p, st in self.opt.state.items()}
        return ([p.detach().clone() for p in self.params], state, self.seed_state.clone(), self.flat.clone())

    def _restore(self, snap):
        """Undo the eager warm-up runs that precede a capture: parameters, Adam moments and step
        counters (a freshly created state is all zeros), {seed, step}, and the flat buffer -- IN
        PLACE, so the tensors the graph captures are the ones the optimizer keeps using."""
        params, state, seed_state, flat = snap
        with torch.no_grad():
            for p, saved in zip(self.params, params):
                p.copy_(saved)
            for p, st in self.opt.state.items():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if p in state and k in state[p]:
                            v.copy_(state[p][k])
                        else:
                            v.zero_()
            self.seed_state.copy_(seed_state)
            self.flat.copy_(flat)

    def _replay(self, key, fn, count_kernels=False):
        """Capture `fn()` in a CUDA graph on first use, then replay it.  Capture needs eager
        warm-up runs (lazy optimizer state, workspaces, cuBLAS handles); their side effects
        (Adam updates, step counters) are rolled back before the capture, so the first replay IS
        step 1.  Every rank runs the same warm-ups, so collectives inside `fn` stay matched.
        Returns False if capture was refused, after which the caller runs eagerly."""
        from . import _lib
        g = self._graphs.get(key)
        if g is False:
            return False
        if g is None:
            snap = self._snapshot()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
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
                self._restore(snap)
                self._graphs[key] = False
                return False
            if count_kernels:
                self.kernels_per_step = int(lib.vibo_launch_count() - n0)
            torch.cuda.synchronize()
            self._restore(snap)
            self._graphs[key] = g
        g.replay()
        return True

    def _static_eps(self, response, eps_item, eps_ability):
        """Injected noise for graph replays lives in static buffers the graph reads."""
        if eps_item is None and eps_ability is None:
            return None, None, False
        assert eps_item is not None and eps_ability is not None, "inject both eps_item and eps_ability"
        key = (tuple(eps_item.shape), tuple(eps_ability.shape))
        bufs = self._inject.get(key)
        if bufs is None:
            bufs = (torch.empty_like(eps_item), torch.empty_like(eps_ability))
            self._inject[key] = bufs
        bufs[0].copy_(eps_item)
        bufs[1].copy_(eps_ability)
        return bufs[0], bufs[1], True

    def _run(self, kind, response, mask, step_index, force_eager, eps_item, eps_ability):
        train = kind == "train"
        if step_index is None:
            step_index = self.step
        self.step = step_index + 1
        body = self._train_body if train else self._eval_body
        if not (self.on_gpu and response.is_cuda):
            # CPU tests (oracle backend) and host-resident rows: eager, integer Philox key
            seed = None if eps_ability is not None else self.seed + step_index
            if train:
                self._train_pre(response, mask, seed, eps_item, eps_ability)
                self._reduce(self.flat)
                self.opt.step()
            else:
                self._eval_pre(response, mask, seed, eps_item, eps_ability)
                self._reduce(self.flat[0:1])
            return self.flat[0]
        # the device-side {seed, step} is authoritative on the GPU; keep it in sync with step_index
        if int(step_index) != getattr(self, "_dev_step", 0):
            self.seed_state[1] = int(step_index)
        self._dev_step = int(step_index) + 1
        e_i, e_a, injected = (eps_item, eps_ability, eps_item is not None)
        seed = None if injected else self.seed_state
        if self.cuda_graph and not force_eager:
            e_i, e_a, injected = self._static_eps(response, eps_item, eps_ability)
            key = (response.data_ptr(), mask.data_ptr(), tuple(response.shape), injected)
            one_graph = self.world_size == 1 or self.peer is not None
            if one_graph:
                if self._replay((kind,) + key, lambda: body(response, mask, seed, e_i, e_a), True):
                    self.graph_replays += 1
                    return self.flat[0]
            else:
                pre = self._train_pre if train else self._eval_pre
                if self._replay((kind + "_pre",) + key, lambda: pre(response, mask, seed, e_i, e_a), True):
                    self._reduce(self.flat if train else self.flat[0:1])
                    if train:
                        if not self._replay(("train_post",), self._train_post):
                            self._train_post()
                    else:
                        self.seed_state[1] += 1
                    self.graph_replays += 1
                    return self.flat[0]
        body(response, mask, seed, e_i, e_a)
        return self.flat[0]

    # ------------------------------------------------------------------ steps
    def train_step(self, response, mask, step_index=None, force_eager=False, eps_item=None,
                   eps_ability=None):
        """zero_grad + fused forward/backward + all-reduce + Adam.  Returns the (all-reduced) loss
        as a 0-d tensor view (read it before the next step).  ``step_index`` defaults to the
        trainer's own step counter; the ability noise of step k is Philox(seed + k, global person
        index) whether the step runs eagerly or as a graph replay.  ``eps_item`` /
        ``eps_ability`` inject pre-drawn noise instead (parity tests)."""
        self.model.train()
        return self._run("train", response, mask, step_index, force_eager, eps_item, eps_ability)

    def eval_step(self, response, mask, step_index=None, force_eager=False, eps_item=None,
                  eps_ability=None):
        self.model.eval()
        return self._run("eval", response, mask, step_index, force_eager, eps_item, eps_ability)

    def close(self):
        if self.peer is not None:
            self.peer.close()
            self.peer = None
