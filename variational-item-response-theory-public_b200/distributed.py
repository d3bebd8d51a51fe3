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


class FusedStep:
    """The step of the UNCONDITIONAL product-of-experts model without flows as five launches of
    libvibo_b200.so and nothing else (no autograd, no torch kernels):

        vibo_param_forward[_draw] -> vibo_fused_elbo[_graph] (+ finalize) -> vibo_step_tail
        [-> all-reduce] -> vibo_adam_step   (peer-memory exchange: all-reduce + Adam are one kernel)

    Parameters, gradients and Adam moments live in flat buffers; the module's parameters are
    re-homed as views of ``theta`` (state_dict / load_state_dict keep working), ``p.grad`` are
    views of ``flat[1:]``.  Mirrors reference vibo.py:243-268 (train) / :285-312 (test)."""

    def __init__(self, trainer):
        import ctypes as C
        from . import _lib, kernels as K
        self.C, self.K, self.lib = C, K, _lib.load()
        self._check = _lib.check
        t = self.t = trainer
        m = t.model
        dev = t.flat.device
        mlp = m.ability_encoder.mlp
        want = [mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias, mlp[4].weight, mlp[4].bias,
                m.item_encoder.mu_lookup.weight, m.item_encoder.logvar_lookup.weight]
        assert len(t.params) == len(want) and all(a is b for a, b in zip(t.params, want)), \
            "unexpected parameter order"
        n = t.flat.numel() - 1
        self.n = n
        self.theta = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in t.params:
                k = p.numel()
                self.theta[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.theta[off:off + k].view_as(p)
                off += k
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.adam_step = torch.zeros(1, dtype=torch.int64, device=dev)
        I, F, D, H = m.num_item, m.item_feat_dim, m.ability_dim, m.hidden_dim
        self.I, self.F, self.D, self.H = I, F, D, H
        f32 = dict(dtype=torch.float32, device=dev)
        self.eps_item = torch.empty(I, F, **f32)
        self.item_feat = torch.empty(I, F, **f32)
        self.table = torch.empty(2, 1, 2 * D, **f32)
        self.hidden = torch.empty(2, 2, H, **f32)
        self.item_term = torch.empty(1, dtype=torch.float64, device=dev)
        self.scalars = torch.empty(2, dtype=torch.float64, device=dev)
        self.g_table = torch.empty(2, 1, 2 * D, **f32)
        self.g_item = torch.empty(I, F, **f32)
        self.lr = float(trainer.lr)

    @staticmethod
    def eligible(trainer, response):
        m = trainer.model
        return (trainer.on_gpu and response.is_cuda and m.ability_merge == 'product'
                and not m.conditional_posterior and m.n_norm_flows == 0 and m.hidden_dim <= 256
                and m.generative_model == 'irt' and m.response_dist == 'bernoulli'
                and getattr(m, "fuse_param_chain", True))

    def run(self, response, mask, train, eps_item=None, eps_ability=None):
        from . import functional as VF
        C, K, lib, t, m = self.C, self.K, self.lib, self.t, self.t.model
        resp, msk = VF.prepare_rows(response, mask)
        K._check_rows(resp, msk)
        P = resp.shape[0]
        dev = resp.device
        form = K.ELBO_KL if t.use_kl else K.ELBO_SAMPLE
        desc = K.make_desc(P, self.I, self.D, m.irt_num, False, m.ability_encoder.missing_policy, form,
                           t.person_offset)
        st = K._stream(dev)
        ptr = K._ptr
        mu, lv = t.params[6], t.params[7]
        w0, b0, w2, b2, w4, b4 = t.params[:6]
        seed = t.seed_state
        if eps_item is None:
            self._check(lib.vibo_param_forward_draw(C.byref(desc), self.H, ptr(mu), ptr(lv), ptr(seed), ptr(w0),
                                                    ptr(b0), ptr(w2), ptr(b2), ptr(w4), ptr(b4),
                                                    ptr(self.eps_item), ptr(self.item_feat), ptr(self.table),
                                                    ptr(self.hidden), ptr(self.item_term), st),
                        "vibo_param_forward_draw")
            e_i = self.eps_item
        else:
            e_i = eps_item.contiguous()
            self._check(lib.vibo_param_forward(C.byref(desc), self.H, ptr(mu), ptr(lv), ptr(e_i), ptr(w0), ptr(b0),
                                               ptr(w2), ptr(b2), ptr(w4), ptr(b4), ptr(self.item_feat),
                                               ptr(self.table), ptr(self.hidden), ptr(self.item_term), st),
                        "vibo_param_forward")
        ws = K.workspace(desc, dev)
        gt, gi = (ptr(self.g_table), ptr(self.g_item)) if train else (None, None)
        if eps_ability is None:
            self._check(lib.vibo_fused_elbo_graph(C.byref(desc), ptr(resp), ptr(msk), ptr(self.table),
                                                  ptr(self.item_feat), ptr(seed), C.c_float(t.beta),
                                                  ptr(self.scalars), None, None, None, gt, gi, ptr(ws),
                                                  ws.numel(), st), "vibo_fused_elbo_graph")
        else:
            self._check(lib.vibo_fused_elbo(C.byref(desc), ptr(resp), ptr(msk), ptr(self.table),
                                            ptr(self.item_feat), ptr(eps_ability.contiguous()), C.c_uint64(0),
                                            C.c_float(t.beta), ptr(self.scalars), None, None, None, gt, gi,
                                            ptr(ws), ws.numel(), st), "vibo_fused_elbo")
        flat = t.flat
        step_ptr = C.c_void_p(seed.data_ptr() + 8)  # seed_state[1]
        if train:
            g = [C.c_void_p(p.grad.data_ptr()) for p in t.params]
            grads = [g[6], g[7], g[0], g[1], g[2], g[3], g[4], g[5]]
        else:
            grads = [None] * 8
        self._check(lib.vibo_step_tail(C.byref(desc), self.H, C.c_float(t.beta), C.c_float(1.0 / t.world_size),
                                       ptr(self.scalars), ptr(self.item_term), ptr(flat), step_ptr,
                                       ptr(self.adam_step) if train else None, ptr(mu), ptr(lv), ptr(e_i),
                                       ptr(w2), ptr(w4), ptr(self.hidden), ptr(self.g_table), ptr(self.g_item),
                                       *grads, st), "vibo_step_tail")
        if train and t.world_size > 1 and t.peer is not None:
            # the exchange and the Adam step as one kernel over NVLink peer memory
            t.peer.all_reduce_adam_(flat, 1, self.theta, self.exp_avg, self.exp_avg_sq, self.adam_step, self.lr)
            return
        t._reduce(flat if train else flat[0:1])
        if train:
            self._check(lib.vibo_adam_step(self.n, ptr(self.theta), C.c_void_p(flat.data_ptr() + 4),
                                           ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.adam_step),
                                           C.c_float(self.lr), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8),
                                           st), "vibo_adam_step")


class ShardedElboTrainer:
    """One training / evaluation step of a drop-in VIBO module over this
    rank's rows: fused ELBO forward+backward, one all-reduce, Adam
    (reference vibo.py:243-268 with the whole resident shard as the batch)."""

    def __init__(self, model, lr=5e-3, world_size=1, rank=0, person_offset=0, beta=1.0,
                 use_kl_divergence=True, group=None, cuda_graph=False, seed=1234, allreduce="auto",
                 fused_step=True):
        self.model = model
        self.lr = lr
        self.want_fused_step = bool(fused_step)
        self.fused = None
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
        self.step_taken_generic = False  # a generic (torch.optim) train step ran: stay on that path
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

    def _use_fused(self, response):
        """The five-launch C-ABI step where it applies (decided on first use, then kept: the
        parameters move into its flat buffer)."""
        if self.world_size > 1 and self.peer is None and self.cuda_graph:
            return False  # torch.distributed exchange: graph(pre) -> all_reduce -> graph(Adam) form
        if self.fused is None and self.want_fused_step and self.step_taken_generic is False \
                and FusedStep.eligible(self, response):
            self.fused = FusedStep(self)
        return self.fused is not None and response.is_cuda

    def adam_steps(self):
        """Number of Adam updates applied so far (device-side count)."""
        if self.fused is not None:
            return int(self.fused.adam_step.item())
        steps = {int(st["step"].item()) if torch.is_tensor(st["step"]) else int(st["step"])
                 for st in self.opt.state.values()}
        return max(steps) if steps else 0

    def _train_pre(self, response, mask, seed, eps_item=None, eps_ability=None):
        self.flat.zero_()
        loss = self._loss(response, mask, seed, eps_item, eps_ability)
        # gradients straight into the flat buffer with ONE multi-tensor copy: loss.backward() would add each
        # parameter's gradient into its (zeroed) view with a launch of its own
        grads = torch.autograd.grad(loss, self.params, allow_unused=True)
        dst = [p.grad for p, g in zip(self.params, grads) if g is not None]
        if dst:
            torch._foreach_copy_(dst, [g for g in grads if g is not None])
        self.flat[0:1].copy_(loss.detach().reshape(1))

    def _train_post(self):
        self.opt.step()
        if self.seed_state is not None:
            self.seed_state[1] += 1

    def _train_body(self, response, mask, seed, eps_item=None, eps_ability=None):
        if self._use_fused(response):
            return self.fused.run(response, mask, True, eps_item, eps_ability)
        self.step_taken_generic = True
        self._train_pre(response, mask, seed, eps_item, eps_ability)
        self._reduce(self.flat)
        self._train_post()

    def _eval_pre(self, response, mask, seed, eps_item=None, eps_ability=None):
        with torch.no_grad():
            loss = self._loss(response, mask, seed, eps_item, eps_ability)
            self.flat[0:1].copy_(loss.reshape(1))

    def _eval_body(self, response, mask, seed, eps_item=None, eps_ability=None):
        if self._use_fused(response):
            return self.fused.run(response, mask, False, eps_item, eps_ability)
        self._eval_pre(response, mask, seed, eps_item, eps_ability)
        self._reduce(self.flat[0:1])
        if self.seed_state is not None:
            self.seed_state[1] += 1

    # ------------------------------------------------------------ CUDA graphs
    def _snapshot(self):
        state = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                 for p, st in self.opt.state.items()}
        fused = None
        if self.fused is not None:
            f = self.fused
            fused = (f.exp_avg.clone(), f.exp_avg_sq.clone(), f.adam_step.clone())
        return ([p.detach().clone() for p in self.params], state, self.seed_state.clone(), self.flat.clone(), fused)

    def _restore(self, snap):
        """Undo the eager warm-up runs that precede a capture: parameters, Adam moments and step
        counters (a freshly created state is all zeros), {seed, step}, and the flat buffer -- IN
        PLACE, so the tensors the graph captures are the ones the optimizer keeps using."""
        params, state, seed_state, flat, fused = snap
        if self.fused is not None:
            f = self.fused
            if fused is None:   # the fused step was created during the warm-up: fresh Adam state
                f.exp_avg.zero_(); f.exp_avg_sq.zero_(); f.adam_step.zero_()
            else:
                f.exp_avg.copy_(fused[0]); f.exp_avg_sq.copy_(fused[1]); f.adam_step.copy_(fused[2])
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
            key = (response.data_ptr(), mask.data_ptr(), tuple(response.shape), injected, float(self.beta),
                   self.person_offset)  # everything a capture bakes in by value
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
