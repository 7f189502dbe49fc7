"""CUDA-graph replay of one APGD iteration.

:class:`GraphedAttack` -- reached through ``GraphedModel.attack`` and used by ``apgd_train`` whenever the
model is a :class:`GraphedModel` -- captures the WHOLE iteration of semseg/attacker.py:385-551 as one graph:
image update (+ the previous iteration's row copies) -> model forward -> fused loss -> input-gradient
backward -> device bookkeeping.  What changes between iterations (iteration index, momentum a, the
check window k, the stage's eps and length) lives in a device control block that the bookkeeping kernel
advances itself (``robseg_apgd_step_ctl`` / ``robseg_apgd_bookkeep_ctl``), and the image buffers are
rotated in place, so the same two graphs (with / without backward) serve every iteration of every stage.
Per iteration the host issues one graph launch and the early-stop poll.

:class:`GraphedModel` alone (the first version of this row) replays only the consumer:

One APGD iteration is: image update -> ``model(x_adv)`` -> fused loss -> ``d logits / d x_adv``
-> bookkeeping (semseg/attacker.py:385-551).  The robseg kernels are a handful of launches; the
model's forward and input-gradient backward are several hundred small ones, and below ~8 images
per batch the GPU waits for the host to issue them.  :class:`GraphedModel` captures the two
passes of a frozen, eval-mode model once for a fixed input shape and replays them; the attack
(``apgd_train`` / ``apgd_largereps``) uses ``model(x)`` and ``model.input_grad(dlogits)`` when the
model offers them.  The fused-loss kernel writes its
gradient straight into the graph's static ``gout`` buffer, so no logits-sized copy is added.

The parameters are constants of the attack: their ``requires_grad`` is switched off while the
backward is captured so the graph holds only the input-gradient path (what
``torch.autograd.grad(loss, [x_adv])`` computes in the reference, attacker.py:350,469).
"""
import contextlib
import gc

import torch


# Capture with thread-local error checking: torch's default ("global") lets a CUDA call that is illegal during capture
# -- cudaHostAlloc, cudaEventQuery, cudaFree -- from ANY thread of the process invalidate the capture in progress.
# The attack is captured lazily, at the first batch of a loader loop, i.e. exactly while a DataLoader's pin-memory
# thread (tools/infer.py builds its loaders with pin_memory=True), an NCCL watchdog or a garbage-collected object of
# another thread may issue such calls; only this thread's own work goes into the graph, so only it is checked.
_CAPTURE_MODE = "thread_local"


@contextlib.contextmanager
def _capturing(graph, pool=None):
    """``torch.cuda.graph`` with the garbage collector out of the way.  Destroying a ``torch.cuda.CUDAGraph`` releases its
    memory pool (cudaFree), which is illegal while a stream of the same thread is capturing and invalidates that
    capture.  A dropped GraphedModel / GraphedAttack sits in a reference cycle, so its graphs die whenever the cyclic
    collector happens to run -- in the middle of the next capture, if it is triggered there (seen on the B200 box:
    'operation not permitted when stream is capturing (function reset)' five times, one per graph of the previous
    model, then cudaErrorStreamCaptureInvalidated at capture_end; torch >= 2.10 no longer collects on entering
    ``torch.cuda.graph`` unless torch.compiler.config.force_cudagraph_gc is set).  So: collect now, and keep the
    collector off until the capture has ended."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, pool=pool, capture_error_mode=_CAPTURE_MODE):
            yield
    finally:
        if was_enabled:
            gc.enable()


class GraphedModel:
    training = False  # apgd_train asserts ``not model.training`` (attacker.py:280)

    def __init__(self, model, example, warmup=3):
        if model.training:
            raise ValueError("GraphedModel captures an eval-mode model")
        if not example.is_cuda:
            raise RuntimeError("GraphedModel needs a CUDA example input (no CPU fallback)")
        self.model = model
        self.x = example.detach().float().contiguous().clone().requires_grad_(True)
        req = [p.requires_grad for p in model.parameters()]
        for p in model.parameters():
            p.requires_grad_(False)
        try:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up off the capture stream: cuDNN plans, lazy inits
                for _ in range(warmup):
                    out = model(self.x)
                    torch.autograd.grad(out, [self.x], grad_outputs=torch.zeros_like(out))
                del out
            cur.wait_stream(side)
            self.fwd = torch.cuda.CUDAGraph()
            with _capturing(self.fwd):
                self.logits = model(self.x)
            self.gout = torch.zeros_like(self.logits)
            self.bwd = torch.cuda.CUDAGraph()
            with _capturing(self.bwd, pool=self.fwd.pool()):
                (self.gx,) = torch.autograd.grad(self.logits, [self.x], grad_outputs=self.gout)
        finally:
            for p, r in zip(model.parameters(), req):
                p.requires_grad_(r)

    def eval(self):
        return self

    def parameters(self):
        return self.model.parameters()

    def __call__(self, x):
        """logits of x, in the graph's static output buffer (overwritten by the next call)."""
        if x.shape != self.x.shape:
            raise ValueError(f"GraphedModel captured for input {tuple(self.x.shape)}, got {tuple(x.shape)}")
        with torch.no_grad():
            self.x.copy_(x)
        self.fwd.replay()
        return self.logits

    def input_grad(self, gout):
        """(d logits / d x)^T gout for the last forward; gout may BE ``self.gout`` (no copy)."""
        if gout.data_ptr() != self.gout.data_ptr():
            self.gout.copy_(gout)
        self.bwd.replay()
        return self.gx


    def attack(self, keep_pred, early_stop, n_iter):
        """The :class:`GraphedAttack` for this model / batch shape (cached).  The argmax map of the best
        point is always tracked (8 B per pixel and iteration): apgd_largereps asks for it in its last stage
        only, and one runner -- one set of static buffers and captured graphs -- then serves all stages."""
        key = bool(early_stop)
        cache = self.__dict__.setdefault("_attacks", {})
        ga = cache.get(key)
        if ga is None or ga.n_iter_max < n_iter:
            ga = cache[key] = GraphedAttack(self, True, early_stop, max(512, n_iter))
        return ga


class GraphedAttack:
    """Static state + three captured graphs per loss kind (initial point, iteration with backward,
    last iteration without) sharing one memory pool.  See the module docstring."""

    def __init__(self, gm, keep_pred, early_stop, n_iter_max):
        from . import ops

        self.gm, self.model = gm, gm.model
        self.keep_pred, self.early_stop, self.n_iter_max = bool(keep_pred), bool(early_stop), int(n_iter_max)
        x = gm.x
        dev = x.device
        B = x.shape[0]
        self.B, self.n_pxl = B, x.shape[-2] * x.shape[-1]
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)  # noqa: E731
        self.x = torch.zeros_like(x, requires_grad=False)
        self.x_adv = torch.zeros_like(x).requires_grad_(True)  # the captured model input (leaf)
        self.x_old, self.x_best, self.x_best_adv = (torch.zeros_like(self.x) for _ in range(3))
        self.grad, self.grad_best = torch.zeros_like(self.x), torch.zeros_like(self.x)
        self.y = torch.zeros((B,) + tuple(x.shape[2:]), dtype=torch.int64, device=dev)
        self.acc, self.loss_best, self.loss_best_last, self.reduced_last, self.step = (z(B) for _ in range(5))
        self.loss_steps = z(self.n_iter_max, B)
        self.flags, self.done = z(3, B, dt=torch.int32), z(1, dt=torch.int32)
        self.ctl = ops.make_ctl(self.n_iter_max, dev)
        self.done_host = torch.zeros([1], dtype=torch.int32).pin_memory()
        self.pred_best = torch.zeros_like(self.y) if keep_pred else None
        self.counts_best = None  # [B,3,C] int64, allocated with the first capture (C is known then)
        self.weights = None  # [C] class weights, allocated on first use
        self.pool = None
        self.graphs = {}     # loss kind -> dict(init=, it=, last=, out0=)

    # ---- one iteration, as launched eagerly for warm-up and under capture -------------------------
    def _body(self, kind, track_loss, with_step, with_grad):
        from . import ops

        if with_step:
            ops.apgd_step_ctl(self.x, self.x_adv.detach(), self.x_old, self.grad, self.step, self.ctl, self.flags,
                              self.x_best_adv, self.x_best, self.grad_best)
        lowres = getattr(self.model, "forward_lowres", None)
        with torch.set_grad_enabled(with_grad):
            logits = lowres(self.x_adv) if lowres is not None else None
            fused = logits is not None and ops.can_fuse_upsample(logits, self.y)
            if not fused:
                logits = self.model(self.x_adv)
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()
        w = self.weights if kind == "mask-ce-bal" else None
        fn = ops.loss_upsampled_fwd_bwd if fused else ops.loss_fwd_bwd
        out = fn(logits, self.y, kind, w, want_grad=with_grad, want_pred=self.keep_pred, want_counts=True)
        if self.counts_best is None:
            self.counts_best = torch.zeros_like(out.counts)
        if with_grad:
            (gx,) = torch.autograd.grad(logits, [self.x_adv], grad_outputs=out.dlogits)
            self.grad.copy_(gx)
        track = out.track_img if track_loss in ("ce", "ce-avg") else out.loss_img
        if with_step:
            ops.apgd_bookkeep_ctl(out.correct, out.valid, track, self.acc, self.loss_best, self.loss_best_last,
                                  self.reduced_last, self.step, self.loss_steps, self.ctl, self.n_pxl,
                                  self.early_stop, self.flags, self.done)
            jobs = [(self.counts_best, out.counts, self.flags[0], None)]
            if self.keep_pred:
                jobs.append((self.pred_best, out.pred, self.flags[0], None))
            ops.row_select(jobs, self.B, self.x.device)
        return out, track

    def _capture(self, kind, track_loss):
        params = list(self.model.parameters())
        req = [p.requires_grad for p in params]
        for p in params:  # the attack differentiates w.r.t. the input only (attacker.py:350,469)
            p.requires_grad_(False)
        try:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up off the capture stream (cuDNN plans, workspaces)
                for _ in range(2):
                    self._body(kind, track_loss, True, True)
                    self._body(kind, track_loss, True, False)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            g = {}
            for name, (with_step, with_grad) in (("init", (False, True)), ("it", (True, True)), ("last", (True, False))):
                graph = torch.cuda.CUDAGraph()
                with _capturing(graph, pool=self.pool):
                    out, track = self._body(kind, track_loss, with_step, with_grad)
                if self.pool is None:
                    self.pool = graph.pool()
                g[name] = graph
                if name == "init":
                    g["out0"], g["track0"] = out, track
            self.graphs[(kind, track_loss)] = g
        finally:
            for p, r in zip(params, req):
                p.requires_grad_(r)
        return g

    def run_stage(self, x, y, x_adv0, eps, n_iter, kind, track_loss, weights, checks):
        """One ``apgd_train`` call (a stage of apgd_largereps) from the starting point ``x_adv0``.
        Returns fresh tensors ``(x_best, acc, loss_best, x_best_adv, pred_best, counts_best)``."""
        from . import ops

        if n_iter > self.n_iter_max:
            raise ValueError("stage longer than the captured control block")
        with torch.no_grad():
            self.x.copy_(x)
            self.y.copy_(y)
            if weights is not None and kind == "mask-ce-bal":
                if self.weights is None:
                    self.weights = torch.empty_like(weights)
                    self.graphs = {k: v for k, v in self.graphs.items() if k[0] != "mask-ce-bal"}
                self.weights.copy_(weights)
            # (capturing warms the iteration up eagerly, which moves x_adv: set the start point afterwards)
            g = self.graphs.get((kind, track_loss)) or self._capture(kind, track_loss)
            self.x_adv.copy_(x_adv0)
            ops.set_ctl(self.ctl, n_iter, eps, checks)
            self.loss_steps.zero_()
            self.flags.zero_()
            self.done.zero_()
            g["init"].replay()
            out0, track0 = g["out0"], g["track0"]
            # initial state (attacker.py:362-383); -1 pixels count as wrong here (:370-371)
            self.acc.copy_(out0.correct.float() / torch.full((), float(self.n_pxl), device=self.x.device))
            self.loss_best.copy_(track0)
            self.loss_best_last.copy_(track0)
            self.reduced_last.fill_(1.0)
            self.step.fill_(2.0 * eps)
            for t in (self.x_best, self.x_best_adv, self.x_old):
                t.copy_(self.x_adv)
            self.grad_best.copy_(self.grad)
            if self.keep_pred:
                self.pred_best.copy_(out0.pred)
            self.counts_best.copy_(out0.counts)
            copied = None
            for i in range(n_iter):
                g["it" if i < n_iter - 1 else "last"].replay()
                if self.early_stop:  # polled one iteration late; the device freezes the state itself
                    if copied is not None:
                        copied.synchronize()
                        if int(self.done_host[0]) != 0:
                            break
                    self.done_host.copy_(self.done, non_blocking=True)
                    copied = torch.cuda.Event()
                    copied.record()
            if n_iter > 0:  # the last iteration's pending row stores
                xa = self.x_adv.detach()
                ops.row_select([(self.x_best_adv, xa, self.flags[0], None), (self.x_best, xa, self.flags[1], None),
                                (self.grad_best, self.grad, self.flags[1], None)], self.B, self.x.device)
            return (self.x_best.clone(), self.acc.clone(), self.loss_best.clone(), self.x_best_adv.clone(),
                    self.pred_best.clone() if self.keep_pred else None, self.counts_best.clone())
