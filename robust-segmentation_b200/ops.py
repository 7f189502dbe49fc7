"""Tensor-level wrappers over the C ABI, plus the ``torch.ops.robseg.*`` registrations.

Every function here launches hand-written sm_100a kernels from librobseg_b200.so on the
caller's current CUDA stream.  CPU tensors are rejected: there is no fallback path.
"""
from collections import namedtuple

import torch

from . import _lib

KIND_IDS = {
    "ce": _lib.LOSS_CE, "ce-avg": _lib.LOSS_CE, "pgd": _lib.LOSS_CE,
    "mask-ce-avg": _lib.LOSS_MASK_CE, "mask-ce-bal": _lib.LOSS_MASK_CE_BAL,
    "js-avg": _lib.LOSS_JS, "argmax": _lib.LOSS_ARGMAX,
}

LossOut = namedtuple("LossOut", "loss_img track_img correct valid dlogits pred loss_pix counts", defaults=(None,))

_workspaces = {}

# Optional per-launch device timing (bench.py turns it on inside its timed region): every
# wrapper brackets its C call with CUDA events on the launching stream and appends
# (name, algorithmic_bytes, start_event, end_event) here.
_prof = None
_prof_k = []   # kernel-level brackets of the same window (robseg_profile_next_kernel)
_last_k = []


def profile_start():
    global _prof, _prof_k
    _prof, _prof_k = [], []


def profile_stop():
    """Returns [(name, bytes, ms)] -- one event bracket per C call; call after a device synchronize.
    ``profile_kernels()`` then returns the brackets the library itself recorded around the main kernel of
    every loss call of the same window (the zeroing / fold / finalize launches of the call left outside)."""
    global _prof, _prof_k, _last_k
    rec, _prof = _prof or [], None
    _last_k = [(n, b, s.elapsed_time(e)) for n, b, s, e in _prof_k]
    _prof_k = []
    return [(n, b, s.elapsed_time(e)) for n, b, s, e in rec]


def profile_kernels():
    return list(_last_k)


class _timed:
    def __init__(self, name, nbytes, main_kernel=False):
        self.name, self.nbytes, self.main_kernel = name, nbytes, main_kernel

    def __enter__(self):
        self.on = _prof is not None and not torch.cuda.is_current_stream_capturing()
        if self.on:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            if self.main_kernel:
                # the library re-records these two around its main kernel; recording them here first creates
                # the cudaEvent_t handles (torch makes them lazily)
                self.ks = torch.cuda.Event(enable_timing=True)
                self.ke = torch.cuda.Event(enable_timing=True)
                self.ks.record()
                self.ke.record()
                _lib.load().robseg_profile_next_kernel(self.ks.cuda_event, self.ke.cuda_event)
            self.s.record()

    def __exit__(self, *a):
        if self.on and _prof is not None:
            self.e.record()
            _prof.append((self.name, self.nbytes, self.s, self.e))
            if self.main_kernel:
                _prof_k.append((self.name, self.nbytes, self.ks, self.ke))


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("robseg-b200 ops need CUDA tensors (no CPU fallback)")


def _workspace(dev, nbytes):
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def loss_fwd_bwd(logits, labels, kind, weights=None, grad_scale=None, upstream=None,
                 want_grad=True, want_pred=False, want_loss_pix=False, ignore_index=-1,
                 dlogits_out=None, want_stats=True, want_counts=False):
    """Fused softmax + loss + dlogits + argmax + per-image sums (robseg_loss_fwd_bwd).

    logits [B,C,*spatial] fp32/bf16 (contiguous), labels [B,*spatial] int64.
    grad_scale: None (1/HW per image), python float, or [B] tensor.
    want_counts: also return ``counts`` [B,3,C] int64 -- the per-image intersection / target /
    prediction class counters of compute_iou_acc taken in the same pass (robseg_loss_fwd_bwd_counts).
    Returns LossOut; fields not requested are None.
    """
    _need_cuda(logits, labels, weights, upstream)
    lib = _lib.load()
    if logits.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"logits dtype {logits.dtype} not supported (fp32 / bf16)")
    logits = logits.detach()
    if not logits.is_contiguous():
        logits = logits.contiguous()
    B, Cn = logits.shape[0], logits.shape[1]
    HW = logits[0, 0].numel()
    labels = labels.detach()
    if labels.dtype != torch.int64:
        labels = labels.long()
    labels = labels.contiguous()
    if labels.numel() != B * HW:
        raise ValueError(f"labels shape {tuple(labels.shape)} does not match logits {tuple(logits.shape)}")
    dev = logits.device
    kid = KIND_IDS[kind]
    if weights is not None:
        weights = weights.detach().to(device=dev, dtype=torch.float32).contiguous()
        if weights.numel() != Cn:
            raise ValueError("class weights must have C entries")
    if grad_scale is not None and not torch.is_tensor(grad_scale):
        grad_scale = torch.full((B,), float(grad_scale), dtype=torch.float32, device=dev)
    if grad_scale is not None:
        grad_scale = grad_scale.detach().to(device=dev, dtype=torch.float32).contiguous()
        if grad_scale.numel() == 1:
            grad_scale = grad_scale.reshape(1).expand(B).contiguous()
    if upstream is not None:
        upstream = upstream.detach().to(dtype=torch.float32).contiguous()
    want_grad = want_grad and kid != _lib.LOSS_ARGMAX
    if want_grad:
        dlogits = dlogits_out if dlogits_out is not None else torch.empty_like(logits)
        if dlogits.shape != logits.shape or dlogits.dtype != logits.dtype or not dlogits.is_contiguous():
            raise ValueError("dlogits_out must match logits")
    else:
        dlogits = None
    spatial = logits.shape[2:]
    pred = torch.empty((B, *spatial), dtype=torch.int64, device=dev) if want_pred else None
    loss_pix = torch.empty((B, *spatial), dtype=torch.float32, device=dev) if want_loss_pix else None
    if want_stats:
        fstat = torch.empty((2, B), dtype=torch.float32, device=dev)
        istat = torch.empty((2, B), dtype=torch.int32, device=dev)
    else:
        fstat = istat = None
    dt = _lib.F32 if logits.dtype == torch.float32 else _lib.BF16
    nws = lib.robseg_loss_workspace_bytes(B, Cn, HW, dt)
    ws = _workspace(dev, nws)
    # algorithmic bytes: logits read (+ gradient write) + int64 labels
    # (+ int64 argmax)
    nbytes = B * Cn * HW * logits.element_size() * (2 if want_grad else 1) + 8 * B * HW * (2 if want_pred else 1)
    counts = torch.empty((B, 3, Cn), dtype=torch.int64, device=dev) if want_counts else None
    head = (logits.data_ptr(), dt, labels.data_ptr(), _ptr(weights), kid, int(ignore_index), B, Cn,
            HW, _ptr(grad_scale), _ptr(upstream), _ptr(dlogits), _ptr(loss_pix), _ptr(pred),
            _ptr(fstat[0]) if want_stats else 0, _ptr(fstat[1]) if want_stats else 0,
            _ptr(istat[0]) if want_stats else 0, _ptr(istat[1]) if want_stats else 0)
    # (a counted launch brackets three more small kernels -- zeroing and folding the counter replicas -- so it is
    # profiled under its own name: bench.py's roofline is the loss kernel's own launch duration)
    tag = ("loss_grad" if want_grad else "loss_only") + ("_counts" if want_counts else "")
    with torch.cuda.device(dev), _timed(tag, nbytes, main_kernel=True):
        if want_counts:
            rc = lib.robseg_loss_fwd_bwd_counts(*head, counts.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        else:
            rc = lib.robseg_loss_fwd_bwd(*head, ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "robseg_loss_fwd_bwd")
    _lib.count((2 if want_stats else 1) + (2 if want_counts else 0))  # (+ counter zeroing and fold kernels)
    if want_stats:
        return LossOut(fstat[0], fstat[1], istat[0], istat[1], dlogits, pred, loss_pix, counts)
    return LossOut(None, None, None, None, dlogits, pred, loss_pix, counts)


FUSED_UP_RATIOS = (2, 4, 8, 16)


def can_fuse_upsample(low, labels):
    """True when ``labels`` are an integer x2/4/8/16 up-sampling of fp32 ``low`` [B,C,h,w]."""
    if low is None or low.dim() != 4 or low.dtype != torch.float32 or not low.is_cuda:
        return False
    h, w = low.shape[-2:]
    H, W = labels.shape[-2:]
    return H % h == 0 and W % w == 0 and H // h == W // w and H // h in FUSED_UP_RATIOS


def loss_upsampled_fwd_bwd(low, labels, kind, weights=None, grad_scale=None, want_grad=True,
                           want_pred=False, ignore_index=-1, dlow_out=None, want_stats=True, want_counts=False):
    """``loss_fwd_bwd(F.interpolate(low, labels.shape[-2:], mode="bilinear"), ...)`` without the
    [B,C,H,W] logits: the loss kernel interpolates on the fly and returns the gradient with respect to
    ``low`` (robseg_loss_upsampled_fwd_bwd).  ``LossOut.dlogits`` is ``dlow``
    [B,C,h,w]; ratios 2/4/8/16, fp32."""
    _need_cuda(low, labels, weights)
    lib = _lib.load()
    if low.dtype != torch.float32 or low.dim() != 4:
        raise TypeError("loss_upsampled_fwd_bwd expects 4-D float32 low-resolution logits")
    low = low.detach().contiguous()
    B, Cn, h, w = low.shape
    labels = labels.detach()
    if labels.dtype != torch.int64:
        labels = labels.long()
    labels = labels.contiguous()
    H, W = labels.shape[-2:]
    if labels.numel() != B * H * W or H % h or W % w or H // h != W // w or H // h not in FUSED_UP_RATIOS:
        raise ValueError(f"labels {tuple(labels.shape)} are not an integer x2/4/8/16 up-sampling of {tuple(low.shape)}")
    dev = low.device
    kid = KIND_IDS[kind]
    if weights is not None:
        weights = weights.detach().to(device=dev, dtype=torch.float32).contiguous()
        if weights.numel() != Cn:
            raise ValueError("class weights must have C entries")
    if grad_scale is not None and not torch.is_tensor(grad_scale):
        grad_scale = torch.full((B,), float(grad_scale), dtype=torch.float32, device=dev)
    if grad_scale is not None:
        grad_scale = grad_scale.detach().to(device=dev, dtype=torch.float32).contiguous()
        if grad_scale.numel() == 1:
            grad_scale = grad_scale.reshape(1).expand(B).contiguous()
    want_grad = want_grad and kid != _lib.LOSS_ARGMAX
    dlow = None
    if want_grad:
        dlow = dlow_out if dlow_out is not None else torch.empty_like(low)
        if dlow.shape != low.shape or dlow.dtype != low.dtype or not dlow.is_contiguous():
            raise ValueError("dlow_out must match low")
    pred = torch.empty((B, H, W), dtype=torch.int64, device=dev) if want_pred else None
    if want_stats:
        fstat = torch.empty((2, B), dtype=torch.float32, device=dev)
        istat = torch.empty((2, B), dtype=torch.int32, device=dev)
    ws = _workspace(dev, lib.robseg_loss_upsampled_workspace_bytes(B, Cn, h, w, H, W))
    # algorithmic bytes: labels (+ argmax map) + the low-resolution tensors
    nbytes = 8 * B * H * W * (2 if want_pred else 1) + low.numel() * 4 * (2 if want_grad else 1)
    counts = torch.empty((B, 3, Cn), dtype=torch.int64, device=dev) if want_counts else None
    head = (low.data_ptr(), labels.data_ptr(), _ptr(weights), kid, int(ignore_index), B, Cn, h, w, H, W,
            _ptr(grad_scale), _ptr(dlow), _ptr(pred),
            _ptr(fstat[0]) if want_stats else 0, _ptr(fstat[1]) if want_stats else 0,
            _ptr(istat[0]) if want_stats else 0, _ptr(istat[1]) if want_stats else 0)
    tag = ("loss_up_grad" if want_grad else "loss_up_only") + ("_counts" if want_counts else "")
    with torch.cuda.device(dev), _timed(tag, nbytes, main_kernel=True):
        if want_counts:
            rc = lib.robseg_loss_upsampled_fwd_bwd_counts(*head, counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                                          _stream())
        else:
            rc = lib.robseg_loss_upsampled_fwd_bwd(*head, ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "robseg_loss_upsampled_fwd_bwd")
    _lib.count(1 + (1 if want_grad else 0) + (1 if want_stats else 0) + (2 if want_counts else 0))
    if want_stats:
        return LossOut(fstat[0], fstat[1], istat[0], istat[1], dlow, pred, None, counts)
    return LossOut(None, None, None, None, dlow, pred, None, counts)


def _f32c(t, name):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError(f"{name} must be a contiguous float32 tensor")
    return t


def apgd_step(x, x_adv, x_old, grad, step, eps, a, out):
    """out <- one L-inf APGD update (robseg_apgd_step); bit-exact with attacker.py:388-410."""
    _need_cuda(x, x_adv, x_old, grad, step, out)
    lib = _lib.load()
    for n, t in (("x", x), ("x_adv", x_adv), ("x_old", x_old), ("grad", grad), ("step", step), ("out", out)):
        _f32c(t, n)
    B = x.shape[0]
    with torch.cuda.device(x.device), _timed("apgd_step", 20 * x.numel()):
        rc = lib.robseg_apgd_step(x.data_ptr(), x_adv.data_ptr(), x_old.data_ptr(), grad.data_ptr(),
                                  step.data_ptr(), float(eps), float(a), float(1.0 - a), B,
                                  x[0].numel(), out.data_ptr(), _stream())
    _lib.check(rc, "robseg_apgd_step")
    _lib.count(1)
    return out


def apgd_step_fused(x, x_adv, x_old, grad, step, eps, a, out, flags, x_best_adv, x_best, grad_best):
    """``apgd_step`` that first applies the previous iteration's flag-driven row copies
    (robseg_apgd_step_fused): x_best_adv / x_best / grad_best rows are stored, restarted rows of
    x_adv / grad are replaced in place.  flags: the [3,B] int32 tensor of ``apgd_bookkeep``."""
    _need_cuda(x, x_adv, x_old, grad, step, out, flags, x_best_adv, x_best, grad_best)
    lib = _lib.load()
    for n, t in (("x", x), ("x_adv", x_adv), ("x_old", x_old), ("grad", grad), ("step", step), ("out", out),
                 ("x_best_adv", x_best_adv), ("x_best", x_best), ("grad_best", grad_best)):
        _f32c(t, n)
    B = x.shape[0]
    if flags.dtype != torch.int32 or flags.shape != (3, B) or not flags.is_contiguous():
        raise TypeError("flags must be a contiguous int32 [3,B] tensor")
    with torch.cuda.device(x.device), _timed("apgd_step", 20 * x.numel()):
        rc = lib.robseg_apgd_step_fused(x.data_ptr(), x_adv.data_ptr(), x_old.data_ptr(), grad.data_ptr(),
                                        step.data_ptr(), float(eps), float(a), float(1.0 - a), B, x[0].numel(),
                                        out.data_ptr(), flags.data_ptr(), x_best_adv.data_ptr(),
                                        x_best.data_ptr(), grad_best.data_ptr(), _stream())
    _lib.check(rc, "robseg_apgd_step_fused")
    _lib.count(1)
    return out


def make_ctl(n_iter_max, device):
    """Device control block of the CUDA-graph iteration (robseg_b200.h ROBSEG_CTL_*)."""
    if n_iter_max > _lib.CTL_MAX_ITER:
        raise ValueError(f"n_iter {n_iter_max} exceeds ROBSEG_CTL_MAX_ITER")
    return torch.zeros(_lib.CTL_SCHED + n_iter_max, dtype=torch.int32, device=device)


def set_ctl(ctl, n_iter, eps, checks):
    """Start a stage: iteration 0, its length, eps and the check schedule {iteration: window k}
    (one small pinned-host -> device copy, stream ordered)."""
    import struct

    host = torch.zeros(ctl.numel(), dtype=torch.int32)
    host[_lib.CTL_NITER] = n_iter
    host[_lib.CTL_EPS] = struct.unpack("i", struct.pack("f", float(eps)))[0]
    for it, k in checks.items():
        host[_lib.CTL_SCHED + it] = k
    ctl.copy_(host.pin_memory(), non_blocking=True)
    return ctl


def apgd_step_ctl(x, x_adv, x_old, grad, step, ctl, flags, x_best_adv, x_best, grad_best):
    """In-place, device-controlled form of ``apgd_step_fused`` (robseg_apgd_step_ctl): x_old <- x_adv,
    x_adv <- new point; a / eps come from ``ctl``.  Fixed addresses, so it can sit in a CUDA graph."""
    lib = _lib.load()
    for n, t in (("x", x), ("x_adv", x_adv), ("x_old", x_old), ("grad", grad), ("step", step),
                 ("x_best_adv", x_best_adv), ("x_best", x_best), ("grad_best", grad_best)):
        _f32c(t, n)
    B = x.shape[0]
    with torch.cuda.device(x.device), _timed("apgd_step", 24 * x.numel()):
        rc = lib.robseg_apgd_step_ctl(x.data_ptr(), x_adv.data_ptr(), x_old.data_ptr(), grad.data_ptr(),
                                      step.data_ptr(), ctl.data_ptr(), B, x[0].numel(), flags.data_ptr(),
                                      x_best_adv.data_ptr(), x_best.data_ptr(), grad_best.data_ptr(), _stream())
    _lib.check(rc, "robseg_apgd_step_ctl")
    _lib.count(1)


def apgd_bookkeep_ctl(correct, valid, loss_indiv, acc, loss_best, loss_best_last, reduced_last, step,
                      loss_steps, ctl, HW, early_stop, flags, done):
    """``apgd_bookkeep`` with (iter, check_k, n_iter) read from -- and iter advanced in -- ``ctl``."""
    lib = _lib.load()
    B = acc.shape[0]
    with torch.cuda.device(acc.device), _timed("bookkeep", 0):
        rc = lib.robseg_apgd_bookkeep_ctl(
            correct.data_ptr(), valid.data_ptr(), loss_indiv.data_ptr(), acc.data_ptr(), loss_best.data_ptr(),
            loss_best_last.data_ptr(), reduced_last.data_ptr(), step.data_ptr(), loss_steps.data_ptr(),
            ctl.data_ptr(), B, int(HW), int(bool(early_stop)), flags.data_ptr(), done.data_ptr(), _stream())
    _lib.check(rc, "robseg_apgd_bookkeep_ctl")
    _lib.count(1)


def project_linf(z, x, eps, noise=None, out=None):
    """clip01(x + clip(z-x, +-eps)), or with noise: clip01(x + eps*noise) (robseg_project_linf)."""
    _need_cuda(z, x, noise, out)
    lib = _lib.load()
    _f32c(x, "x")
    if out is None:
        out = torch.empty_like(x)
    _f32c(out, "out")
    if z is not None:
        _f32c(z, "z")
    if noise is not None:
        _f32c(noise, "noise")
    with torch.cuda.device(x.device), _timed("project", 12 * x.numel()):
        rc = lib.robseg_project_linf(_ptr(z), x.data_ptr(), _ptr(noise), float(eps), x.numel(),
                                     out.data_ptr(), _stream())
    _lib.check(rc, "robseg_project_linf")
    _lib.count(1)
    return out


def pgd_step(X, delta, grad, alpha, eps, mask_outside=False, x_next=None, clamp_next=True):
    """In-place PIR-AT delta update (robseg_pgd_step), optionally emitting the next input."""
    _need_cuda(X, delta, grad, x_next)
    lib = _lib.load()
    for n, t in (("X", X), ("delta", delta), ("grad", grad)):
        _f32c(t, n)
    if x_next is not None:
        _f32c(x_next, "x_next")
    with torch.cuda.device(X.device), _timed("pgd_step", (16 + (4 if x_next is not None else 0)) * X.numel()):
        rc = lib.robseg_pgd_step(X.data_ptr(), delta.data_ptr(), grad.data_ptr(), float(alpha),
                                 float(eps), int(mask_outside), int(clamp_next), X.numel(),
                                 _ptr(x_next), _stream())
    _lib.check(rc, "robseg_pgd_step")
    _lib.count(1)
    return delta


def apgd_bookkeep(correct, valid, loss_indiv, acc, loss_best, loss_best_last, reduced_last, step,
                  loss_steps, it, check_k, HW, early_stop, flags, done):
    """Device-side step-size / best-point bookkeeping (robseg_apgd_bookkeep)."""
    lib = _lib.load()
    n_iter, B = loss_steps.shape
    with torch.cuda.device(acc.device), _timed("bookkeep", 0):
        rc = lib.robseg_apgd_bookkeep(
            correct.data_ptr(), valid.data_ptr(), loss_indiv.data_ptr(), acc.data_ptr(),
            loss_best.data_ptr(), loss_best_last.data_ptr(), reduced_last.data_ptr(),
            step.data_ptr(), loss_steps.data_ptr(), n_iter, int(it), int(check_k), B, int(HW),
            int(bool(early_stop)), flags.data_ptr(), done.data_ptr(), 0, _stream())
    _lib.check(rc, "robseg_apgd_bookkeep")
    _lib.count(1)
    return flags


def row_select(jobs, B, device):
    """jobs: list of (dst, src, flags[B] int32, unless[B] int32 | None); rows dst[b] <- src[b]."""
    lib = _lib.load()
    arr = (_lib.RowJob * len(jobs))()
    for k, (dst, src, flags, unless) in enumerate(jobs):
        if dst.shape != src.shape or dst.dtype != src.dtype:
            raise ValueError("row_select: dst/src mismatch")
        if not (dst.is_contiguous() and src.is_contiguous()):
            raise ValueError("row_select: tensors must be contiguous")
        arr[k] = _lib.RowJob(dst.data_ptr(), src.data_ptr(), flags.data_ptr(), _ptr(unless),
                             dst[0].numel() * dst.element_size())
    with torch.cuda.device(device), _timed("row_select", 0):
        rc = lib.robseg_row_select(arr, len(jobs), B, _stream())
    _lib.check(rc, "robseg_row_select")
    _lib.count(1)


def pixel_hist(pred, labels, n_cls, ignore_index=-1, want_hist=False, hist_total=None,
               want_counts=True):
    """Integer confusion / intersection / union counters (robseg_pixel_hist).

    pred [n_img,*spatial] int64, labels [n_lab,*spatial] int64 with n_img % n_lab == 0
    (image i is scored against labels[i % n_lab]).  Returns dict of int64 tensors:
    hist [n_img,C,C] (if want_hist), inter/tgt/prd [n_img,C] (if want_counts); hist_total
    [C,C] int64 is accumulated in place when given."""
    _need_cuda(pred, labels, hist_total)
    lib = _lib.load()
    pred = pred.detach()
    labels = labels.detach()
    if pred.dtype != torch.int64:
        pred = pred.long()
    if labels.dtype != torch.int64:
        labels = labels.long()
    pred, labels = pred.contiguous(), labels.contiguous()
    n_img, n_lab = pred.shape[0], labels.shape[0]
    HW = pred[0].numel()
    if labels[0].numel() != HW or n_img % n_lab != 0:
        raise ValueError("pred / labels shape mismatch")
    dev = pred.device
    out = {}
    if want_hist:
        out["hist"] = torch.zeros((n_img, n_cls, n_cls), dtype=torch.int64, device=dev)
    if want_counts:
        cnt = torch.zeros((3, n_img, n_cls), dtype=torch.int64, device=dev)
        out["inter"], out["tgt"], out["prd"] = cnt[0], cnt[1], cnt[2]
    if hist_total is not None and (hist_total.dtype != torch.int64 or not hist_total.is_contiguous()):
        raise TypeError("hist_total must be contiguous int64")
    with torch.cuda.device(dev), _timed("pixel_hist", 16 * n_img * HW):
        rc = lib.robseg_pixel_hist(pred.data_ptr(), labels.data_ptr(), n_img, n_lab, HW, int(n_cls),
                                   int(ignore_index), _ptr(out.get("hist")), _ptr(hist_total),
                                   _ptr(out.get("inter")), _ptr(out.get("tgt")), _ptr(out.get("prd")),
                                   _stream())
    _lib.check(rc, "robseg_pixel_hist")
    _lib.count(1)
    return out


def sea_worst_acc(inter, tgt):
    """inter/tgt [A,N,C] int64 -> (acc[A,N] f32, worst[N] f32) (robseg_sea_worst_acc)."""
    _need_cuda(inter, tgt)
    lib = _lib.load()
    inter, tgt = inter.contiguous(), tgt.contiguous()
    A, N, Cn = inter.shape
    acc = torch.empty((A, N), dtype=torch.float32, device=inter.device)
    worst = torch.empty((N,), dtype=torch.float32, device=inter.device)
    with torch.cuda.device(inter.device):
        rc = lib.robseg_sea_worst_acc(inter.data_ptr(), tgt.data_ptr(), A, N, Cn, acc.data_ptr(),
                                      worst.data_ptr(), _stream())
    _lib.check(rc, "robseg_sea_worst_acc")
    _lib.count(1)
    return acc, worst


def _upsample_fwd(x, H, W):
    _need_cuda(x)
    lib = _lib.load()
    if x.dtype != torch.float32 or x.dim() != 4:
        raise TypeError("upsample_bilinear expects a 4-D float32 CUDA tensor")
    x = x.contiguous()
    B, Cn, h, w = x.shape
    out = torch.empty((B, Cn, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _timed("upsample_fwd", 4 * (x.numel() + out.numel())):
        rc = lib.robseg_upsample_bilinear_fwd(x.data_ptr(), B * Cn, h, w, out.data_ptr(), H, W, _stream())
    _lib.check(rc, "robseg_upsample_bilinear_fwd")
    _lib.count(1)
    return out


def _upsample_bwd(g, h, w):
    _need_cuda(g)
    lib = _lib.load()
    if g.dtype != torch.float32:
        g = g.float()
    B, Cn, H, W = g.shape
    # a channel slice of a torch.cat gradient (rows contiguous, planes strided) is read in place
    if not (g.stride(3) == 1 and g.stride(2) == W and g.stride(1) >= H * W and g.stride(0) >= 0):
        g = g.contiguous()
    gin = torch.empty((B, Cn, h, w), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device), _timed("upsample_bwd", 4 * (g.numel() + gin.numel())):
        rc = lib.robseg_upsample_bilinear_bwd_strided(g.data_ptr(), B, Cn, g.stride(0), g.stride(1), H, W,
                                                      gin.data_ptr(), h, w, _stream())
    _lib.check(rc, "robseg_upsample_bilinear_bwd_strided")
    _lib.count(1)
    return gin


class _UpsampleBilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W):
        ctx.hw = (x.shape[2], x.shape[3])
        return _upsample_fwd(x, H, W)

    @staticmethod
    def backward(ctx, g):
        return _upsample_bwd(g, *ctx.hw), None, None


def upsample_bilinear(x, size):
    """Drop-in for ``F.interpolate(x, size=size, mode="bilinear", align_corners=False)`` on fp32
    NCHW CUDA tensors (robseg_upsample_bilinear_fwd / _bwd): streaming forward, deterministic
    gather backward.  SURVEY.md section 8f rank 1."""
    H, W = (size, size) if isinstance(size, int) else (int(size[0]), int(size[1]))
    return _UpsampleBilinear.apply(x, H, W)


_stock_interpolate = torch.nn.functional.interpolate


def interpolate(input, size=None, scale_factor=None, mode="nearest", align_corners=None, **kw):
    """``torch.nn.functional.interpolate`` with the bilinear / align_corners=False / fp32 / CUDA /
    4-D case routed to :func:`upsample_bilinear`; every other call goes to the stock function."""
    if (mode == "bilinear" and not align_corners and size is not None and scale_factor is None and not kw
            and input.is_cuda and input.dim() == 4 and input.dtype == torch.float32):
        return upsample_bilinear(input, size)
    return _stock_interpolate(input, size=size, scale_factor=scale_factor, mode=mode,
                              align_corners=align_corners, **kw)


class patched_interpolate:
    """Context manager: ``torch.nn.functional.interpolate`` -> :func:`interpolate` while a model
    that calls it by that name (the reference's UperNet head, semseg/models/uperforseg.py:193,
    282,297) runs its forward.  Restores the stock function on exit."""

    def __enter__(self):
        self.prev = torch.nn.functional.interpolate
        torch.nn.functional.interpolate = interpolate
        return self

    def __exit__(self, *a):
        torch.nn.functional.interpolate = self.prev


# ----------------------------------------------------------------------------------------------
# torch.ops.robseg.* : the same kernels as dispatcher-visible custom ops.  ``pixel_loss`` carries
# an autograd formula so the criterion_dict-compatible callables stay differentiable.
# ----------------------------------------------------------------------------------------------
_registered = False


def register_custom_ops():
    global _registered
    if _registered:
        return
    _registered = True
    lib = torch.library

    @lib.custom_op("robseg::pixel_loss", mutates_args=())
    def pixel_loss(logits: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, kind: str,
                   ignore_index: int) -> torch.Tensor:
        w = weights if weights.numel() else None
        return loss_fwd_bwd(logits, labels, kind, w, want_grad=False, want_loss_pix=True,
                            ignore_index=ignore_index, want_stats=False).loss_pix

    @pixel_loss.register_fake
    def _(logits, labels, weights, kind, ignore_index):
        return logits.new_empty((logits.shape[0], *logits.shape[2:]), dtype=torch.float32)

    @lib.custom_op("robseg::pixel_loss_bwd", mutates_args=())
    def pixel_loss_bwd(logits: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, kind: str,
                       ignore_index: int, gout: torch.Tensor) -> torch.Tensor:
        w = weights if weights.numel() else None
        return loss_fwd_bwd(logits, labels, kind, w, grad_scale=1.0, upstream=gout, want_grad=True,
                            ignore_index=ignore_index, want_stats=False).dlogits

    @pixel_loss_bwd.register_fake
    def _(logits, labels, weights, kind, ignore_index, gout):
        return torch.empty_like(logits)

    def _setup(ctx, inputs, output):
        logits, labels, weights, kind, ignore_index = inputs
        ctx.save_for_backward(logits, labels, weights)
        ctx.kind, ctx.ignore_index = kind, ignore_index

    def _backward(ctx, gout):
        logits, labels, weights = ctx.saved_tensors
        g = torch.ops.robseg.pixel_loss_bwd(logits, labels, weights, ctx.kind, ctx.ignore_index,
                                            gout.contiguous())
        return g, None, None, None, None

    pixel_loss.register_autograd(_backward, setup_context=_setup)

    @lib.custom_op("robseg::loss_fwd_bwd", mutates_args=())
    def loss_fwd_bwd_op(logits: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, kind: str,
                        ignore_index: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor,
                                                     torch.Tensor, torch.Tensor, torch.Tensor]:
        w = weights if weights.numel() else None
        o = loss_fwd_bwd(logits, labels, kind, w, want_grad=True, want_pred=True,
                         ignore_index=ignore_index)
        return o.dlogits, o.loss_img.clone(), o.track_img.clone(), o.correct.clone(), o.valid.clone(), o.pred

    @loss_fwd_bwd_op.register_fake
    def _(logits, labels, weights, kind, ignore_index):
        B = logits.shape[0]
        f = logits.new_empty((B,), dtype=torch.float32)
        i = logits.new_empty((B,), dtype=torch.int32)
        return (torch.empty_like(logits), f, f.clone(), i, i.clone(),
                logits.new_empty((B, *logits.shape[2:]), dtype=torch.int64))

    @lib.custom_op("robseg::apgd_step", mutates_args=())
    def apgd_step_op(x: torch.Tensor, x_adv: torch.Tensor, x_old: torch.Tensor, grad: torch.Tensor,
                     step: torch.Tensor, eps: float, a: float) -> torch.Tensor:
        return apgd_step(x, x_adv, x_old, grad, step, eps, a, torch.empty_like(x))

    @apgd_step_op.register_fake
    def _(x, x_adv, x_old, grad, step, eps, a):
        return torch.empty_like(x)

    @lib.custom_op("robseg::pixel_hist", mutates_args=())
    def pixel_hist_op(pred: torch.Tensor, labels: torch.Tensor, n_cls: int,
                      ignore_index: int) -> torch.Tensor:
        return pixel_hist(pred, labels, n_cls, ignore_index, want_hist=True, want_counts=False)["hist"]

    @pixel_hist_op.register_fake
    def _(pred, labels, n_cls, ignore_index):
        return pred.new_empty((pred.shape[0], n_cls, n_cls), dtype=torch.int64)


def pixel_loss(logits, labels, kind, weights=None, ignore_index=-1):
    """Differentiable per-pixel criterion [B,*spatial] (criterion_dict semantics)."""
    register_custom_ops()
    w = weights if weights is not None else logits.new_empty((0,), dtype=torch.float32)
    w = w.to(device=logits.device, dtype=torch.float32)
    if labels.dtype != torch.int64:
        labels = labels.long()
    return torch.ops.robseg.pixel_loss(logits.contiguous(), labels.contiguous(), w, kind, ignore_index)
