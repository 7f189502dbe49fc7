"""B200 drop-in for the reference's ``semseg/attacker.py`` (SEA losses + APGD, L-inf path).

Same public names, positional order, defaults and return tuples as the reference
(semseg/attacker.py:9-52, :143-257, :260-571, :662-728), so ``tools/infer.py`` and
``tools/train_rob_seg.py`` can bind to this module unchanged.  What differs is the
execution: per APGD iteration the attack-side work is

  K_step  robseg_apgd_step       one launch  (replaces ~14 elementwise ATen kernels, :388-410)
  model forward                  PyTorch (cuDNN/cuBLAS) -- the surrounding consumer
  K_loss  robseg_loss_fwd_bwd    one pass: softmax, loss, track CE, argmax, accuracy AND
                                 d(loss)/d(logits)      (replaces ~7-25 passes, :143-240)
  model backward                 torch.autograd.grad(logits, x, grad_outputs=dlogits)
  K_book  robseg_apgd_bookkeep   device-side flags: no .nonzero()/.sum() host syncs (:485-551)
  (the boolean-index row copies those flags select ride in the NEXT K_step launch --
   robseg_apgd_step_fused -- and one robseg_row_select flushes them after the last iteration)

The only data-dependent host decision left is ``early_stop`` (:568-569); it is read one
iteration late from a pinned flag while the device freezes the state itself, so results
do not depend on when the host notices.

Out of scope: L2 / L1 norms and pgd_filters -- never reached from
tools/infer.py (norm is hard-set to "Linf" at tools/infer.py:327); they raise
NotImplementedError.  ``dlr_loss`` / ``dlr_loss_targeted`` / ``margin_loss`` are kept for name
parity as plain differentiable PyTorch expressions (no kernel: no driver selects them), and
``apgd_restarts`` (untargeted) runs on top of the accelerated ``apgd_train``.
"""
from functools import partial

import torch

from .. import ops

__all__ = ["compute_iou_acc", "masked_cross_entropy", "masked_cross_entropy_balanced",
           "js_div_fn", "js_loss", "pixel_to_img_loss", "check_oscillation", "criterion_dict",
           "apgd_train", "apgd_largereps", "apgd_restarts", "apgd_schedule", "dlr_loss",
           "dlr_loss_targeted", "margin_loss"]


class Logger:
    """Minimal stand-in for autoattack.other_utils.Logger (semseg/attacker.py:6,592,692)."""

    def __init__(self, log_path=None):
        self.log_path = log_path

    def log(self, str_to_log):
        print(str_to_log)
        if self.log_path is not None:
            with open(self.log_path, "a") as f:
                f.write(str_to_log + "\n")
                f.flush()


# ---------------------------------------------------------------------------------- metrics
def _iou_acc_from_counts(inter, tgt, prd):
    """fp32 finaliser of compute_iou_acc (semseg/attacker.py:29-50) on exact integer counts."""
    a, n = inter.float(), tgt.float()
    u = n + prd.float() - a
    ind = n > 0
    m_acc = (a[ind] / n[ind]).mean().cpu()
    a_acc = (a.sum() / n.sum()).cpu()
    ind = u > 0
    m_iou = (a[ind] / u[ind]).mean().cpu()
    return m_acc, a_acc, m_iou


def compute_iou_acc(pred, target, n_cls, verbose=False, ignore_index=-1, device=None):
    """mAcc / aAcc / mIoU of a batch of predictions (semseg/attacker.py:9-52).

    One ``robseg_pixel_hist`` launch instead of 2*n_cls masked reductions.  Like the
    reference, ``pred`` is modified in place (ignored pixels are set to ``ignore_index``,
    :20).  Returns three 0-dim CPU tensors."""
    pred[target == ignore_index] = ignore_index
    cnt = ops.pixel_hist(pred, target, n_cls, ignore_index)
    m_acc, a_acc, m_iou = _iou_acc_from_counts(cnt["inter"].sum(0), cnt["tgt"].sum(0),
                                               cnt["prd"].sum(0))
    if verbose:
        print(f"mAcc={m_acc:.2%} aAcc={a_acc:.2%}", f" mIoU={m_iou:.2%}")
    return m_acc, a_acc, m_iou


# ----------------------------------------------------------------------------------- losses
def _reduce(loss, pred, reduction):
    if reduction == "mean":
        return loss.view(pred.shape[0], -1).mean(-1)
    return loss


def masked_cross_entropy(pred, target, weights=None, reduction="none", ignore_index=-1):
    """Cross-entropy of only correctly classified pixels (semseg/attacker.py:143-152)."""
    return _reduce(ops.pixel_loss(pred, target, "mask-ce-avg", None, ignore_index), pred, reduction)


def masked_cross_entropy_balanced(pred, target, weights=None, reduction="none", ignore_index=-1):
    """Class-balanced variant (semseg/attacker.py:155-173); ``weights=None`` = unweighted."""
    return _reduce(ops.pixel_loss(pred, target, "mask-ce-bal", weights, ignore_index), pred, reduction)


def js_div_fn(p, q, weights=None, softmax_output=False, reduction="none", red_dim=None,
              ignore_index=-1):
    """JS divergence between softmax(p) and one-hot(q) (semseg/attacker.py:187-226).

    Only the form the attack uses is fused: logits in, ``reduction="none"``, summed over the
    class axis (``red_dim=1``) -- closed form of SURVEY.md section 10, finite where the
    reference produces NaN (section 9-Q15)."""
    if softmax_output or reduction != "none" or red_dim not in (1, (1,)):
        raise NotImplementedError("only js_div_fn(logits, labels, red_dim=1) is accelerated")
    return ops.pixel_loss(p, q, "js-avg", None, ignore_index)


def js_loss(p, q, num_classes=21, reduction="mean"):
    """semseg/attacker.py:229-234."""
    loss = js_div_fn(p, q, red_dim=(1))
    if reduction == "mean":
        return loss.view(p.shape[0], -1).mean(-1)
    elif reduction == "none":
        return loss


# Name parity only (semseg/attacker.py:123-141,176-184): the three margin-type losses of the reference's
# module.  No SEA / PIR-AT driver selects them, so they stay stock PyTorch on whatever device the logits
# live on; written from the published definitions (Croce & Hein 2020, DLR), class axis = 1.
def _top_logits(x, k):
    return x.topk(k, dim=1).values  # descending along the class axis


def dlr_loss(x, y, reduction="none"):
    """Difference-of-logits-ratio: -(z_y - max_{k != y} z_k) / (z_(1) - z_(3) + 1e-12)."""
    top = _top_logits(x, 3)
    z_y = x.gather(1, y.unsqueeze(1)).squeeze(1)
    y_is_top = (x.argmax(1) == y).to(x.dtype)
    best_other = top[:, 1] * y_is_top + top[:, 0] * (1.0 - y_is_top)
    return -(z_y - best_other) / (top[:, 0] - top[:, 2] + 1e-12)


def dlr_loss_targeted(x, y, y_target):
    """Targeted DLR: -(z_y - z_t) / (z_(1) - (z_(3) + z_(4)) / 2 + 1e-12)."""
    top = _top_logits(x, 4)
    z_y = x.gather(1, y.unsqueeze(1)).squeeze(1)
    z_t = x.gather(1, y_target.unsqueeze(1)).squeeze(1)
    return -(z_y - z_t) / (top[:, 0] - 0.5 * (top[:, 2] + top[:, 3]) + 1e-12)


def margin_loss(pred, target):
    """max_{k != y} z_k - z_y per pixel (the label channel pushed down by 1e10 before the max)."""
    onehot = torch.zeros_like(pred).scatter_(1, target.unsqueeze(1), 1.0)
    return (pred - 1e10 * onehot).amax(1) - (pred * onehot).sum(1)


def pixel_to_img_loss(loss, mask_background=None):
    """semseg/attacker.py:237-240."""
    if mask_background is not None:
        loss = mask_background * loss
    return loss.view(loss.shape[0], -1).mean(-1)


def check_oscillation(x, j, k, y5, k3=0.75):
    """semseg/attacker.py:243-248 (kept for API parity; the attack uses the device version)."""
    rows = [(j - c) % x.shape[0] for c in range(k + 1)]       # negative indices wrap
    ups = (x[rows[:-1]] > x[rows[1:]]).float().sum(0)           # increases inside the window
    return (ups <= k * k3).float()


def _ce(x, y, weights=None):  # accepts the third argument the caller passes
    return ops.pixel_loss(x, y, "ce", None, -1)


criterion_dict = {
    "ce": _ce,
    "ce-avg": _ce,
    "mask-ce-avg": masked_cross_entropy,
    "mask-ce-bal": masked_cross_entropy_balanced,
    "js-avg": partial(js_loss, reduction="none"),
}


# ------------------------------------------------------------------------------------- APGD
def apgd_schedule(n_iter):
    """Iterations at which the step-size check fires and its window k.  Depends only on
    n_iter (semseg/attacker.py:322-329,528-551), so the host never has to look at data."""
    n_iter_2 = max(int(0.22 * n_iter), 1)
    n_iter_min = max(int(0.06 * n_iter), 1)
    size_decr = max(int(0.03 * n_iter), 1)
    k, counter3, checks = n_iter_2, 0, {}
    for i in range(n_iter):
        counter3 += 1
        if counter3 == k:
            checks[i] = k
            counter3 = 0
            k = max(k - size_decr, n_iter_min)
    return checks


def apgd_train(model, x, y, norm, eps, n_iter=10, use_rs=False, loss="ce", verbose=False,
               is_train=False, early_stop=False, track_loss=None, logger=None, y_target=None,
               ignore_index=-1, x_init=None, num_classes=21, weights=None, return_pred=False,
               return_counts=False, gpuu=None):
    """APGD (L-inf) with the SEA losses; returns ``(x_best, acc, loss_best, x_best_adv)``.

    ``gpuu`` is accepted and ignored: the reference trainer's APGD branch passes it
    (tools/train_rob_seg.py:303-315) although the reference's own ``apgd_train`` does not take it
    (that branch cannot run there; here it does, on the tensors' own device).

    ``return_pred=True`` (extension) appends ``pred_best``: the argmax map of
    ``x_best_adv`` as seen during the attack, which lets the SEA driver skip the re-forward of
    every adversarial batch (tools/infer.py:82-90).  ``return_counts=True`` appends ``counts_best``
    [B,3,C] int64: the per-image intersection / target / prediction class counters of that same
    point (what eval_performance / evalSEA derive from the prediction map), taken inside the loss
    kernel's argmax pass -- no prediction map, no histogram launch.

    Mirrors semseg/attacker.py:260-571 step for step; see the module docstring for the
    kernel each block of the reference maps to."""
    assert not model.training
    assert ignore_index == -1, "Only `ignore_index = 1` is supported."
    if norm != "Linf":
        raise NotImplementedError("only the L-inf path is implemented (SURVEY.md section 2)")
    if loss not in criterion_dict:
        raise KeyError(loss)
    if track_loss is None:
        track_loss = loss
    if track_loss not in criterion_dict:
        raise KeyError(track_loss)
    if not x.is_cuda:
        raise RuntimeError("robseg-b200 apgd_train needs CUDA tensors (no CPU fallback)")
    device = x.device
    x = x.detach().float().contiguous()
    y = y.to(device)
    bs = x.shape[0]
    n_pxl = x.shape[-2] * x.shape[-1]
    eps = float(eps)

    # random start: the RNG draw stays on the torch side and happens even if x_init
    # overrides it (attacker.py:292-297)
    if not use_rs:
        x_adv = x.clone()
    else:
        t = 2 * torch.rand_like(x) - 1
        x_adv = ops.project_linf(None, x, eps, noise=t)
    if x_init is not None:
        x_adv = x_init.detach().float().contiguous().clone()
    if verbose and logger is None:
        logger = Logger(None)
    if logger is not None:  # attacker.py:302-305 (one host read per call, labels only)
        n_bg = int((y == ignore_index).sum())
        if n_bg > 0:
            logger.log(f"{n_bg / y.numel():.2%} pixels are masked out.")
    x_adv = x_adv.clamp_(0.0, 1.0)

    w_dev = None
    if weights is not None and loss == "mask-ce-bal":
        w_dev = weights.to(device=device, dtype=torch.float32)  # hoisted H2D

    keep_pred = return_pred
    keep_counts = verbose or return_counts

    def result(x_best, acc, loss_best, x_best_adv, pred_best, counts_best):
        return (x_best, acc, loss_best, x_best_adv) + ((pred_best,) if return_pred else ()) + (
            (counts_best,) if return_counts else ())

    graphed = hasattr(model, "input_grad")  # graphs.GraphedModel: replayed forward / input gradient
    if graphed and hasattr(model, "attack") and not verbose and track_loss in ("ce", "ce-avg", loss):
        # the whole iteration as one CUDA graph (graphs.GraphedAttack)
        return result(*model.attack(keep_pred, early_stop, n_iter).run_stage(
            x, y, x_adv, eps, n_iter, loss, track_loss, w_dev, apgd_schedule(n_iter)))
    # dropin.accelerate(model, fuse_loss=True): the model hands out its logits BEFORE the final bilinear
    # up-sampling and the loss kernel interpolates on the fly (robseg_loss_upsampled_fwd_bwd)
    lowres = getattr(model, "forward_lowres", None) if not graphed else None

    def forward_backward(x_in_buf, need_grad, dbuf):
        fused = False
        if graphed:
            logits = model(x_in_buf)
        else:
            x_in = x_in_buf.detach().requires_grad_(need_grad)
            with torch.set_grad_enabled(need_grad):
                logits = lowres(x_in) if lowres is not None else None
                fused = logits is not None and ops.can_fuse_upsample(logits, y)
                if not fused:
                    logits = model(x_in)
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()  # fp16 under autocast: F.cross_entropy up-casts too (attacker.py:147)
        if fused:
            def run(kind, **kw):
                return ops.loss_upsampled_fwd_bwd(logits, y, kind, w_dev, **kw)
            out = run(loss, want_grad=need_grad, want_pred=keep_pred, want_counts=keep_counts, dlow_out=dbuf)
        else:
            def run(kind, **kw):
                return ops.loss_fwd_bwd(logits, y, kind, w_dev, **kw)
            out = run(loss, want_grad=need_grad, want_pred=keep_pred, want_counts=keep_counts, dlogits_out=dbuf)
        g = None
        if need_grad:
            if graphed:
                g = model.input_grad(out.dlogits).clone()
            else:
                (g,) = torch.autograd.grad(logits, [x_in], grad_outputs=out.dlogits)
        # per-image value driving best-loss / step-size decisions (attacker.py:353-361,473-475)
        if track_loss in ("ce", "ce-avg"):
            track = out.track_img
        elif track_loss == loss:
            track = out.loss_img
        else:
            track = run(track_loss, want_grad=False).loss_img
        return out, g, track, logits

    # ---- initial point (attacker.py:342-383) ------------------------------------------------
    # (a graphed model hands out its static gradient buffer: the loss kernel writes into it)
    out, grad, track, logits = forward_backward(x_adv, True, model.gout if graphed else None)
    dbuf = out.dlogits  # reused every iteration: the model backward has consumed it by then
    n_cls = logits.shape[1]
    del logits
    grad = grad.contiguous()
    # ignored pixels count as wrong here (:370-371).  Tensor / tensor is an IEEE division on
    # the device (tensor / python-scalar would multiply by a rounded reciprocal), matching
    # the bookkeeping kernel and the reference's CPU mean.
    acc = out.correct.float() / torch.full((), float(n_pxl), device=device)
    loss_best = track.clone()
    loss_best_last = loss_best.clone()
    reduced_last = torch.ones_like(loss_best)
    step = 2.0 * eps * torch.ones([bs], device=device)
    loss_steps = torch.zeros([n_iter, bs], device=device)
    x_best = x_adv.clone()
    x_best_adv = x_adv.clone()
    grad_best = grad.clone()
    x_old = x_adv.clone()
    x_new = torch.empty_like(x_adv)
    pred_best = out.pred if keep_pred else None
    counts_best = out.counts if keep_counts else None
    flags = torch.zeros([3, bs], dtype=torch.int32, device=device)
    done = torch.zeros([1], dtype=torch.int32, device=device)
    done_host = torch.zeros([1], dtype=torch.int32).pin_memory() if early_stop else None
    copied = None
    checks = apgd_schedule(n_iter)

    pending = False  # row copies of the previous iteration, folded into the next step launch
    for i in range(n_iter):
        # ---- gradient step (attacker.py:388-410) ---------------------------------------------
        a = 0.75 if i > 0 else 1.0
        if pending:
            # + the previous iteration's x_best_adv / x_best / grad_best row stores and restart rows
            # (:494-495,523-525,546-548): the step reads x_adv and grad anyway
            ops.apgd_step_fused(x, x_adv, x_old, grad, step, eps, a, x_new, flags, x_best_adv, x_best, grad_best)
        else:
            ops.apgd_step(x, x_adv, x_old, grad, step, eps, a, x_new)
        x_old, x_adv, x_new = x_adv, x_new, x_old

        # ---- forward, fused loss, backward (attacker.py:459-475) -----------------------------
        need_grad = i < n_iter - 1  # the reference saves the last backward pass (:467-469)
        out, g, track, _ = forward_backward(x_adv, need_grad, dbuf)
        if g is not None:
            grad = g.contiguous()

        # ---- accuracy / best-point / step-size bookkeeping (attacker.py:485-551) -------------
        ops.apgd_bookkeep(out.correct, out.valid, track, acc, loss_best, loss_best_last,
                          reduced_last, step, loss_steps, i, checks.get(i, 0), n_pxl, early_stop,
                          flags, done)
        pending = True
        if keep_pred or keep_counts:
            ops.row_select(([(pred_best, out.pred, flags[0], None)] if keep_pred else []) +
                           ([(counts_best, out.counts, flags[0], None)] if keep_counts else []), bs, device)

        if verbose:
            # compute_iou_acc of the best point's prediction (:496-498) from its fused counters
            m_acc, a_acc, m_iou = _iou_acc_from_counts(counts_best[:, 0].sum(0), counts_best[:, 1].sum(0),
                                                       counts_best[:, 2].sum(0))
            logger.log(
                "iteration: {} - best loss: {:.6f} curr loss {:.6f} - mAcc={:.2%} aAcc={:.2%} "
                "mIoU={:.2%} - step size: {:.5f}".format(i, loss_best.sum(), track.sum(), m_acc,
                                                         a_acc, m_iou, step.mean()))

        if early_stop:
            # poll the flag of the PREVIOUS iteration: the device freezes the state itself once
            # acc.sum()==0 (:568-569), so running one iteration late cannot change the result.
            if copied is not None:
                copied.synchronize()
                if int(done_host[0]) != 0:
                    break
            done_host.copy_(done, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record()

    if pending:  # the last iteration's row stores (restart copies only matter to a following step)
        ops.row_select([(x_best_adv, x_adv, flags[0], None), (x_best, x_adv, flags[1], None),
                        (grad_best, grad, flags[1], None)], bs, device)
    return result(x_best, acc, loss_best, x_best_adv, pred_best, counts_best)


def apgd_largereps(model, x, y, weights, norm="Linf", eps=8.0 / 255.0, n_iter=10, loss="ce",
                   verbose=False, n_restarts=1, log_path=None, early_stop=False, eot_iter=0,
                   track_loss=None, use_rs=False, ignore_index=-1, num_classes=21, return_pred=False,
                   return_counts=False):
    """SEA's 3-stage large-eps schedule (semseg/attacker.py:662-728): iterations
    ``[.3n, .3n, rest]`` at ``[2 eps, 1.5 eps, eps]``, each stage started from the projection of
    the previous stage's lowest-accuracy point.  Returns ``(x_adv, loss_best, acc)``
    (+ ``pred`` of ``x_adv`` with ``return_pred=True``; + its per-image class
    counters [B,3,C] with ``return_counts=True``, see :func:`apgd_train`)."""
    if norm != "Linf":
        raise NotImplementedError()
    logger = Logger(log_path)
    n_iters = [int(c * n_iter) for c in [0.3, 0.3]]
    n_iters.append(n_iter - sum(n_iters))
    epss = [c * eps for c in [2, 1.5, 1]]

    acc = torch.ones([x.shape[0]], device=x.device)
    x = x.detach().float().contiguous()
    x_init = None
    loss_best = pred = counts = None
    last = len(n_iters) - 1
    for stage, (inner_it, inner_eps) in enumerate(zip(n_iters, epss)):
        if x_init is not None:
            x_init = ops.project_linf(x_init, x, inner_eps)
        res = apgd_train(
            model, x, y, n_iter=inner_it, use_rs=use_rs, verbose=verbose, loss=loss,
            eps=inner_eps, norm=norm, logger=logger, early_stop=early_stop,
            track_loss=track_loss, y_target=None, ignore_index=ignore_index, x_init=x_init,
            num_classes=num_classes, weights=weights, return_pred=return_pred and stage == last,
            return_counts=return_counts and stage == last)
        _, acc, loss_best, x_init = res[:4]
        if stage == last:
            extra = list(res[4:])
            pred = extra.pop(0) if return_pred else None
            counts = extra.pop(0) if return_counts else None
    return (x_init, loss_best, acc) + ((pred,) if return_pred else ()) + ((counts,) if return_counts else ())


def apgd_restarts(model, x, y, norm="Linf", eps=8.0 / 255.0, n_iter=10, loss="ce", verbose=False,
                  n_restarts=1, log_path=None, early_stop=False, eot_iter=0, track_loss=None,
                  use_rs=False, ignore_index=-1):
    """APGD with restarts (semseg/attacker.py:574-659): ``n_restarts`` independent ``apgd_train`` runs
    over the images whose accuracy is still positive, keeping per image the adversarial point with the
    lowest pixel accuracy (ignored pixels count as correct, :639).  Returns ``(x_adv, None, acc)``.
    The accuracy of each run's point comes from the argmax map the attack already holds
    (``return_pred``) instead of the reference's extra forward.  Untargeted losses only."""
    if "targeted" in loss:
        raise NotImplementedError("targeted losses are outside the accelerated path")
    logger = Logger(log_path)
    x = x.detach().float().contiguous()
    acc = torch.ones([x.shape[0]], device=x.device)
    x_adv = x.clone()
    for i in range(n_restarts):
        rows = (acc > 0).nonzero().flatten()
        if rows.numel() == 0:
            break
        yr = y[rows]
        res = apgd_train(model, x[rows], yr, n_iter=n_iter, use_rs=use_rs, verbose=verbose, loss=loss, eps=eps,
                         norm=norm, logger=logger, early_stop=early_stop, track_loss=track_loss, y_target=None,
                         ignore_index=ignore_index, return_pred=True)
        x_cur, pred = res[3], res[4]
        ok = (pred == yr) | (yr == ignore_index)
        acc_cur = ok.float().view(rows.numel(), -1).mean(-1)
        better = acc_cur < acc[rows]
        x_adv[rows[better]] = x_cur[better]
        acc[rows[better]] = acc_cur[better]
        note = " (warning: this is only upper bound on aAcc)" if bool((yr == ignore_index).any()) else ""
        logger.log(f"restart {i + 1} robust accuracy={acc.float().mean():.1%}{note}")
    return x_adv, None, acc


def L1_projection(*args, **kwargs):
    raise NotImplementedError("L1 attacks are outside the accelerated path")


def pgd_filters(*args, **kwargs):
    raise NotImplementedError("pgd_filters is outside the accelerated path")

