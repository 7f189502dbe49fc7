"""B200 drop-in for the hot-path functions of the reference's ``tools/infer.py``:
``eval_performance`` (:56-133), ``evaluate`` (:136-155) and ``check_imgs`` (:39-53).

Adversarial-example bookkeeping (north-star piece 5).  The reference moves every adversarial
batch to the CPU (:151, 50 MB per batch at B=16), later moves it back and re-forwards it
(:82-84), then runs 2*C masked reductions per batch on the CPU (:94-116).  Here:

* ``evaluate`` keeps the same signature and return value (a list of ``(x_adv, target)`` CPU
  pairs) but the D2H copies are asynchronous into pinned memory, overlapped with the next
  attack; ``keep_on_device=True`` skips the round trip entirely;
* ``eval_performance`` takes the argmax with the fused ARGMAX kernel and accumulates exact
  int64 intersection / target / prediction counts with ``robseg_pixel_hist``; the final
  mAcc / aAcc / mIoU replay the reference's float32 finaliser.  The returned ``l_output`` is
  the ``[N,H,W]`` int64 argmax store with -1 at ignored pixels, as ``evalSEA`` expects.
"""
import functools

import torch

from .. import ops
from ..semseg import attacker as _attacker


class AdvLoader(list):
    """What ``evaluate`` returns: the reference's list of ``(x_adv, target)`` pairs, plus -- when the
    attack is this package's ``apgd_largereps`` -- the argmax map of every adversarial batch as
    the attack saw it (``return_pred``) and the device copy of its targets.
    ``eval_performance(model, adv_loader)`` then needs neither the H2D copy of ``x_adv`` nor the
    re-forward of tools/infer.py:82-84; the pairs themselves stay as the reference defines them."""

    def __init__(self, model=None):
        super().__init__()
        self.model = model
        self.device_items = []  # per batch (pred [B,H,W] int64, target [B,H,W]) on the device


def _returns_pred(attack_fn):
    f = attack_fn.func if isinstance(attack_fn, functools.partial) else attack_fn
    return f is _attacker.apgd_largereps


def check_imgs(adv, x, norm, verbose=False):
    """Perturbation-size / range report of an adversarial batch (tools/infer.py:39-53)."""
    flat = (adv - x).flatten(1)
    size = {"Linf": lambda d: d.abs().amax(1), "L2": lambda d: d.square().sum(1).sqrt(),
            "L1": lambda d: d.abs().sum(1)}[norm](flat)
    report = (f"max {norm} pert: {size.max():.5f}, nan in imgs: {adv.isnan().sum()}, "
              f"max in imgs: {adv.max():.5f}, min in imgs: {adv.min():.5f}")
    if verbose:
        print(report)
    return report


def eval_performance(model, data_loader, n_batches=-1, n_cls=21, return_output=False,
                     ignore_index=-1, return_preds=False, verbose=False, device="cuda"):
    """Accuracy and mIoU of ``model`` over ``data_loader`` (tools/infer.py:56-133).
    Returns ``(stats, l_output)`` with ``stats = {"mAcc","aAcc","mIoU"}``."""
    model.eval()
    dev = torch.device(device)
    inter = torch.zeros(n_cls, dtype=torch.int64, device=dev)
    tgt, prd = torch.zeros_like(inter), torch.zeros_like(inter)
    l_output = []
    cached = None
    if isinstance(data_loader, AdvLoader) and data_loader.model is model and \
            len(data_loader.device_items) == len(data_loader):
        cached = data_loader.device_items
    for i, vals in enumerate(data_loader):
        if cached is not None:  # argmax maps handed over by the attack: no re-forward (8f-2)
            pred, target = cached[i][0].clone(), cached[i][1]
        else:
            input, target = vals[0].to(dev, non_blocking=True), vals[1].to(dev, non_blocking=True)
            with torch.no_grad():
                output = model(input)
            if output.dtype not in (torch.float32, torch.bfloat16):
                output = output.float()
            pred = ops.loss_fwd_bwd(output, target, "argmax", want_grad=False, want_pred=True,
                                    ignore_index=ignore_index, want_stats=False).pred
        pred[target == ignore_index] = ignore_index  # :90
        c = ops.pixel_hist(pred, target, n_cls, ignore_index)
        inter += c["inter"].sum(0)
        tgt += c["tgt"].sum(0)
        prd += c["prd"].sum(0)
        l_output.append(pred.cpu())
        if verbose:
            s = _finalize(inter, tgt, prd)
            print(f"batch={i} running mAcc={s['mAcc']:.2%} running aAcc={s['aAcc']:.2%}",
                  f" running mIoU={s['mIoU']:.2%}")
        if i + 1 == n_batches:
            print("enough batches seen")
            break
    return _finalize(inter, tgt, prd), torch.cat(l_output)


def _finalize(inter, tgt, prd):
    """float32 finaliser of tools/infer.py:99-116 on the exact running counts."""
    a, n = inter.float().cpu(), tgt.float().cpu()
    u = n + prd.float().cpu() - a
    ind = n > 0
    m_acc = (a[ind] / n[ind]).mean()
    a_acc = a.sum() / n.sum()
    ind = u > 0
    m_iou = (a[ind] / u[ind]).mean()
    return {"mAcc": m_acc.item(), "aAcc": a_acc.item(), "mIoU": m_iou.item()}


def evaluate(val_loader, model, attack_fn, n_batches=-1, args=None, weights=None,
             keep_on_device=False, device="cuda"):
    """Run ``attack_fn`` on every batch (tools/infer.py:136-155); returns the adversarial
    "loader": a list of ``(x_adv, target)`` pairs (CPU tensors unless keep_on_device)."""
    model.eval()
    dev = torch.device(device)
    adv_loader, pending = AdvLoader(model), None
    with_pred = _returns_pred(attack_fn)
    for i, (input, target, _) in enumerate(val_loader):
        input = input.to(dev, non_blocking=True)
        target = target.to(dev, non_blocking=True)
        if with_pred:
            x_adv, _, acc, pred = attack_fn(model, input.clone(), target, weights, return_pred=True)
            adv_loader.device_items.append((pred, target))
        else:
            x_adv, _, acc = attack_fn(model, input.clone(), target, weights)
        if args is not None and getattr(args, "norm", None):
            check_imgs(input, x_adv, norm=args.norm)
        if keep_on_device:
            adv_loader.append((x_adv, target))
        else:
            x_host = torch.empty(x_adv.shape, dtype=x_adv.dtype).pin_memory()
            t_host = torch.empty(target.shape, dtype=target.dtype).pin_memory()
            x_host.copy_(x_adv, non_blocking=True)
            t_host.copy_(target, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending.synchronize()
            pending = ev
            adv_loader.append((x_host, t_host))
        if i + 1 == n_batches:
            break
    if pending is not None:
        pending.synchronize()
    return adv_loader
