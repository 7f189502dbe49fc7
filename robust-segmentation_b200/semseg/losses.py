"""B200 drop-in for the reference's ``semseg/losses.py`` (semseg/losses.py:6-109).

``CrossEntropy`` (mean reduction, ignore label, optional class weights, aux-head weighting)
runs on the fused loss kernel: forward = one pass producing the per-pixel weighted CE, backward
= one pass producing d/dlogits (``torch.ops.robseg.pixel_loss`` carries the autograd formula).
``OhemCrossEntropy`` and ``Dice`` are not on the SEA / PIR-AT path;
they are kept as plain-PyTorch bodies for name compatibility only.
"""
import math

import torch
from torch import Tensor, nn
from torch.nn import functional as F

from .. import ops


class CrossEntropy(nn.Module):
    def __init__(self, ignore_label: int = 255, weight: Tensor = None,
                 aux_weights: list = [1, 0.4, 0.4]) -> None:
        super().__init__()
        self.aux_weights = aux_weights
        self.ignore_label = ignore_label
        self.weight = weight

    def _forward(self, preds: Tensor, labels: Tensor) -> Tensor:
        # nn.CrossEntropyLoss(weight, ignore_index) with mean reduction:
        #   sum_p w_y * ce_p / sum_p w_y  over non-ignored pixels
        if self.weight is None:
            lp = ops.pixel_loss(preds, labels, "ce", None, self.ignore_label)
            denom = (labels != self.ignore_label).sum().clamp(min=1).to(lp.dtype)
            return lp.sum() / denom
        w = self.weight.to(preds.device, torch.float32)
        ce = ops.pixel_loss(preds, labels, "ce", None, self.ignore_label)
        keep = labels != self.ignore_label
        wy = w[labels.clamp(min=0, max=w.numel() - 1)] * keep
        return (wy * ce).sum() / wy.sum()

    def forward(self, preds, labels: Tensor) -> Tensor:
        return _weighted_sum(self._forward, preds, labels, self.aux_weights)


def _weighted_sum(fn, preds, labels, aux_weights):
    """Main + auxiliary heads: preds may be a tuple (main, aux...) weighted by aux_weights."""
    if isinstance(preds, tuple):
        return sum(w * fn(p, labels) for p, w in zip(preds, aux_weights))
    return fn(preds, labels)


class OhemCrossEntropy(nn.Module):
    """Online hard example mining CE (semseg/losses.py:30-63).  Not on the SEA / PIR-AT path; the
    per-pixel CE comes from the fused kernel, the hard-example selection stays in PyTorch."""

    def __init__(self, ignore_label: int = 255, weight: Tensor = None, thresh: float = 0.7,
                 aux_weights: list = [1, 1]) -> None:
        super().__init__()
        self.ignore_label, self.aux_weights, self.weight = ignore_label, aux_weights, weight
        self.thresh = -math.log(thresh)

    def _forward(self, preds: Tensor, labels: Tensor) -> Tensor:
        keep = labels != self.ignore_label
        ce = ops.pixel_loss(preds, labels, "ce", None, self.ignore_label)
        if self.weight is not None:
            w = self.weight.to(preds.device, torch.float32)
            ce = ce * w[labels.clamp(0, w.numel() - 1)] * keep
        ce = ce.flatten()
        n_min = int(keep.sum()) // 16
        hard = ce[ce > self.thresh]
        if hard.numel() < n_min:
            hard = ce.topk(n_min).values
        return hard.mean()

    def forward(self, preds, labels: Tensor) -> Tensor:
        return _weighted_sum(self._forward, preds, labels, self.aux_weights)


class Dice(nn.Module):
    """Dice / Tversky loss (semseg/losses.py:66-93; delta weighs FN against FP).  Not on the
    SEA / PIR-AT path: plain PyTorch."""

    def __init__(self, delta: float = 0.5, aux_weights: list = [1, 0.4, 0.4]):
        super().__init__()
        self.delta, self.aux_weights = delta, aux_weights

    def _forward(self, preds: Tensor, labels: Tensor) -> Tensor:
        n = preds.shape[1]
        onehot = F.one_hot(labels, n).movedim(-1, 1).to(preds.dtype)
        dims = tuple(range(2, preds.dim()))
        tp = (onehot * preds).sum(dims)
        fn = (onehot * (1 - preds)).sum(dims)
        fp = ((1 - onehot) * preds).sum(dims)
        score = (tp + 1e-6) / (tp + self.delta * fn + (1 - self.delta) * fp + 1e-6)
        return ((1 - score).sum(-1) / n).mean()

    def forward(self, preds, targets: Tensor) -> Tensor:
        return _weighted_sum(self._forward, preds, targets, self.aux_weights)


__all__ = ["CrossEntropy", "OhemCrossEntropy", "Dice"]


def get_loss(loss_fn_name: str = "CrossEntropy", ignore_label: int = 255, cls_weights: Tensor = None):
    assert loss_fn_name in __all__, (
        f"Unavailable loss function name >> {loss_fn_name}.\nAvailable loss functions: {__all__}")
    if loss_fn_name == "Dice":
        return Dice()
    return {"CrossEntropy": CrossEntropy, "OhemCrossEntropy": OhemCrossEntropy}[loss_fn_name](
        ignore_label, cls_weights)
