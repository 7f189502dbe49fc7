"""B200 drop-in for the reference's ``semseg/losses.py`` (semseg/losses.py:6-109).

``CrossEntropy`` (mean reduction, ignore label, optional class weights, aux-head weighting)
runs on the fused loss kernel: forward = one pass producing the per-pixel weighted CE, backward
= one pass producing d/dlogits (``torch.ops.robseg.pixel_loss`` carries the autograd formula).
``OhemCrossEntropy`` and ``Dice`` are not on the SEA / PIR-AT path;
they are kept as plain-PyTorch bodies for name compatibility only.
"""
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
        if isinstance(preds, tuple):
            return sum([w * self._forward(pred, labels) for (pred, w) in zip(preds, self.aux_weights)])
        return self._forward(preds, labels)


class OhemCrossEntropy(nn.Module):
    """Unaccelerated (not on the hot path): same arithmetic as semseg/losses.py:30-63."""

    def __init__(self, ignore_label: int = 255, weight: Tensor = None, thresh: float = 0.7,
                 aux_weights: list = [1, 1]) -> None:
        super().__init__()
        self.ignore_label = ignore_label
        self.aux_weights = aux_weights
        self.thresh = -torch.log(torch.tensor(thresh, dtype=torch.float))
        self.criterion = nn.CrossEntropyLoss(weight=weight, ignore_index=ignore_label, reduction="none")

    def _forward(self, preds: Tensor, labels: Tensor) -> Tensor:
        n_min = labels[labels != self.ignore_label].numel() // 16
        loss = self.criterion(preds, labels).view(-1)
        hard = loss[loss > self.thresh]
        if hard.numel() < n_min:
            hard, _ = loss.topk(n_min)
        return torch.mean(hard)

    def forward(self, preds, labels: Tensor) -> Tensor:
        if isinstance(preds, tuple):
            return sum([w * self._forward(pred, labels) for (pred, w) in zip(preds, self.aux_weights)])
        return self._forward(preds, labels)


class Dice(nn.Module):
    """Unaccelerated (not on the hot path): same arithmetic as semseg/losses.py:66-93."""

    def __init__(self, delta: float = 0.5, aux_weights: list = [1, 0.4, 0.4]):
        super().__init__()
        self.delta = delta
        self.aux_weights = aux_weights

    def _forward(self, preds: Tensor, labels: Tensor) -> Tensor:
        n = preds.shape[1]
        onehot = F.one_hot(labels, n).permute(0, 3, 1, 2)
        tp = torch.sum(onehot * preds, dim=(2, 3))
        fn = torch.sum(onehot * (1 - preds), dim=(2, 3))
        fp = torch.sum((1 - onehot) * preds, dim=(2, 3))
        score = (tp + 1e-6) / (tp + self.delta * fn + (1 - self.delta) * fp + 1e-6)
        return (torch.sum(1 - score, dim=-1) / n).mean()

    def forward(self, preds, targets: Tensor) -> Tensor:
        if isinstance(preds, tuple):
            return sum([w * self._forward(pred, targets) for (pred, w) in zip(preds, self.aux_weights)])
        return self._forward(preds, targets)


__all__ = ["CrossEntropy", "OhemCrossEntropy", "Dice"]


def get_loss(loss_fn_name: str = "CrossEntropy", ignore_label: int = 255, cls_weights: Tensor = None):
    assert loss_fn_name in __all__, (
        f"Unavailable loss function name >> {loss_fn_name}.\nAvailable loss functions: {__all__}")
    if loss_fn_name == "Dice":
        return Dice()
    return {"CrossEntropy": CrossEntropy, "OhemCrossEntropy": OhemCrossEntropy}[loss_fn_name](
        ignore_label, cls_weights)
