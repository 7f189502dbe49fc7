"""B200 drop-in for the reference's ``semseg/metrics.py`` (semseg/metrics.py:21-60).

``Metrics.update`` is one fused argmax launch (``robseg_loss_fwd_bwd`` in ARGMAX mode) plus one
``robseg_pixel_hist`` launch that accumulates the exact int64 confusion matrix
``hist[target, pred]``; the reference's boolean gathers + ``bincount`` + float32 accumulation
(:27-33) are gone.  ``hist`` is exposed as the float32 view the reference keeps (exact while
every cell < 2**24); the finalisers replay the reference's float32
arithmetic on it so the printed numbers match.
"""
import torch
from torch import Tensor

from .. import ops


class Metrics:
    def __init__(self, num_classes: int, ignore_label: int, device) -> None:
        self.ignore_label = ignore_label
        self.num_classes = num_classes
        self.hist_int = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=device)
        self._hist_override = None

    @property
    def hist(self) -> Tensor:
        if self._hist_override is not None:
            return self._hist_override
        return self.hist_int.to(torch.float32)

    @hist.setter
    def hist(self, value: Tensor) -> None:  # the reference lets callers assign / reset hist
        self._hist_override = value

    def update(self, pred: Tensor, target: Tensor) -> None:
        """pred: [B,C,H,W] scores (any monotone transform of the logits), target: [B,H,W]."""
        if self._hist_override is not None:
            self.hist_int = self._hist_override.round().to(torch.int64).to(self.hist_int.device)
            self._hist_override = None
        if pred.dtype not in (torch.float32, torch.bfloat16):
            pred = pred.float()
        am = ops.loss_fwd_bwd(pred, target, "argmax", want_grad=False, want_pred=True,
                              ignore_index=self.ignore_label, want_stats=False).pred
        ops.pixel_hist(am, target, self.num_classes, self.ignore_label, hist_total=self.hist_int,
                       want_counts=False)

    # ---- finalisers: float32 arithmetic in the reference's order (semseg/metrics.py:35-60) ----
    @staticmethod
    def _percent(per_class: Tensor):
        """(per-class list, NaN-skipping mean), both x100 and rounded to 2 decimals."""
        mean = per_class[~per_class.isnan()].mean().item() * 100
        return (per_class * 100).cpu().numpy().round(2).tolist(), round(mean, 2)

    def _marginals(self):
        h = self.hist
        return h.diag(), h.sum(0), h.sum(1)

    def compute_iou(self):
        tp, pred_total, target_total = self._marginals()
        return self._percent(tp / (pred_total + target_total - tp))

    def compute_f1(self):
        tp, pred_total, target_total = self._marginals()
        return self._percent(2 * tp / (pred_total + target_total))

    def compute_pixel_acc(self):
        tp, _, target_total = self._marginals()
        per_class, macc = self._percent(tp / target_total)
        overall = (tp.sum() / self.hist.sum()) * 100
        return per_class, macc, overall.cpu().numpy().round(2)
