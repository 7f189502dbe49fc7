"""B200 drop-in for the hot-path part of the reference's ``semseg/val.py``:
``Pgd_Attack`` (:130-178), ``Pgd_Attack_1`` (:181-218), ``losses`` (:121-127) and
``evaluate`` (:14-32).

Per PGD step the loss, its logit gradient and the per-image loss come out of ONE
``robseg_loss_fwd_bwd`` pass, and the delta update (sign step, [0,1] clamp, eps-ball clamp,
next model input) out of ONE ``robseg_pgd_step`` launch.

Reference quirks handled:
  Q6  the trainer passes ``epsilon=`` although ``Pgd_Attack.__init__`` names it ``eps``: both
      spellings are accepted; the scalar loss ``"pgd"`` cannot run through the reference's
      per-image best tracking (IndexError), so for scalar losses ``Pgd_Attack`` follows
      ``Pgd_Attack_1``'s "last delta" rule while keeping its own zero start and clamped input.
  Q7  the reference back-propagates with ``loss.backward()``, which also accumulates the
      attack-time parameter gradients into ``param.grad`` (and fires DDP's all-reduce).
      ``input_grad_only=False`` (default) reproduces that; ``True`` differentiates wrt the
      input only -- about a third less backward work and no attack-time all-reduce.
"""
import torch

from .. import ops
from .metrics import Metrics

_PER_IMAGE = ("mask-ce-avg", "js-avg")
_KINDS = ("pgd",) + _PER_IMAGE


def _kind_check(los):
    if los not in _KINDS:
        raise KeyError(f"loss {los!r} is not on the accelerated path; choose one of {_KINDS}")
    return los


class _PgdBase:
    clamp_input = True
    random_start = False

    def _setup(self, epsilon, alpha, num_iter, los, input_grad_only):
        self.epsilon = epsilon
        self.num_iter = num_iter
        self.los_name = _kind_check(los)
        self.alpha = alpha
        self.input_grad_only = input_grad_only

    def _attack(self, model, X, y):
        model.eval()
        if not X.is_cuda:
            X = X.cuda()
        X = X.detach().float().contiguous()
        y = y.to(X.device).long()
        B = X.shape[0]
        delta = torch.zeros_like(X)
        if self.random_start:
            delta.uniform_(-self.epsilon, self.epsilon)
        per_image = self.los_name in _PER_IMAGE
        track_best = per_image and not self.random_start
        if track_best:
            best_loss = torch.zeros(B, device=X.device)
            best_delta = torch.zeros_like(X)
            flags = torch.zeros(B, dtype=torch.int32, device=X.device)
        # "pgd" = F.cross_entropy(x, y): mean over pixels whose label is not -100 (val.py:122)
        ignore = -100 if self.los_name == "pgd" else -1
        gscale = None
        if self.los_name == "pgd":
            gscale = (1.0 / ((y != ignore) & (y >= 0)).sum().clamp(min=1).float()).reshape(1)
        x_in = torch.empty_like(X)
        s = X + delta
        x_in.copy_(s.clamp(0.0, 1.0) if self.clamp_input else s)
        logits = None
        dbuf = None
        for _ in range(self.num_iter):
            xin = x_in.detach().requires_grad_(True)
            logits = model(xin)
            if logits.dtype not in (torch.float32, torch.bfloat16):
                logits = logits.float()  # fp16 under autocast(AMP): F.cross_entropy up-casts too
            out = ops.loss_fwd_bwd(logits, y, self.los_name, None, grad_scale=gscale,
                                   want_grad=True, ignore_index=ignore, dlogits_out=dbuf)
            dbuf = out.dlogits
            if self.input_grad_only:
                (g,) = torch.autograd.grad(logits, [xin], grad_outputs=out.dlogits)
            else:  # loss.backward() semantics: parameter grads accumulate too (val.py:167,208)
                logits.backward(out.dlogits)
                g = xin.grad
            if track_best:  # val.py:158-163: rows whose loss is >= the running best
                ind = out.loss_img >= best_loss
                best_loss = torch.where(ind, out.loss_img, best_loss)
                flags.copy_(ind)
            ops.pgd_step(X, delta, g.contiguous(), self.alpha, self.epsilon,
                         mask_outside=self.clamp_input, x_next=x_in, clamp_next=self.clamp_input)
            if track_best:  # val.py:175: best_delta[ind] = delta[ind] (after the update)
                ops.row_select([(best_delta, delta, flags, None)], B, X.device)
        final = best_delta if track_best else delta
        x_adv = (X + final).clamp(0.0, 1.0)
        return x_adv.detach(), logits


class Pgd_Attack(_PgdBase):
    """semseg/val.py:130-178: zero start, clamped model input, per-image best-loss delta."""

    def __init__(self, eps=4.0 / 255.0, alpha=1e-2, num_iter=2, los="pgd", epsilon=None,
                 input_grad_only=False):
        self._setup(eps if epsilon is None else epsilon, alpha, num_iter, los, input_grad_only)

    def adv_attack(self, model, X, y, wt=None):
        x_adv, _ = self._attack(model, X, y)
        return x_adv, None, None


class Pgd_Attack_1(_PgdBase):
    """semseg/val.py:181-218: uniform random start, unclamped model input, last delta."""

    clamp_input = False
    random_start = True

    def __init__(self, epsilon=4.0 / 255.0, alpha=1e-2, num_iter=2, los="pgd", eps=None,
                 input_grad_only=False):
        self._setup(epsilon if eps is None else eps, alpha, num_iter, los, input_grad_only)

    def adv_attack(self, model, X, y):
        x_adv, logits = self._attack(model, X, y)
        return x_adv, logits, None


@torch.no_grad()
def evaluate(model, dataloader, device, cls, n_batches=-1):
    """Clean / adversarial validation (semseg/val.py:14-32).  The reference's redundant
    ``softmax`` before ``Metrics.update`` (:25) is dropped: update only takes the argmax."""
    print("Evaluating...")
    model.eval()
    metrics = Metrics(cls, -1, device)
    for i, (images, labels) in enumerate(dataloader):
        images = images.to(device)
        labels = labels.to(device)
        preds = model(images)
        metrics.update(preds, labels)
        if i + 1 == n_batches:
            break
    ious, miou = metrics.compute_iou()
    cla_acc, macc, aacc = metrics.compute_pixel_acc()
    f1, mf1 = metrics.compute_f1()
    return cla_acc, macc, aacc, f1, mf1, ious, miou
