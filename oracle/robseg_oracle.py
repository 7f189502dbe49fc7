"""CPU oracle for the SEA / PIR-AT attack-side hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of what the reference (nmndeep/Robust-Segmentation)
computes on the path named in BASELINE.json; every function cites the reference
file:line it follows.  It is *not* product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the CPU baseline.  The product path
(``robust-segmentation_b200/``) never imports this module and has no CPU fallback.

Parity status: PINNED.  The reference ships no tests or golden vectors of its own
(SURVEY.md §4), so the pin is against outputs of the reference itself, generated in the
build container by ``tests/golden/make_golden.py`` (which imports the unmodified
reference from /root/reference through two import shims) and committed under
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every function here
against those fixtures.

Conventions: logits are ``[B, C, P]`` (P = H*W flattened, NCHW order), labels ``[B, P]``
int64 with -1 = ignore.  Float work is done in float64 from the closed forms of
SURVEY.md §10 unless a function says it replays float32 rounding step by step (the
APGD/PGD update, whose result must be bit-exact).
"""
from __future__ import annotations

import math
import random
import statistics

import numpy as np

LOSS_KINDS = ("ce", "mask-ce-avg", "mask-ce-bal", "js-avg")
_ALIAS = {"ce-avg": "ce", "pgd": "ce"}


def canonical_kind(kind: str) -> str:
    kind = _ALIAS.get(kind, kind)
    if kind not in LOSS_KINDS:
        raise ValueError(f"unknown loss kind {kind!r}")
    return kind


# --------------------------------------------------------------------------------------
# (1) fused per-pixel softmax + loss + d(loss)/d(logits)
# --------------------------------------------------------------------------------------
def pixel_terms(logits, labels, kind, weights=None, ignore_index=-1, dtype=np.float64):
    """Per-pixel quantities of one loss.

    Follows semseg/attacker.py:143-152 (masked_cross_entropy), :155-173
    (masked_cross_entropy_balanced), :187-234 (js_div_fn/js_loss), :251-257
    (criterion_dict) in the closed forms of SURVEY.md §10:

      ce          l = v (lse - z_y)                         coef = v
      mask-ce-avg l = m (lse - z_y)                         coef = m
      mask-ce-bal l = m w_y (lse - z_y)                     coef = m w_y
      js-avg      l = v/2 [2 ln2 + p_y ln p_y - (1+p_y) ln(1+p_y)]
                                                            coef = -v/2 p_y ln(p_y/(1+p_y))
      d l / d z_k = coef (p_k - 1[k = y])

    with v = [y != ignore], yhat = argmax (lowest index on ties, Tensor.max(1)[1]),
    m = v [yhat = y].
    """
    kind = canonical_kind(kind)
    z = np.asarray(logits).astype(dtype)
    y = np.asarray(labels).astype(np.int64)
    B, C, P = z.shape
    zmax = z.max(axis=1)
    pred = z.argmax(axis=1).astype(np.int64)  # first (lowest) index on ties
    e = np.exp(z - zmax[:, None, :])
    s = e.sum(axis=1)
    lse = zmax + np.log(s)
    v = y != ignore_index
    ys = np.where(v, y, 0)
    z_y = np.take_along_axis(z, ys[:, None, :], axis=1)[:, 0, :]
    logp_y = z_y - lse
    p_y = np.exp(logp_y)
    ce = np.where(v, -logp_y, 0.0)
    hit = v & (pred == y)
    if kind == "ce":
        loss, coef = ce, v.astype(dtype)
    elif kind == "mask-ce-avg":
        loss, coef = np.where(hit, ce, 0.0), hit.astype(dtype)
    elif kind == "mask-ce-bal":
        if weights is None:  # weights=None means unweighted
            w_y = np.ones_like(ce)
        else:
            w_y = np.asarray(weights).astype(dtype)[ys]
        loss, coef = np.where(hit, w_y * ce, 0.0), np.where(hit, w_y, 0.0)
    else:  # js-avg
        log1p = np.log1p(p_y)
        js = 0.5 * (2.0 * math.log(2.0) + p_y * logp_y - (1.0 + p_y) * log1p)
        loss = np.where(v, js, 0.0)
        coef = np.where(v, -0.5 * p_y * (logp_y - log1p), 0.0)
    p = e / s[:, None, :]
    return dict(loss=loss, coef=coef, p=p, ce=ce, pred=pred, valid=v, hit=hit, ys=ys)


def loss_fwd_bwd(logits, labels, kind, weights=None, grad_scale=None, ignore_index=-1,
                 dtype=np.float64, want_grad=True):
    """Fused pass: what one launch of ``robseg_loss_fwd_bwd`` must produce.

    loss_img[b]   = g_b * sum_p l          (pixel_to_img_loss, semseg/attacker.py:237-240,
                                             with g_b = 1/P; `.sum()` over b at :349)
    track_img[b]  = 1/P * sum_p v (lse - z_y)   (track loss "ce-avg", :353-361,473-475)
    dlogits       = g_b * coef * (p - onehot(y))  (autograd of the above wrt logits)
    pred          = argmax_c z              (:370-373, :485)
    correct[b]    = #[pred == y]            (:370-371; y = -1 never equals a class)
    valid[b]      = #[y != -1]
    """
    t = pixel_terms(logits, labels, kind, weights, ignore_index, dtype)
    B, C, P = np.asarray(logits).shape
    if grad_scale is None:
        g = np.full((B,), 1.0 / P, dtype=dtype)
    else:
        g = np.broadcast_to(np.asarray(grad_scale, dtype=dtype), (B,)).copy()
    out = dict(
        loss_pix=t["loss"],
        loss_img=g * t["loss"].sum(axis=1),
        track_img=t["ce"].sum(axis=1) / P,
        pred=t["pred"],
        correct=(t["pred"] == np.asarray(labels)).sum(axis=1).astype(np.int64),
        valid=t["valid"].sum(axis=1).astype(np.int64),
    )
    if want_grad:
        d = t["p"].copy()
        np.put_along_axis(d, t["ys"][:, None, :],
                          np.take_along_axis(d, t["ys"][:, None, :], axis=1) - 1.0, axis=1)
        out["dlogits"] = d * (g[:, None] * t["coef"])[:, None, :]
    return out


def loss_pixel_bwd(logits, labels, kind, gout_pix, weights=None, ignore_index=-1,
                   dtype=np.float64):
    """Backward of the per-pixel criterion (criterion_dict[...] -> [B,H,W]) for an
    arbitrary upstream gradient ``gout_pix[B,P]`` (semseg/attacker.py:251-257 + autograd)."""
    t = pixel_terms(logits, labels, kind, weights, ignore_index, dtype)
    d = t["p"].copy()
    np.put_along_axis(d, t["ys"][:, None, :],
                      np.take_along_axis(d, t["ys"][:, None, :], axis=1) - 1.0, axis=1)
    return d * (np.asarray(gout_pix, dtype=dtype) * t["coef"])[:, None, :]


# --------------------------------------------------------------------------------------
# (2) APGD / PGD step -- float32 replay, bit-exact
# --------------------------------------------------------------------------------------
_f32 = np.float32


def _sign(g):
    # torch.sign: (0 < x) - (x < 0); NaN -> 0
    return (g > 0).astype(_f32) - (g < 0).astype(_f32)


def apgd_step(x, x_adv, x_old, grad, step, eps, a):
    """One Linf APGD update, semseg/attacker.py:388-410, each torch op rounded to fp32:

      g2 = x_adv - x_old
      z  = clip01(min(max(x_adv + step_b * sign(grad), x - eps), x + eps))
      x' = clip01(min(max(x_adv + (z - x_adv) * a + g2 * (1 - a), x - eps), x + eps))

    x, x_adv, x_old, grad: float32 [B, ...]; step: float32 [B]; eps, a python floats.
    Returns x' (the new x_adv); the caller rotates x_old <- x_adv (:391).
    """
    x, x_adv, x_old, grad = (np.asarray(t, dtype=_f32) for t in (x, x_adv, x_old, grad))
    st = np.asarray(step, dtype=_f32).reshape((-1,) + (1,) * (x.ndim - 1))
    eps32, a32, b32 = _f32(eps), _f32(a), _f32(1.0 - a)
    lo, hi = x - eps32, x + eps32
    g2 = x_adv - x_old
    z = x_adv + st * _sign(grad)
    z = np.clip(np.minimum(np.maximum(z, lo), hi), _f32(0), _f32(1))
    t = x_adv + (z - x_adv) * a32
    t = t + g2 * b32
    return np.clip(np.minimum(np.maximum(t, lo), hi), _f32(0), _f32(1))


def project_linf(z, x, eps):
    """Stage hand-off projection, semseg/attacker.py:683-690: clip01(x + clip(z-x, +-eps))."""
    z, x = np.asarray(z, dtype=_f32), np.asarray(x, dtype=_f32)
    e = _f32(eps)
    return np.clip(x + np.clip(z - x, -e, e), _f32(0), _f32(1))


def random_start(x, eps, t):
    """semseg/attacker.py:292-294 given t = 2*rand_like(x)-1 (fp32): clip01(x + eps*t)."""
    x, t = np.asarray(x, dtype=_f32), np.asarray(t, dtype=_f32)
    return np.clip(x + _f32(eps) * t, _f32(0), _f32(1))


def pgd_step(X, delta, grad, alpha, eps):
    """PIR-AT inner step, semseg/val.py:169-172 / :210-213:
    delta <- clip(clip01(X + (delta + alpha*sign(grad))) - X, +-eps)."""
    X, delta, grad = (np.asarray(t, dtype=_f32) for t in (X, delta, grad))
    d = delta + _f32(alpha) * _sign(grad)
    d = np.clip(X + d, _f32(0), _f32(1)) - X
    return np.clip(d, -_f32(eps), _f32(eps))


# --------------------------------------------------------------------------------------
# (2b) step-size / best-point bookkeeping
# --------------------------------------------------------------------------------------
def apgd_schedule(n_iter):
    """Data-independent check cadence of semseg/attacker.py:322-329,528-551: returns the
    list of (iteration index i, window k) at which the oscillation check fires."""
    n_iter_2 = max(int(0.22 * n_iter), 1)
    n_iter_min = max(int(0.06 * n_iter), 1)
    size_decr = max(int(0.03 * n_iter), 1)
    k, counter3, checks = n_iter_2, 0, []
    for i in range(n_iter):
        counter3 += 1
        if counter3 == k:
            checks.append((i, k))
            counter3 = 0
            k = max(k - size_decr, n_iter_min)
    return checks


def check_oscillation(loss_steps, j, k, k3=0.75):
    """semseg/attacker.py:243-248; index j-k == -1 wraps to the last row."""
    t = np.zeros(loss_steps.shape[1], dtype=_f32)
    for c in range(k):
        t += (loss_steps[j - c] > loss_steps[j - c - 1]).astype(_f32)
    return (t <= _f32(k * k3)).astype(_f32)


class ApgdState:
    """Per-call state of apgd_train (semseg/attacker.py:308-383)."""

    def __init__(self, x_adv, grad, loss_indiv, acc0, pred0, n_iter, eps):
        B = x_adv.shape[0]
        self.x_adv = x_adv.copy()
        self.x_old = x_adv.copy()
        self.x_best = x_adv.copy()
        self.x_best_adv = x_adv.copy()
        self.grad = grad.copy()
        self.grad_best = grad.copy()
        self.pred_best = pred0.copy()
        self.loss_steps = np.zeros((n_iter, B), dtype=_f32)
        self.loss_best = loss_indiv.astype(_f32).copy()
        self.loss_best_last = self.loss_best.copy()
        self.reduced_last = np.ones(B, dtype=_f32)
        self.acc = acc0.astype(_f32).copy()
        self.step = (_f32(2.0 * eps) * np.ones(B, dtype=_f32))
        self.done = False


def apgd_bookkeep(st: ApgdState, i, x_adv, grad, loss_indiv, avg_acc, pred, check_k):
    """Everything after the forward/backward of iteration i, semseg/attacker.py:485-551.

    x_adv/grad are this iteration's point and gradient (grad is the stale one on the last
    iteration, :467-469); loss_indiv the per-image track loss; avg_acc the per-image accuracy
    with ignored pixels counted correct (:489); check_k = window k if the oscillation check
    fires at i else 0.  Mutates st (including the restart of rows of st.x_adv / st.grad)."""
    st.x_adv, st.grad = x_adv.copy(), grad.copy()
    ind_pred = avg_acc <= st.acc
    st.acc = np.minimum(st.acc, avg_acc)
    st.x_best_adv[ind_pred] = x_adv[ind_pred]
    st.pred_best[ind_pred] = pred[ind_pred]
    y1 = loss_indiv.astype(_f32)
    st.loss_steps[i] = y1
    ind = y1 > st.loss_best
    st.x_best[ind] = x_adv[ind]
    st.grad_best[ind] = grad[ind]
    st.loss_best[ind] = y1[ind]
    if check_k:
        osc = check_oscillation(st.loss_steps, i, check_k)
        no_impr = (_f32(1.0) - st.reduced_last) * (st.loss_best_last >= st.loss_best).astype(_f32)
        osc = np.maximum(osc, no_impr)
        st.reduced_last = osc.copy()
        st.loss_best_last = st.loss_best.copy()
        red = osc > 0
        st.step[red] /= _f32(2.0)
        st.x_adv[red] = st.x_best[red]
        st.grad[red] = st.grad_best[red]
    return ind_pred, ind


# --------------------------------------------------------------------------------------
# (3) per-image accuracy + confusion histogram, and the metric finalisers
# --------------------------------------------------------------------------------------
def pixel_hist(pred, labels, n_cls, ignore_index=-1):
    """Integer counters behind compute_iou_acc (semseg/attacker.py:9-52), Metrics.update
    (semseg/metrics.py:27-33), eval_performance (tools/infer.py:86-116) and evalSEA
    (tools/worse_only.py:30-66,383-394), per image:

      hist[b,t,p]  confusion counts over pixels with t != ignore  (hist[target, pred])
      inter[b,c]   #[pred == c and target == c]
      tgt[b,c]     #[target == c]
      prd[b,c]     #[pred == c and target != ignore]   (pred is set to ignore there, :20)
    """
    pred = np.asarray(pred).reshape(pred.shape[0], -1).astype(np.int64)
    lab = np.asarray(labels).reshape(labels.shape[0], -1).astype(np.int64)
    B = pred.shape[0]
    hist = np.zeros((B, n_cls, n_cls), dtype=np.int64)
    for b in range(B):
        keep = lab[b] != ignore_index
        idx = lab[b][keep] * n_cls + pred[b][keep]
        hist[b] = np.bincount(idx, minlength=n_cls * n_cls).reshape(n_cls, n_cls)
    inter = np.einsum("bcc->bc", hist).copy()
    return dict(hist=hist, inter=inter, tgt=hist.sum(2), prd=hist.sum(1))


def iou_acc_from_counts(inter, tgt, prd):
    """Finaliser of compute_iou_acc (semseg/attacker.py:29-50) in its float32 arithmetic.
    inter/tgt/prd are [C] integer totals.  Returns (m_acc, a_acc, m_iou) as float32."""
    a = np.asarray(inter).astype(_f32)
    n = np.asarray(tgt).astype(_f32)
    u = n + np.asarray(prd).astype(_f32) - a
    ind = n > 0
    m_acc = (a[ind] / n[ind]).astype(_f32)
    m_acc = _f32(m_acc.astype(np.float64).mean()) if ind.any() else _f32(np.nan)
    a_acc = _f32(a.astype(np.float64).sum()) / _f32(n.astype(np.float64).sum())
    indu = u > 0
    iou = (a[indu] / u[indu]).astype(_f32)
    m_iou = _f32(iou.astype(np.float64).mean()) if indu.any() else _f32(np.nan)
    return m_acc, _f32(a_acc), m_iou


def metrics_from_hist(hist):
    """Finalisers of semseg/metrics.py:35-60 on a [C,C] histogram (float32 arithmetic, NaN
    for empty classes skipped by the means, x100, round 2).  Returns a dict of python
    lists / floats."""
    h = np.asarray(hist).astype(_f32)
    diag = np.diag(h)
    with np.errstate(divide="ignore", invalid="ignore"):
        ious = diag / (h.sum(0) + h.sum(1) - diag)
        f1 = _f32(2) * diag / (h.sum(0) + h.sum(1))
        acc = diag / h.sum(1)
        a_acc = _f32(diag.astype(np.float64).sum()) / _f32(h.astype(np.float64).sum())

    def fin(v):
        m = float(_f32(v[~np.isnan(v)].astype(np.float64).mean())) * 100
        return np.round(v * _f32(100), 2).tolist(), round(m, 2)

    ious_l, miou = fin(ious)
    f1_l, mf1 = fin(f1)
    acc_l, macc = fin(acc)
    return dict(ious=ious_l, miou=miou, f1=f1_l, mf1=mf1, acc=acc_l, macc=macc,
                aacc=float(np.round(a_acc * _f32(100), 2)))


# --------------------------------------------------------------------------------------
# (4) SEA worst-case aggregation
# --------------------------------------------------------------------------------------
def sea_image_acc(inter, tgt):
    """Per-attack per-image aACC of evalSEA.worse_case_eval (tools/worse_only.py:383-398):
    sum_c inter / sum_c n_tgt in float32.  inter/tgt: [A,N,C] integer counts."""
    a = np.asarray(inter).sum(-1).astype(_f32)
    n = np.asarray(tgt).sum(-1).astype(_f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return a / n


def sea_worst_acc(acc_an):
    """tools/worse_only.py:403-408: min over attacks, mean over images; per-attack means.
    Means are taken in float64 here and compared to the reference's fp32 `.mean()` with
    a 1-ulp(fp32) tolerance in the tests; the product uses torch's own `.mean()`."""
    acc_an = np.asarray(acc_an, dtype=_f32)
    worst = acc_an.min(axis=0)
    return float(worst.astype(np.float64).mean()), acc_an.astype(np.float64).mean(axis=1)


def _miou(ints, unions):
    # tools/worse_only.py:69-76 (_compute_miou): python-double ratios, statistics.mean
    iou = [float(a) / float(b) for a, b in zip(ints, unions) if b != 0]
    return statistics.mean(iou)


def sea_worst_miou(cons_ints, cons_unions, rng=None, n_rounds=1000):
    """Greedy randomised worst-case mIoU, tools/worse_only.py:267-334.

    cons_ints/cons_unions: [A,N,C] per-attack per-image intersections / unions (exact
    integer counts; the reference holds them in fp32).  Running sums start from attack 0
    (:236-246), are re-rounded to fp32 whenever they pass through torch.tensor(list)
    (:311-316,323-326) and every candidate is scored as
    mean_c((run_int+d_int)/(run_union+d_union+1e-8)) over classes whose *running* union is
    non-zero (:79-93).  Class alignment is kept for globally absent classes (the reference
    shortens its lists there; fixtures keep every class present).
    Returns (final_miou, selected_attack_per_image)."""
    rng = rng or random
    ci = np.asarray(cons_ints, dtype=np.float64)
    cu = np.asarray(cons_unions, dtype=np.float64)
    A, N, C = ci.shape
    run_i = [float(_f32(v)) for v in ci[0].sum(0)]
    run_u = [float(_f32(v)) for v in cu[0].sum(0)]
    final = _miou(run_i, run_u)
    sel = [0] * N
    prev_best = 10
    for _ in range(n_rounds):
        order = list(range(N))
        rng.shuffle(order)
        for idx in order:
            for a in range(A):
                ri = [float(_f32(v)) for v in run_i]
                ru = [float(_f32(v)) for v in run_u]
                di = ci[a, idx] - ci[sel[idx], idx]
                du = cu[a, idx] - cu[sel[idx], idx]
                new_i = [ri[c] + di[c] for c in range(C)]
                new_u = [ru[c] + du[c] for c in range(C)]
                est = statistics.mean(
                    [new_i[c] / (new_u[c] + 1e-8) for c in range(C) if ru[c] != 0])
                if est < final:
                    sel[idx] = a
                    run_i, run_u = new_i, new_u
            final = _miou([float(_f32(v)) for v in run_i], [float(_f32(v)) for v in run_u])
        if prev_best - final <= 1e-6:
            break
        prev_best = final
    return final, sel


# --------------------------------------------------------------------------------------
# (5) the attack loops, with the model as a black box
# --------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------
# Bilinear up-sampling of the logits / pyramid maps
# ------------------------------------------------------------------------------------------
def _bilinear_taps(out_size, in_size):
    """Source taps of ``nn.functional.interpolate(..., mode="bilinear", align_corners=False)``
    as the reference calls it (semseg/models/uperforseg.py:193-198,282-303,416-418;
    semseg/models/segmenter.py:228).  The arithmetic lives in PyTorch, a dependency of the
    reference (torch==2.2.0 pinned, requirements.txt:70; 2.11 here, same formula): ATen's
    ``area_pixel_compute_source_index`` in float32 -- scale = in/out, src = scale*(dst+0.5)-0.5
    clamped at 0, i0 = floor(src), i1 = i0 + (i0 < in-1), w1 = src - i0, w0 = 1 - w1."""
    dst = np.arange(out_size, dtype=_f32)
    scale = _f32(in_size) / _f32(out_size)
    src = (scale * (dst + _f32(0.5)) - _f32(0.5)).astype(_f32)
    src = np.maximum(src, _f32(0.0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    w1 = (src - i0.astype(_f32)).astype(_f32)
    w0 = (_f32(1.0) - w1).astype(_f32)
    return i0, i1, w0, w1


def upsample_bilinear(x, H, W):
    """[..., h, w] float32 -> [..., H, W]; float32 products in ATen's order
    (w0y*(w0x*a + w1x*b) + w1y*(w0x*c + w1x*d))."""
    x = np.asarray(x, dtype=_f32)
    y0, y1, wy0, wy1 = _bilinear_taps(H, x.shape[-2])
    x0, x1, wx0, wx1 = _bilinear_taps(W, x.shape[-1])
    top = (wx0 * x[..., y0, :][..., :, x0] + wx1 * x[..., y0, :][..., :, x1]).astype(_f32)
    bot = (wx0 * x[..., y1, :][..., :, x0] + wx1 * x[..., y1, :][..., :, x1]).astype(_f32)
    return (wy0[:, None] * top + wy1[:, None] * bot).astype(_f32)


def upsample_bilinear_bwd(g, h, w):
    """Adjoint of :func:`upsample_bilinear`: [..., H, W] -> [..., h, w], accumulated in float64
    (the product's gather and ATen's atomic scatter differ from it only by float32 rounding)."""
    g = np.asarray(g, dtype=np.float64)
    H, W = g.shape[-2:]
    y0, y1, wy0, wy1 = _bilinear_taps(H, h)
    x0, x1, wx0, wx1 = _bilinear_taps(W, w)
    My = np.zeros((H, h))
    np.add.at(My, (np.arange(H), y0), wy0.astype(np.float64))
    np.add.at(My, (np.arange(H), y1), wy1.astype(np.float64))
    Mx = np.zeros((W, w))
    np.add.at(Mx, (np.arange(W), x0), wx0.astype(np.float64))
    np.add.at(Mx, (np.arange(W), x1), wx1.astype(np.float64))
    # two matrix products (x first, then y) instead of one three-operand einsum, which numpy evaluates as a single
    # O(H W h w) loop nest per plane: 14 s for two 512 x 512 planes against 0.02 s
    return np.matmul(My.T, np.matmul(g, Mx))


class TorchModelAdapter:
    """Wraps a torch.nn.Module (CPU, eval) as the black-box consumer the oracle drives:
    ``forward(x) -> logits [B,C,P]`` and ``vjp(dlogits) -> d/dx`` (float32 numpy)."""

    def __init__(self, module):
        import torch

        self.torch, self.m = torch, module
        self._x = self._out = None

    def forward(self, x, need_grad=True):
        torch = self.torch
        xt = torch.from_numpy(np.ascontiguousarray(x, dtype=_f32)).requires_grad_(need_grad)
        with torch.set_grad_enabled(need_grad):
            out = self.m(xt)
        self._x, self._out = xt, out
        self.hw = tuple(out.shape[2:])
        return out.detach().numpy().reshape(out.shape[0], out.shape[1], -1)

    def vjp(self, dlogits):
        torch = self.torch
        g = torch.from_numpy(np.ascontiguousarray(dlogits, dtype=_f32)).reshape(self._out.shape)
        (gx,) = torch.autograd.grad(self._out, [self._x], grad_outputs=g)
        return gx.numpy()


def _eval_point(model, x_adv, y, kind, weights, want_grad):
    logits = model.forward(x_adv, need_grad=want_grad)
    o = loss_fwd_bwd(logits, y, kind, weights, want_grad=want_grad)
    grad = model.vjp(o["dlogits"]) if want_grad else None
    return o, grad


def apgd_train(model, x, y, eps, n_iter=10, use_rs=False, loss="ce", early_stop=False,
               x_init=None, weights=None, rand_t=None, trace=None):
    """Linf apgd_train, semseg/attacker.py:260-571, with track_loss="ce-avg".

    x: float32 [B,3,H,W]; y: int64 [B,H,W]; rand_t = 2*rand_like(x)-1 supplied by the
    caller when use_rs (the RNG draw itself stays with the caller, :292-297).
    Returns (x_best, acc, loss_best, x_best_adv) like the reference."""
    x = np.asarray(x, dtype=_f32)
    B = x.shape[0]
    yf = np.asarray(y).reshape(B, -1)
    P = yf.shape[1]
    x_adv = x.copy()
    if use_rs:
        x_adv = random_start(x, eps, rand_t)
    if x_init is not None:
        x_adv = np.asarray(x_init, dtype=_f32).copy()
    x_adv = np.clip(x_adv, _f32(0), _f32(1))
    o, grad = _eval_point(model, x_adv, yf, loss, weights, True)
    acc0 = (o["correct"].astype(_f32) / _f32(P)).astype(_f32)  # -1 pixels count wrong here (:370-371)
    st = ApgdState(x_adv, grad, o["track_img"].astype(_f32), acc0, o["pred"], n_iter, eps)
    checks = dict(apgd_schedule(n_iter))
    if trace is not None:
        trace.append(x_adv.copy())
    for i in range(n_iter):
        a = 0.75 if i > 0 else 1.0
        x_new = apgd_step(x, st.x_adv, st.x_old, st.grad, st.step, eps, a)
        st.x_old = st.x_adv.copy()
        if trace is not None:
            trace.append(x_new.copy())
        want_grad = i < n_iter - 1
        o, g = _eval_point(model, x_new, yf, loss, weights, want_grad)
        if g is None:
            g = st.grad
        # ignored pixels count as correct inside the loop (:489)
        avg_acc = ((o["correct"] + (P - o["valid"])).astype(_f32) / _f32(P)).astype(_f32)
        apgd_bookkeep(st, i, x_new, g, o["track_img"].astype(_f32), avg_acc, o["pred"],
                      checks.get(i, 0))
        if early_stop and st.acc.sum() == 0:
            break
    return st.x_best, st.acc, st.loss_best, st.x_best_adv


def apgd_largereps(model, x, y, weights, eps=8.0 / 255.0, n_iter=10, loss="ce",
                   early_stop=False, use_rs=False, rand_ts=None, trace=None):
    """3-stage large-eps schedule, semseg/attacker.py:662-728: iterations
    [.3n,.3n,rest] at [2eps,1.5eps,eps], each stage started from the projection of the
    previous stage's x_best_adv.  rand_ts: one noise tensor per stage (drawn even when
    x_init overrides it)."""
    n_iters = [int(c * n_iter) for c in (0.3, 0.3)]
    n_iters.append(n_iter - sum(n_iters))
    epss = [c * eps for c in (2, 1.5, 1)]
    x_init, acc, loss_best = None, None, None
    for s, (it, e) in enumerate(zip(n_iters, epss)):
        if x_init is not None:
            x_init = project_linf(x_init, x, e)
        _, acc, loss_best, x_init = apgd_train(
            model, x, y, e, n_iter=it, use_rs=use_rs, loss=loss, early_stop=early_stop,
            x_init=x_init, weights=weights,
            rand_t=None if rand_ts is None else rand_ts[s], trace=trace)
    return x_init, loss_best, acc


def pgd_attack(model, X, y, eps, alpha, num_iter, loss="pgd", random_start_delta=None,
               clamp_input=True, track_best=True, ignore_index=-100):
    """PIR-AT inner attack, semseg/val.py:130-178 (Pgd_Attack: delta0 = 0, input clamped,
    per-image best-loss tracking for per-image losses) and :181-218 (Pgd_Attack_1: random
    start, unclamped input, last delta).  loss "pgd" = mean CE over valid pixels (val.py:122)."""
    X = np.asarray(X, dtype=_f32)
    B = X.shape[0]
    yf = np.asarray(y).reshape(B, -1)
    P = yf.shape[1]
    delta = np.zeros_like(X) if random_start_delta is None else np.asarray(random_start_delta, _f32).copy()
    best_delta = np.zeros_like(X)
    best_loss = np.zeros(B, dtype=_f32)
    kind = canonical_kind(loss)
    scalar_loss = loss == "pgd"
    for _ in range(num_iter):
        xin = np.clip(X + delta, _f32(0), _f32(1)) if clamp_input else X + delta
        logits = model.forward(xin, need_grad=True)
        if scalar_loss:
            n_valid = max(int((yf != ignore_index).sum()), 1)
            o = loss_fwd_bwd(logits, yf, kind, grad_scale=1.0 / n_valid, ignore_index=ignore_index)
        else:
            o = loss_fwd_bwd(logits, yf, kind, ignore_index=ignore_index)
        g = model.vjp(o["dlogits"])
        if track_best and not scalar_loss:
            li = o["loss_img"].astype(_f32)
            ind = li >= best_loss
            best_loss[ind] = li[ind]
        delta = pgd_step(X, delta, g, alpha, eps)
        if track_best and not scalar_loss:
            best_delta[ind] = delta[ind]
    final = best_delta if (track_best and not scalar_loss) else delta
    return np.clip(X + final, _f32(0), _f32(1))
