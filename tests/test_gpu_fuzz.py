"""Seeded random sweep of the loss entry points over the launcher's dispatch boundaries, against the
CPU oracle (semseg/attacker.py:143-240,370-373,485-498 restated in oracle/robseg_oracle.py).

The fixed shape lists of test_gpu_parity.py pin the paths one by one; this sweep draws (B, C, H, W),
loss kind, dtype, requested outputs, per-image gradient scales and the byte alignment of the logits /
gradient / label pointers at random, so that every combination the launcher can pick -- TMA rows of
4 / 2 / 1 pixels per lane, logits in registers (C <= 24), 16-byte over-fetch, 4-byte-copy strided and
one-pixel kernels, counted and uncounted -- is hit with ragged last tiles, images smaller than a tile
and pointers that are only 4- or 8-byte aligned.  Integers bit-exact, fp32 <= 1e-5, bf16 <= 1e-2."""
import random

import numpy as np
import pytest
import torch

import robseg_oracle as O

pytestmark = pytest.mark.gpu
KINDS = ["mask-ce-avg", "mask-ce-bal", "js-avg", "ce-avg"]
# class counts either side of every launcher threshold: kRegChannels = 24, the 2-stage over-fetch limits
# (13 / 27), the TMA stage caps (24 KB / 40 KB at 4, 8, 16 bytes per lane), ADE20K (150 / 151), > 255
CLASSES = [1, 2, 3, 7, 12, 13, 14, 19, 21, 24, 25, 27, 28, 40, 48, 49, 64, 79, 80, 81, 96, 97, 150, 151, 171, 256, 300]


def _dev():
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _carve(t, off_elems):
    """A contiguous copy of ``t`` whose data pointer sits ``off_elems`` elements past a 256-byte boundary."""
    buf = torch.empty(t.numel() + 64, dtype=t.dtype, device=_dev())
    v = buf[off_elems:off_elems + t.numel()].view(t.shape)
    v.copy_(t)
    return v


def _draw(seed):
    r = random.Random(seed)
    C = r.choice(CLASSES)
    B = r.randint(1, 4)
    # a third of the cases: tiny images (smaller than one tile); else rows up to ~1.5k pixels, any parity
    if r.random() < 0.33:
        H, W = r.randint(1, 6), r.randint(1, 9)
    else:
        H, W = r.randint(3, 40), r.randint(3, 40)
    if C >= 150:
        H, W = min(H, 24), min(W, 24)
    return dict(B=B, C=C, H=H, W=W, kind=r.choice(KINDS), bf16=r.random() < 0.3,
                off_z=r.choice([0, 0, 1, 2, 3, 4]), off_d=r.choice([0, 0, 1, 2, 4]), off_y=r.choice([0, 0, 1]),
                want_pred=r.random() < 0.5, want_counts=r.random() < 0.5, want_loss_pix=r.random() < 0.3,
                scale=r.choice([None, "tensor", "float"]), frac_ignore=r.choice([0.0, 0.1, 0.6, 1.0]),
                sigma=r.choice([0.5, 3.0, 12.0]), weights=r.random() < 0.7)


@pytest.mark.parametrize("seed", range(96))
def test_loss_entry_points_random_dispatch_vs_oracle(pkg, seed):
    from importlib import import_module

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ops = import_module("robseg_b200.ops")
    c = _draw(seed)
    B, C, H, W, kind = c["B"], c["C"], c["H"], c["W"], c["kind"]
    g = torch.Generator().manual_seed(1000 + seed)
    z = c["sigma"] * torch.randn(B, C, H, W, generator=g)
    if c["bf16"]:
        z = z.bfloat16()
    zf = z.float()
    y = torch.randint(0, C, (B, H, W), generator=g)
    y = torch.where(torch.rand(B, H, W, generator=g) < 0.5, zf.argmax(1), y)
    y = torch.where(torch.rand(B, H, W, generator=g) < c["frac_ignore"], torch.full_like(y, -1), y)
    w = 0.5 + torch.rand(C, generator=g) if c["weights"] else None
    scale = {None: None, "float": 0.37, "tensor": 0.1 + torch.rand(B, generator=g)}[c["scale"]]
    zd, yd = _carve(z, c["off_z"]), _carve(y, c["off_y"])
    dl = _carve(torch.zeros_like(z), c["off_d"])
    out = ops.loss_fwd_bwd(zd, yd, kind, None if w is None else w.to(_dev()),
                           grad_scale=scale.to(_dev()) if torch.is_tensor(scale) else scale,
                           want_pred=c["want_pred"], want_loss_pix=c["want_loss_pix"], want_counts=c["want_counts"],
                           dlogits_out=dl)
    ref = O.loss_fwd_bwd(zf.numpy().reshape(B, C, -1), y.numpy().reshape(B, -1), kind,
                         None if w is None else w.numpy(),
                         grad_scale=scale.numpy() if torch.is_tensor(scale) else scale)
    tol = 1e-2 if c["bf16"] else 1e-5
    assert out.dlogits.data_ptr() == dl.data_ptr()
    assert np.array_equal(out.correct.cpu().numpy(), ref["correct"]), c
    assert np.array_equal(out.valid.cpu().numpy(), ref["valid"]), c
    if c["want_pred"]:
        assert np.array_equal(out.pred.cpu().numpy().reshape(B, -1), ref["pred"]), c
    if c["want_counts"]:
        h = O.pixel_hist(ref["pred"].reshape(B, H, W), y.numpy(), C)
        cnt = out.counts.cpu().numpy()
        for i, k in enumerate(("inter", "tgt", "prd")):
            assert np.array_equal(cnt[:, i], h[k]), (k, c)
    d = out.dlogits.float().cpu().numpy().reshape(B, C, -1)
    dmax = np.abs(ref["dlogits"]).max()
    if dmax > 0:
        assert np.abs(d - ref["dlogits"]).max() <= tol * dmax, c
    else:
        assert not d.any(), c
    if c["want_loss_pix"]:
        lp = out.loss_pix.cpu().numpy().reshape(B, -1)
        assert np.abs(lp - ref["loss_pix"]).max() <= 1e-5 * max(np.abs(ref["loss_pix"]).max(), 1e-30) + 1e-7, c
    np.testing.assert_allclose(out.loss_img.cpu().numpy(), ref["loss_img"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out.track_img.cpu().numpy(), ref["track_img"], rtol=1e-5, atol=1e-7)
    # the loss-only and argmax launches of the same inputs (other kernels / stage budgets): identical statistics
    lo = ops.loss_fwd_bwd(zd, yd, kind, None if w is None else w.to(_dev()),
                          grad_scale=scale.to(_dev()) if torch.is_tensor(scale) else scale, want_grad=False,
                          want_counts=c["want_counts"])
    assert torch.equal(lo.correct, out.correct) and torch.equal(lo.valid, out.valid), c
    np.testing.assert_allclose(lo.loss_img.cpu().numpy(), ref["loss_img"], rtol=1e-5, atol=1e-7)
    am = ops.loss_fwd_bwd(zd, yd, "argmax", want_grad=False, want_pred=True, want_counts=True)
    assert np.array_equal(am.pred.cpu().numpy().reshape(B, -1), ref["pred"]), c
    if c["want_counts"]:
        assert torch.equal(am.counts, out.counts) and torch.equal(lo.counts, out.counts), c


def _draw_up(seed):
    r = random.Random(10_000 + seed)
    mode = r.choice(["pow2", "pow2", "pow2", "any", "any", "same", "down"])
    h, w = r.choice([1, 2, 3, 5, 6, 8, 15, 16, 17, 30, 31, 32, 33, 40, 64]), r.choice([1, 2, 3, 4, 7, 14, 16, 17, 29, 30, 31, 32, 33, 45, 64, 70])
    if mode == "pow2":
        R = r.choice([2, 4, 8, 16])
        if R >= 8:
            h, w = min(h, 33), min(w, 33)
        H, W = R * h, R * w
    elif mode == "any":
        H, W = h + r.randint(0, 3 * h + 5), w + r.randint(0, 3 * w + 5)
    elif mode == "same":
        H, W = h, w
    else:  # down-sampling: the tile fallback
        H, W = max(1, h - r.randint(0, h // 2)), max(1, w - r.randint(0, w // 2))
    planes = r.choice([1, 2, 3, 7, 40, 300]) if h * w * 16 < 40_000 else r.choice([1, 2, 5])
    B = r.choice([1, 2]) if planes % 2 == 0 else 1
    return dict(B=B, C=planes // B, h=h, w=w, H=H, W=W, sliced=r.random() < 0.3)


@pytest.mark.parametrize("seed", range(72))
def test_upsample_random_shapes_vs_aten(pkg, seed):
    """Seeded random sweep of robseg_upsample_bilinear_fwd / _bwd / _bwd_strided over the launchers' dispatch space
    (x2 cell kernels, pow2 with 1 / 2 / 4 threads per cell, planes too small for them, walk-down kernels for unaligned
    rows and non-integer ratios, the tile fallback for down-sampling; backward with and without halo lanes, short and
    long strips, gradients read in place from a channel slice) against ATen on the same device
    (F.interpolate(..., 'bilinear', align_corners=False) and its autograd; semseg/models/uperforseg.py:193-198,282-303,
    416-418, segmenter.py:228): forward <= 2e-6, backward <= 1e-5 relative, repeated runs bit-identical."""
    from importlib import import_module

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ops = import_module("robseg_b200.ops")
    c = _draw_up(seed)
    B, C, h, w, H, W = c["B"], c["C"], c["h"], c["w"], c["H"], c["W"]
    g = torch.Generator().manual_seed(20_000 + seed)
    x = torch.randn(B, C, h, w, generator=g).to(_dev())
    if c["sliced"]:  # the decode head's case: the gradient is a channel slice of a concatenated gradient
        big = torch.randn(B, C + 3, H, W, generator=g).to(_dev())
        go = big[:, 2:2 + C]
    else:
        go = torch.randn(B, C, H, W, generator=g).to(_dev())
    xr = x.clone().requires_grad_()
    ref = torch.nn.functional.interpolate(xr, size=(H, W), mode="bilinear", align_corners=False)
    (gref,) = torch.autograd.grad(ref, [xr], grad_outputs=go)
    xo = x.clone().requires_grad_()
    out = ops.upsample_bilinear(xo, (H, W))
    (gours,) = torch.autograd.grad(out, [xo], grad_outputs=go)
    assert out.shape == ref.shape and gours.shape == gref.shape, c
    assert _rel(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= 2e-6, c
    assert _rel(gours.cpu().numpy(), gref.cpu().numpy()) <= 1e-5, c
    out2 = ops.upsample_bilinear(xo, (H, W))
    (g2,) = torch.autograd.grad(out2, [xo], grad_outputs=go)
    assert torch.equal(out2, out) and torch.equal(g2, gours), c
