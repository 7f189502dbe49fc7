"""Fused bilinear up-sampling + loss + gradient w.r.t. the low-resolution logits
(robseg_loss_upsampled_fwd_bwd) against
    F.interpolate(low, size, mode="bilinear", align_corners=False)  ->  the two-kernel path / the oracle.
Tolerances: the in-kernel interpolation is a lerp (3 FMAs), ATen's a 4-term weighted sum -> logits agree
to ~1 ulp, losses / gradients to <= 1e-5 relative (BASELINE.json's fp32 bound); counters exact unless the
1-ulp difference flips a near-tie of the argmax (none on these seeds: asserted equal)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import robseg_oracle as O

pytestmark = pytest.mark.gpu
KINDS = ["mask-ce-avg", "mask-ce-bal", "js-avg", "ce-avg"]


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as ge

    ge.load_package()
    from importlib import import_module

    import_module("robseg_b200._lib").load()
    return import_module("robseg_b200.ops")


def _problem(B, C, h, w, R, seed, frac_ignore=0.05):
    g = torch.Generator().manual_seed(seed)
    low = 3 * torch.randn(B, C, h, w, generator=g)
    up = F.interpolate(low, size=(h * R, w * R), mode="bilinear", align_corners=False)
    y = torch.randint(0, C, (B, h * R, w * R), generator=g)
    y = torch.where(torch.rand(y.shape, generator=g) < 0.5, up.argmax(1), y)
    y = torch.where(torch.rand(y.shape, generator=g) < frac_ignore, torch.full_like(y, -1), y)
    wts = 0.5 + torch.rand(C, generator=g)
    return low, up, y, wts


@pytest.mark.parametrize("shape", [(2, 21, 8, 8, 4), (1, 150, 6, 5, 16), (2, 7, 5, 9, 8), (1, 33, 16, 16, 2),
                                   (1, 151, 3, 3, 4), (2, 19, 1, 1, 16), (1, 5, 40, 24, 4)])
@pytest.mark.parametrize("kind", KINDS)
def test_fused_vs_oracle(ops, shape, kind):
    B, C, h, w, R = shape
    low, up, y, wts = _problem(B, C, h, w, R, seed=sum(shape))
    dev = torch.device("cuda")
    out = ops.loss_upsampled_fwd_bwd(low.to(dev), y.to(dev), kind, wts.to(dev), want_pred=True)
    # oracle: float64 loss / dlogits on the ATen-interpolated logits, pulled back through the interpolation
    ref = O.loss_fwd_bwd(up.numpy().reshape(B, C, -1), y.numpy().reshape(B, -1), kind, wts.numpy())
    dl = O.upsample_bilinear_bwd(ref["dlogits"].reshape(B, C, h * R, w * R), h, w)
    d = out.dlogits.cpu().numpy()
    assert np.abs(d - dl).max() <= 1e-5 * np.abs(dl).max() + 1e-12
    np.testing.assert_allclose(out.loss_img.cpu().numpy(), ref["loss_img"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out.track_img.cpu().numpy(), ref["track_img"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(out.pred.cpu().numpy().reshape(B, -1), ref["pred"])
    assert np.array_equal(out.correct.cpu().numpy(), ref["correct"])
    assert np.array_equal(out.valid.cpu().numpy(), ref["valid"])


@pytest.mark.parametrize("R,h", [(4, 128), (16, 32)])
def test_fused_full_size_vs_two_kernel_path(ops, R, h):
    """BASELINE config-2 / config-3 logit shapes: fused == up-sampling kernel + loss kernel + gather
    backward; argmax-only and loss-only launches agree with the fused one; runs are bit-reproducible."""
    B, C = 2, 150
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    low = 3 * torch.randn(B, C, h, h, device=dev, generator=g)
    up = ops.upsample_bilinear(low, (h * R, h * R))
    y = torch.randint(0, C, (B, h * R, h * R), device=dev, generator=g)
    y = torch.where(torch.rand(y.shape, device=dev, generator=g) < 0.5, up.argmax(1), y)
    wts = 0.5 + torch.rand(C, device=dev, generator=g)
    for kind in KINDS:
        a = ops.loss_upsampled_fwd_bwd(low, y, kind, wts, want_pred=True)
        b = ops.loss_fwd_bwd(up, y, kind, wts, want_pred=True)
        db = ops._upsample_bwd(b.dlogits, h, h)
        assert float((a.dlogits - db).abs().max()) <= 1e-5 * float(db.abs().max())
        torch.testing.assert_close(a.loss_img, b.loss_img, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(a.track_img, b.track_img, rtol=1e-5, atol=1e-7)
        mism = (a.pred != b.pred)
        assert int(mism.sum()) <= 2  # 1-ulp logit differences may flip an exact near-tie
        assert int((a.correct - b.correct).abs().max()) <= 2 and torch.equal(a.valid, b.valid)
        a2 = ops.loss_upsampled_fwd_bwd(low, y, kind, wts, want_pred=True)
        assert torch.equal(a.dlogits, a2.dlogits) and torch.equal(a.loss_img, a2.loss_img)
        lo = ops.loss_upsampled_fwd_bwd(low, y, kind, wts, want_grad=False)
        assert torch.equal(lo.loss_img, a.loss_img) and lo.dlogits is None
    am = ops.loss_upsampled_fwd_bwd(low, y, "argmax", want_grad=False, want_pred=True)
    assert torch.equal(am.pred, a.pred) and torch.equal(am.correct, a.correct)


def test_fused_autograd_through_interpolate(ops):
    """dlow equals torch.autograd's pull-back of the reference loss through F.interpolate."""
    B, C, h, R = 2, 21, 16, 4
    dev = torch.device("cuda")
    low, up, y, wts = _problem(B, C, h, h, R, seed=11)
    low_d = low.to(dev).requires_grad_()
    z = F.interpolate(low_d, size=(h * R, h * R), mode="bilinear", align_corners=False)
    yd = y.to(dev)
    mask = ((z.max(1)[1] == yd) & (yd != -1)).float()
    lp = mask * F.cross_entropy(z, yd, reduction="none", ignore_index=-1)
    (gt,) = torch.autograd.grad(lp.view(B, -1).mean(-1).sum(), [low_d])
    out = ops.loss_upsampled_fwd_bwd(low.to(dev), yd, "mask-ce-avg")
    assert float((out.dlogits - gt).abs().max()) <= 1e-5 * float(gt.abs().max())


def test_fused_rejects_unsupported(ops):
    dev = torch.device("cuda")
    low = torch.zeros(1, 3, 5, 5, device=dev)
    with pytest.raises(ValueError):
        ops.loss_upsampled_fwd_bwd(low, torch.zeros(1, 15, 15, dtype=torch.int64, device=dev), "ce-avg")  # x3
    with pytest.raises(RuntimeError):
        ops.loss_upsampled_fwd_bwd(low.cpu(), torch.zeros(1, 20, 20, dtype=torch.int64), "ce-avg")


def test_apgd_step_fused_equals_step_after_row_select(ops):
    """robseg_apgd_step_fused == the row copies of attacker.py:494-495,523-525,546-548 followed by
    robseg_apgd_step, bit for bit, for every flag combination (incl. restart rows)."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(5)
    B, n = 8, (3, 20, 24)
    mk = lambda: torch.rand(B, *n, device=dev, generator=g)  # noqa: E731
    x, xa, xo, gr, xba, xb, gb = mk(), mk(), mk(), mk() - 0.5, mk(), mk(), mk() - 0.5
    step = torch.rand(B, device=dev, generator=g) * 0.05
    flags = torch.tensor([[1, 0, 1, 0, 1, 0, 1, 0], [1, 1, 0, 0, 1, 1, 0, 0], [1, 1, 1, 1, 0, 0, 0, 0]],
                         dtype=torch.int32, device=dev)
    # reference sequence on copies
    r = [t.clone() for t in (xa, gr, xba, xb, gb)]
    ops.row_select([(r[2], r[0], flags[0], None), (r[3], r[0], flags[1], None), (r[4], r[1], flags[1], None)], B, dev)
    ops.row_select([(r[0], r[3], flags[2], flags[1]), (r[1], r[4], flags[2], flags[1])], B, dev)
    want = ops.apgd_step(x, r[0], xo, r[1], step, 8 / 255, 0.75, torch.empty_like(x))
    f = [t.clone() for t in (xa, gr, xba, xb, gb)]
    got = ops.apgd_step_fused(x, f[0], xo, f[1], step, 8 / 255, 0.75, torch.empty_like(x), flags, f[2], f[3], f[4])
    assert torch.equal(got, want)
    for a, b in zip(r, f):
        assert torch.equal(a, b)


class _UpNet(torch.nn.Module):
    """conv -> tanh -> conv at 1/4 resolution, bilinear x4 to the input size: the reference models'
    output structure (uperforseg.py:416-418) in miniature."""

    def __init__(self, C, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.c1 = torch.nn.Conv2d(3, 8, 4, stride=4)
        self.c2 = torch.nn.Conv2d(8, C, 3, padding=1)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.4 if p.dim() > 1 else 0.1))

    def low(self, x):
        return self.c2(torch.tanh(self.c1(x - 0.5)))

    def forward(self, x):
        return F.interpolate(self.low(x), size=x.shape[2:], mode="bilinear", align_corners=False)


@pytest.mark.parametrize("loss", ["mask-ce-avg", "js-avg"])
def test_attack_through_fused_logit_upsampling(ops, loss):
    """apgd_largereps on a model that offers forward_lowres (dropin.accelerate(..., fuse_loss=True)
    protocol) == the same attack on the full-resolution logits, up to sign flips at |grad| ~ 0."""
    from importlib import import_module

    att = import_module("robseg_b200.semseg.attacker")
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda")
    C = 9
    plain = _UpNet(C).to(dev).eval()
    fused = _UpNet(C).to(dev).eval()
    fused.forward_lowres = lambda x: fused.low(x)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(3, 3, 64, 64, generator=g).to(dev)
    with torch.no_grad():
        y = plain(x).argmax(1)
    kw = dict(norm="Linf", eps=8 / 255, n_iter=10, loss=loss, track_loss="ce-avg", use_rs=True, early_stop=True,
              num_classes=C, return_pred=True)
    torch.manual_seed(9)
    xa, la, aa, pa = att.apgd_largereps(plain, x, y, None, **kw)
    n0 = import_module("robseg_b200._lib").launches
    torch.manual_seed(9)
    xb, lb, ab, pb = att.apgd_largereps(fused, x, y, None, **kw)
    assert float((xa - xb).abs().gt(1e-6).float().mean()) <= 0.05
    assert float((aa - ab).abs().max()) <= 5.0 / (64 * 64)
    torch.testing.assert_close(la, lb, rtol=2e-2, atol=1e-4)
    with torch.no_grad():  # the returned argmax map is that of the returned point
        assert float((fused(xb).argmax(1) != pb).float().mean()) <= 1e-3
    assert float(ab.mean()) < 1.0 and import_module("robseg_b200._lib").launches > n0
