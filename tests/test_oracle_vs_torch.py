"""The oracle's closed forms against torch autograd in float64 on random problems (CPU).
Complements test_oracle_golden.py (which pins the oracle to the reference's own fp32 outputs)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import robseg_oracle as O


def torch_loss(z, y, kind, w):
    """Per-pixel losses written with stock torch ops in float64 (attack semantics, ignore = -1)."""
    valid = y != -1
    ce = F.cross_entropy(z, y, reduction="none", ignore_index=-1)
    hit = (z.argmax(1) == y) & valid
    if kind == "ce-avg":
        return ce
    if kind == "mask-ce-avg":
        return hit.double() * ce
    if kind == "mask-ce-bal":
        return hit.double() * F.cross_entropy(z, y, reduction="none", ignore_index=-1, weight=w)
    p = F.softmax(z, 1)
    q = F.one_hot(y.clamp(min=0), z.shape[1]).movedim(-1, 1).double()
    m = (p + q) / 2
    kl_qm = -m.gather(1, y.clamp(min=0).unsqueeze(1)).squeeze(1).log()  # KL(onehot || m) = -log m_y
    js = 0.5 * ((p * (p / m).log()).sum(1) + kl_qm)
    return valid.double() * js


@pytest.mark.parametrize("kind", ["ce-avg", "mask-ce-avg", "mask-ce-bal", "js-avg"])
@pytest.mark.parametrize("shape", [(2, 3, 11), (1, 21, 40), (3, 150, 17)])
def test_closed_forms_match_autograd_float64(kind, shape):
    B, C, P = shape
    g = torch.Generator().manual_seed(B * 1000 + C)
    z = (2.5 * torch.randn(B, C, P, generator=g, dtype=torch.float64)).requires_grad_()
    y = torch.randint(0, C, (B, P), generator=g)
    y = torch.where(torch.rand(B, P, generator=g) < 0.5, z.detach().argmax(1), y)
    y = torch.where(torch.rand(B, P, generator=g) < 0.15, torch.full_like(y, -1), y)
    w = 0.5 + torch.rand(C, generator=g, dtype=torch.float64)
    gscale = 0.1 + torch.rand(B, generator=g, dtype=torch.float64)
    lp = torch_loss(z, y, kind, w)
    (gz,) = torch.autograd.grad((lp.sum(1) * gscale).sum(), [z])
    o = O.loss_fwd_bwd(z.detach().numpy(), y.numpy(), kind, w.numpy(), grad_scale=gscale.numpy())
    np.testing.assert_allclose(o["loss_pix"], lp.detach().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(o["dlogits"], gz.numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(o["loss_img"], (lp.sum(1) * gscale).detach().numpy(), rtol=1e-10)
    up = torch.rand(B, P, generator=g, dtype=torch.float64)
    (gu,) = torch.autograd.grad((torch_loss(z, y, kind, w) * up).sum(), [z])
    np.testing.assert_allclose(O.loss_pixel_bwd(z.detach().numpy(), y.numpy(), kind, up.numpy(), w.numpy()),
                               gu.numpy(), rtol=1e-9, atol=1e-13)


def test_update_kernels_match_torch_fp32_chain():
    g = torch.Generator().manual_seed(0)
    shape = (4, 3, 9, 7)
    x = torch.rand(shape, generator=g)
    eps = 8 / 255
    xa = (x + eps * (2 * torch.rand(shape, generator=g) - 1)).clamp(0, 1)
    xo = (x + eps * (2 * torch.rand(shape, generator=g) - 1)).clamp(0, 1)
    gr = torch.randn(shape, generator=g)
    st = torch.tensor([2 * eps, eps, eps / 2, eps / 4]).view(-1, 1, 1, 1)
    for a in (1.0, 0.75):  # semseg/attacker.py:395-410 with torch CPU fp32 ops
        z = torch.clamp(torch.min(torch.max(xa + st * torch.sign(gr), x - eps), x + eps), 0.0, 1.0)
        ref = torch.clamp(torch.min(torch.max(xa + (z - xa) * a + (xa - xo) * (1 - a), x - eps), x + eps), 0.0, 1.0)
        assert np.array_equal(O.apgd_step(x.numpy(), xa.numpy(), xo.numpy(), gr.numpy(), st.flatten().numpy(), eps, a),
                              ref.numpy())
    d = eps * (2 * torch.rand(shape, generator=g) - 1)
    ref = ((x + (d + 1e-2 * torch.sign(gr))).clamp(0.0, 1.0) - x).clamp(-eps, eps)  # semseg/val.py:210-213
    assert np.array_equal(O.pgd_step(x.numpy(), d.numpy(), gr.numpy(), 1e-2, eps), ref.numpy())
    zz = x + 0.1 * torch.randn(shape, generator=g)
    ref = (x + (zz - x).clamp(-eps, eps)).clamp(0.0, 1.0)  # semseg/attacker.py:683-690
    assert np.array_equal(O.project_linf(zz.numpy(), x.numpy(), eps), ref.numpy())


@pytest.mark.parametrize("shape", [(2, 3, 8, 8, 32, 32), (1, 2, 5, 6, 13, 17), (1, 2, 30, 30, 119, 119), (1, 2, 14, 14, 119, 119),
                                   (1, 1, 6, 6, 6, 6), (1, 2, 1, 1, 16, 16), (1, 2, 9, 7, 5, 4), (1, 2, 32, 32, 512, 512)])
def test_bilinear_upsampling_restatement_matches_torch(shape):
    """The oracle's restatement of ATen's bilinear interpolation (align_corners=False), the op the
    reference's models call for their logits and pyramid maps, against torch itself: forward and
    the adjoint (autograd) on the CPU."""
    B, C, h, w, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(B, C, h, w, generator=g).requires_grad_()
    go = torch.randn(B, C, H, W, generator=g)
    ref = torch.nn.functional.interpolate(x, size=(H, W), mode="bilinear", align_corners=False)
    (gref,) = torch.autograd.grad(ref, [x], go)
    out = O.upsample_bilinear(x.detach().numpy(), H, W)
    gin = O.upsample_bilinear_bwd(go.numpy(), h, w)
    assert np.abs(out - ref.detach().numpy()).max() <= 1e-5 * np.abs(ref.detach().numpy()).max()
    assert np.abs(gin - gref.numpy()).max() <= 1e-5 * np.abs(gref.numpy()).max()
    # adjointness of the restatement itself, in float64
    lhs = float((out.astype(np.float64) * go.numpy()).sum())
    rhs = float((x.detach().numpy().astype(np.float64) * gin).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
