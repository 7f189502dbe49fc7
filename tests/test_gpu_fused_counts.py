"""Per-image class counters taken inside the loss kernels' argmax pass (robseg_loss_fwd_bwd_counts /
robseg_loss_upsampled_fwd_bwd_counts; compute_iou_acc, semseg/attacker.py:9-52) against the oracle's
restatement of the reference's counting and against robseg_pixel_hist on the prediction map of the same
launch.  Integer outputs: bit-exact (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import robseg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(pkg):
    from importlib import import_module

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    names = dict(ops=".ops", attacker=".semseg.attacker", consumers=".consumers", graphs=".graphs", sea=".tools.sea",
                 worse=".tools.worse_only")
    return type("M", (), {k: import_module("robseg_b200" + v) for k, v in names.items()})


DEV = torch.device("cuda:0")


def _labels(z, C, g, coherent, frac_ignore=0.1):
    B, _, H, W = z.shape
    if coherent:  # 16x16 constant regions: whole warp rows share a class (one reduction per row)
        blk = torch.randint(0, C, (B, (H + 15) // 16, (W + 15) // 16), generator=g)
        y = blk.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :H, :W].contiguous()
    else:
        y = torch.randint(0, C, (B, H, W), generator=g)
    y = torch.where(torch.rand(B, H, W, generator=g) < 0.5, z.float().argmax(1), y)
    return torch.where(torch.rand(B, H, W, generator=g) < frac_ignore, torch.full_like(y, -1), y)


def _check(out, y, C):
    B = y.shape[0]
    pred = out.pred.cpu().numpy().reshape(B, -1)
    ref = O.pixel_hist(pred, y.numpy().reshape(B, -1), C)
    cnt = out.counts.cpu().numpy()
    assert cnt.shape == (B, 3, C) and cnt.dtype == np.int64
    for k, name in enumerate(("inter", "tgt", "prd")):
        assert np.array_equal(cnt[:, k], ref[name]), name
    return cnt


@pytest.mark.parametrize("shape", [(2, 21, 64, 64), (1, 150, 32, 96), (3, 7, 19, 23), (2, 21, 37, 41), (1, 151, 16, 24),
                                   (2, 300, 8, 8)])
@pytest.mark.parametrize("kind", ["argmax", "mask-ce-avg", "js-avg"])
@pytest.mark.parametrize("coherent", [False, True])
def test_loss_kernel_counts(mods, shape, kind, coherent):
    """TMA path (aligned rows, partial last tile), the generic strided / one-pixel paths (odd H*W, C > 256),
    ARGMAX-only and loss launches, random and spatially coherent label maps."""
    B, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape) + coherent)
    z = 3 * torch.randn(B, C, H, W, generator=g)
    y = _labels(z, C, g, coherent)
    out = mods.ops.loss_fwd_bwd(z.to(DEV), y.to(DEV), kind, None, want_pred=True, want_counts=True,
                                want_grad=kind != "argmax")
    cnt = _check(out, y, C)
    hist = mods.ops.pixel_hist(out.pred, y.to(DEV), C)
    for k, name in enumerate(("inter", "tgt", "prd")):
        assert np.array_equal(cnt[:, k], hist[name].cpu().numpy())
    if kind != "argmax":  # the counters ride along: nothing else changes
        plain = mods.ops.loss_fwd_bwd(z.to(DEV), y.to(DEV), kind, None, want_pred=True)
        assert torch.equal(plain.dlogits, out.dlogits) and torch.equal(plain.loss_img, out.loss_img)
        assert torch.equal(plain.pred, out.pred) and plain.counts is None
        assert np.array_equal(cnt[:, 0].sum(1), out.correct.cpu().numpy())
        assert np.array_equal(cnt[:, 1].sum(1), out.valid.cpu().numpy())


def test_loss_kernel_counts_bf16_and_alternative_schedule(mods, monkeypatch):
    B, C, H, W = 2, 150, 64, 64
    g = torch.Generator().manual_seed(9)
    z = (3 * torch.randn(B, C, H, W, generator=g)).bfloat16()
    y = _labels(z, C, g, True)
    out = mods.ops.loss_fwd_bwd(z.to(DEV), y.to(DEV), "mask-ce-bal", None, want_pred=True, want_counts=True)
    _check(out, y, C)
    monkeypatch.setenv("ROBSEG_LOSS_G", "2")  # two warps per stage: only one of them counts
    out2 = mods.ops.loss_fwd_bwd(z.float().to(DEV), y.to(DEV), "mask-ce-avg", None, want_pred=True, want_counts=True)
    _check(out2, y, C)


def test_loss_kernel_counts_full_size(mods):
    """BASELINE config-2 tile stream (C=150, 512x512), uniformly random labels: the worst case for the
    warp-aggregated reductions."""
    B, C, S = 2, 150, 512
    g = torch.Generator(device=DEV).manual_seed(0)
    z = 3 * torch.randn(B, C, S, S, device=DEV, generator=g)
    y = torch.randint(0, C, (B, S, S), device=DEV, generator=g)
    y = torch.where(torch.rand(B, S, S, device=DEV, generator=g) < 0.5, z.argmax(1), y)
    out = mods.ops.loss_fwd_bwd(z, y, "mask-ce-avg", None, want_pred=True, want_counts=True)
    hist = mods.ops.pixel_hist(out.pred, y, C)
    for k, name in enumerate(("inter", "tgt", "prd")):
        assert torch.equal(out.counts[:, k], hist[name]), name
    assert int(out.counts[:, 1].sum()) == B * S * S


@pytest.mark.parametrize("shape", [(2, 21, 8, 8, 4), (1, 150, 6, 5, 16), (2, 7, 5, 9, 8), (1, 33, 16, 16, 2)])
@pytest.mark.parametrize("kind", ["argmax", "mask-ce-bal"])
def test_fused_upsampling_loss_counts(mods, shape, kind):
    B, C, h, w, R = shape
    g = torch.Generator().manual_seed(sum(shape))
    low = 3 * torch.randn(B, C, h, w, generator=g)
    up = F.interpolate(low, size=(h * R, w * R), mode="bilinear", align_corners=False)
    y = _labels(up, C, g, False, frac_ignore=0.05)
    out = mods.ops.loss_upsampled_fwd_bwd(low.to(DEV), y.to(DEV), kind, None, want_pred=True, want_counts=True,
                                          want_grad=kind != "argmax")
    _check(out, y, C)


@pytest.mark.parametrize("graph", [False, True])
def test_attack_returns_the_counters_of_its_adversarial_point(mods, graph):
    """apgd_largereps(return_counts=True): the counters equal robseg_pixel_hist on the prediction map the same
    attack returns (return_pred), eagerly and with one CUDA graph per iteration; verbose=True reports
    compute_iou_acc from them."""
    C = 9
    model = mods.consumers.TinySegNet(C, seed=4).to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    x = torch.rand(3, 3, 32, 32, generator=g).to(DEV)
    with torch.no_grad():
        y = model(x).argmax(1)
    y[0, :2] = -1
    m = mods.graphs.GraphedModel(model, x) if graph else model
    det = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        torch.manual_seed(3)
        x_adv, lb, acc, pred, cnt = mods.attacker.apgd_largereps(
            m, x, y, None, eps=8 / 255, n_iter=12, loss="mask-ce-avg", track_loss="ce-avg", use_rs=True,
            early_stop=True, num_classes=C, return_pred=True, return_counts=True)
        hist = mods.ops.pixel_hist(pred, y, C)
        for k, name in enumerate(("inter", "tgt", "prd")):
            assert torch.equal(cnt[:, k], hist[name]), name
        torch.manual_seed(3)
        x2, lb2, acc2, cnt2 = mods.attacker.apgd_largereps(
            m, x, y, None, eps=8 / 255, n_iter=12, loss="mask-ce-avg", track_loss="ce-avg", use_rs=True,
            early_stop=True, num_classes=C, return_counts=True)
        assert torch.equal(x2, x_adv) and torch.equal(cnt2, cnt) and torch.equal(acc2, acc)
    finally:
        torch.backends.cudnn.deterministic = det
