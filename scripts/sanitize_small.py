"""Small-shape pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
ge.load_package()
from importlib import import_module
ops = import_module("robseg_b200.ops"); att = import_module("robseg_b200.semseg.attacker")
cons = import_module("robseg_b200.consumers")
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
import os
for B, C, H, W in [(2, 21, 32, 32), (1, 150, 24, 40), (1, 151, 13, 11), (1, 64, 16, 24), (1, 21, 33, 37), (1, 300, 6, 6)]:
    z = (3 * torch.randn(B, C, H, W, generator=g)).to(dev)
    y = torch.randint(-1, C, (B, H, W), generator=g).to(dev)
    w = (0.5 + torch.rand(C, generator=g)).to(dev)
    for kind in ("mask-ce-bal", "js-avg", "argmax"):
        for env in ({}, {"ROBSEG_LOSS_G": "2"}, {"ROBSEG_LOSS_VEC": "1", "ROBSEG_LOSS_SLOTS": "2", "ROBSEG_LOSS_WARPS": "2"}):
            os.environ.update(env)
            ops.loss_fwd_bwd(z, y, kind, w, want_pred=True, want_loss_pix=kind != "argmax", want_grad=kind != "argmax")
            ops.loss_fwd_bwd(z.bfloat16(), y, kind, w, want_grad=kind != "argmax")
            for k in env: os.environ.pop(k)
    # round 2: class counters in the argmax pass, over-fetch generic path at every width, the 4-byte-copy kernels
    for env in ({}, {"ROBSEG_LOSS_GENERIC_OVF": "4"}, {"ROBSEG_LOSS_GENERIC_OVF": "1"}, {"ROBSEG_LOSS_GENERIC_OVF": "0"},
                {"ROBSEG_LOSS_G": "2"}):
        os.environ.update(env)
        ops.loss_fwd_bwd(z, y, "mask-ce-avg", w, want_pred=True, want_counts=True)
        ops.loss_fwd_bwd(z, y, "argmax", want_grad=False, want_stats=False, want_counts=True)
        ops.loss_fwd_bwd(z.bfloat16(), y, "js-avg", w, want_counts=True)
        for k in env: os.environ.pop(k)
    pred = z.argmax(1)
    ops.pixel_hist(pred, y, C, want_hist=True)
    ops.pixel_hist(pred, y, C)
x = torch.rand(3, 3, 17, 19, generator=g).to(dev)
ops.apgd_step(x, x.clone(), x.clone(), torch.randn_like(x), torch.full((3,), 0.05, device=dev), 0.03, 0.75, torch.empty_like(x))
ops.project_linf(x + 0.1, x, 0.03); ops.project_linf(None, x, 0.03, noise=torch.rand_like(x))
d = torch.zeros_like(x); ops.pgd_step(x, d, torch.randn_like(x), 0.01, 0.03, mask_outside=True, x_next=torch.empty_like(x))
for shp in [(1, 3, 8, 8, 32, 32), (1, 2, 7, 9, 28, 36), (1, 2, 5, 6, 13, 17), (1, 2, 2, 2, 32, 32), (1, 35, 33, 31, 132, 124),
            (1, 3, 7, 9, 14, 18), (1, 2, 5, 61, 10, 122), (2, 3, 16, 16, 128, 128), (1, 2, 3, 5, 48, 80), (1, 2, 1, 1, 16, 16),
            (1, 2, 6, 6, 16, 16), (1, 2, 33, 32, 66, 64), (1, 2, 30, 30, 119, 119), (1, 2, 14, 14, 119, 119),
            (1, 2, 59, 60, 119, 121), (1, 1, 70, 5, 100, 9),
            # exact x2 with even sides: the 2x2-cells-per-thread kernels (borders, one 2x2 plane, plane loop)
            (2, 3, 16, 16, 32, 32), (1, 2, 6, 10, 12, 20), (3, 4, 2, 2, 4, 4), (1, 700, 4, 4, 8, 8),
            # x16 / x8 forward with four / two threads per cell, backward without halo lanes (w <= 32) and with
            # halved strips (few planes of a tall image), w = 32 exactly (all 32 lanes own a column)
            (1, 2, 8, 8, 128, 128), (1, 3, 5, 7, 80, 112), (1, 2, 130, 6, 520, 24), (1, 2, 9, 32, 36, 128),
            (1, 2, 9, 33, 72, 264), (2, 2, 32, 32, 256, 256)]:
    a = torch.randn(*shp[:4], generator=g).to(dev).requires_grad_()
    o = ops.upsample_bilinear(a, shp[4:]); o.sum().backward()
# gradient read in place from a channel slice of a concatenated gradient (strided planes)
aa = [torch.randn(2, 3, 6, 10, generator=g).to(dev).requires_grad_() for _ in range(2)]
torch.cat([ops.upsample_bilinear(t, (12, 20)) for t in aa], 1).square().sum().backward()
# histogram kernels: blocks crossing image boundaries, odd plane sizes, several folds of the byte counters
for n, C, H, W in [(5, 150, 37, 41), (3, 21, 128, 130), (2, 513, 16, 16), (1, 150, 600, 512), (1, 700, 8, 8)]:
    tt = torch.randint(-1, C, (n, H, W), generator=g).to(dev)
    pp = torch.where(torch.rand(n, H, W, generator=g).to(dev) < 0.5, tt.clamp(min=0), torch.randint(0, C, (n, H, W), generator=g).to(dev))
    ops.pixel_hist(pp, tt, C)
    if C <= 226:
        ops.pixel_hist(pp, tt, C, want_hist=True)
os.environ["ROBSEG_CNT_PER_SM"] = "1"
tt = torch.randint(0, 150, (40, 512, 512), generator=g).to(dev)  # > 240 pixels per lane: folds inside the loop
ops.pixel_hist(tt.flip(0), tt, 150)
os.environ.pop("ROBSEG_CNT_PER_SM")
model = cons.TinySegNet(7, seed=1).to(dev).eval()
xx = torch.rand(2, 3, 16, 16, generator=g).to(dev)
yy = model(xx).argmax(1)
att.apgd_largereps(model, xx, yy, None, eps=8 / 255, n_iter=10, loss="mask-ce-avg", track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=7, return_pred=True)
att.apgd_largereps(model, xx, yy, None, eps=8 / 255, n_iter=10, loss="js-avg", track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=7, return_counts=True)
gm = import_module("robseg_b200.graphs").GraphedModel(model, xx)  # control-block step / bookkeeping kernels under graph replay
att.apgd_largereps(gm, xx, yy, None, eps=8 / 255, n_iter=10, loss="mask-ce-avg", track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=7, return_counts=True)
# fused up-sampling loss (x2 / x4 / x16, gradient + counters)
for B, C, h, w_, R in [(2, 21, 8, 8, 4), (1, 150, 3, 5, 16), (1, 33, 6, 6, 2), (1, 7, 5, 9, 8)]:
    low = (3 * torch.randn(B, C, h, w_, generator=g)).to(dev)
    yl = torch.randint(-1, C, (B, h * R, w_ * R), generator=g).to(dev)
    ops.loss_upsampled_fwd_bwd(low, yl, "mask-ce-bal", None, want_pred=True, want_counts=True)
    ops.loss_upsampled_fwd_bwd(low, yl, "argmax", want_grad=False, want_counts=True)
inter = torch.randint(0, 9, (3, 4, 7)).to(dev); ops.sea_worst_acc(inter, inter + 1)
torch.cuda.synchronize()
print("sanitize pass done")
