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
    pred = z.argmax(1)
    ops.pixel_hist(pred, y, C, want_hist=True)
    ops.pixel_hist(pred, y, C)
x = torch.rand(3, 3, 17, 19, generator=g).to(dev)
ops.apgd_step(x, x.clone(), x.clone(), torch.randn_like(x), torch.full((3,), 0.05, device=dev), 0.03, 0.75, torch.empty_like(x))
ops.project_linf(x + 0.1, x, 0.03); ops.project_linf(None, x, 0.03, noise=torch.rand_like(x))
d = torch.zeros_like(x); ops.pgd_step(x, d, torch.randn_like(x), 0.01, 0.03, mask_outside=True, x_next=torch.empty_like(x))
for shp in [(1, 3, 8, 8, 32, 32), (1, 2, 7, 9, 28, 36), (1, 2, 5, 6, 13, 17), (1, 2, 2, 2, 32, 32), (1, 35, 33, 31, 132, 124)]:
    a = torch.randn(*shp[:4], generator=g).to(dev).requires_grad_()
    o = ops.upsample_bilinear(a, shp[4:]); o.sum().backward()
model = cons.TinySegNet(7, seed=1).to(dev).eval()
xx = torch.rand(2, 3, 16, 16, generator=g).to(dev)
yy = model(xx).argmax(1)
att.apgd_largereps(model, xx, yy, None, eps=8 / 255, n_iter=10, loss="mask-ce-avg", track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=7, return_pred=True)
inter = torch.randint(0, 9, (3, 4, 7)).to(dev); ops.sea_worst_acc(inter, inter + 1)
torch.cuda.synchronize()
print("sanitize pass done")
