"""Device time of the fused up-sampling loss kernel vs the three-kernel path it replaces.
   python scripts/loss_up_probe.py B C h R [kind]"""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.load_package()
from importlib import import_module
ops = import_module("robseg_b200.ops")
B, C, h, R = (int(v) for v in sys.argv[1:5])
kind = sys.argv[5] if len(sys.argv) > 5 else "mask-ce-avg"
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
low = 3 * torch.randn(B, C, h, h, device=dev, generator=g)
H = h * R
up = ops.upsample_bilinear(low, (H, H))
y = torch.randint(0, C, (B, H, H), device=dev, generator=g)
y = torch.where(torch.rand(y.shape, device=dev, generator=g) < 0.5, up.argmax(1), y)
dbuf = torch.empty_like(up)
dlow = torch.empty_like(low)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=7):
    """Median DEVICE time of the robseg launches inside fn (the wrappers' own CUDA events around each C call:
    host-side gaps between launches are not counted)."""
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        ops.profile_start(); fn(); torch.cuda.synchronize()
        ts.append(sum(ms for _, _, ms in ops.profile_stop()))
    return statistics.median(ts)
def three():
    u = ops._upsample_fwd(low, H, H)
    o = ops.loss_fwd_bwd(u, y, kind, None, dlogits_out=dbuf)
    ops._upsample_bwd(o.dlogits, h, h)
t3 = timeit(three)
tf = timeit(lambda: ops.loss_upsampled_fwd_bwd(low, y, kind, None, dlow_out=dlow))
tl = timeit(lambda: ops.loss_upsampled_fwd_bwd(low, y, kind, None, want_grad=False))
ta = timeit(lambda: ops.loss_upsampled_fwd_bwd(low, y, "argmax", want_grad=False, want_pred=True))
print(f"B={B} C={C} {h}^2 x{R} -> {H}^2 {kind}: three-kernel path {t3:.3f} ms | fused loss+grad {tf:.3f} ms | fused loss only {tl:.3f} ms | fused argmax {ta:.3f} ms "
      f"| {B*C*H*H/tf/1e6:.1f} G up-sampled logits/s")
