"""Time / profile the x4 up-sampling kernels on the config-2 logits shape."""
import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
ge.load_package()
from importlib import import_module
ops = import_module("robseg_b200.ops")
dev = torch.device("cuda:0")
B, C, S = 16, 150, 512
g = torch.Generator(device=dev).manual_seed(0)
low = torch.randn(B, C, S // 4, S // 4, device=dev, generator=g)
gup = torch.randn(B, C, S, S, device=dev, generator=g)
for name, fn in (("fwd", lambda: ops._upsample_fwd(low, S, S)), ("bwd", lambda: ops._upsample_bwd(gup, S // 4, S // 4))):
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[3]
    print(name, f"{ms:.3f} ms", f"{4 * (low.numel() + gup.numel()) / ms / 1e6:.0f} GB/s", flush=True)
