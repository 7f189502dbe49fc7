import sys, time, torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
ge.load_package()
from importlib import import_module
ops = import_module("robseg_b200.ops")
B, C, S, kind, grad = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], bool(int(sys.argv[5]))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
z = 3 * torch.randn(B, C, S, S, device=dev, generator=g)
if len(sys.argv) > 6 and sys.argv[6] == "bf16":
    z = z.bfloat16()
y = torch.randint(0, C, (B, S, S), device=dev, generator=g)
y = torch.where(torch.rand(B, S, S, device=dev, generator=g) < 0.5, z.argmax(1), y)
d = torch.empty_like(z) if grad else None
ts = []
for i in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    o = ops.loss_fwd_bwd(z, y, kind, None, want_grad=grad, dlogits_out=d)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
nbytes = (2 if grad else 1) * z.numel() * z.element_size() + 8 * y.numel()
print(f"{sys.argv[1:]} {ms:.3f} ms {nbytes / ms / 1e6:.0f} GB/s", flush=True)
