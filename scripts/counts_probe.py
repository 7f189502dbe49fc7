"""loss kernel with / without the fused class counters on one shape: python scripts/counts_probe.py B C S [coherent]"""
import sys, statistics, torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
ge.load_package()
from importlib import import_module
ops = import_module("robseg_b200.ops")
B, C, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
coherent = len(sys.argv) > 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
z = 3 * torch.randn(B, C, S, S, device=dev, generator=g)
if coherent:
    nb = (S + 63) // 64
    blk = torch.randint(0, C, (B, nb, nb), device=dev, generator=g)
    y = blk.repeat_interleave(64, 1).repeat_interleave(64, 2)[:, :S, :S].contiguous()
    z.scatter_add_(1, y.unsqueeze(1), torch.full((B, 1, S, S), 30.0, device=dev))
else:
    y = torch.randint(0, C, (B, S, S), device=dev, generator=g)
d = torch.empty_like(z)
for wc in (False, True, False, True):
    ts = []
    for i in range(6):
        torch.cuda._sleep(1_000_000)
        ops.profile_start()
        ops.loss_fwd_bwd(z, y, "mask-ce-avg", None, dlogits_out=d, want_counts=wc)
        torch.cuda.synchronize()
        ts.append(sum(ms for _, _, ms in ops.profile_stop()))
    print(f"counts={wc}: {statistics.median(ts[1:]) * 1e3:.1f} us", flush=True)
