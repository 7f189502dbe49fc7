"""torch.profiler over ONE bench step (full SEA on 16 x 512^2): GPU busy time vs wall time, and where the
GPU time outside the consumer's convolutions / GEMMs goes.  python scripts/step_profile.py [--graph]"""
import importlib
import sys
import time
import types

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import bench

args = types.SimpleNamespace(batch=16, classes=150, size=512, n_iter=10, eps=8.0, reforward=False, variant="T")
import __graft_entry__ as ge

ge.load_package()
mods = {k: importlib.import_module("robseg_b200." + v) for k, v in dict(
    attacker="semseg.attacker", ops="ops", dist="dist", lib="_lib", consumers="consumers",
    worse="tools.worse_only", graphs="graphs").items()}
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = mods["consumers"].upernet_convnext("T", 150, fast_upsample="all").to(dev).eval()
if "--graph" in sys.argv:
    model = mods["graphs"].GraphedModel(model, torch.rand(16, 3, 512, 512, device=dev))
w = (0.5 + torch.rand(150, generator=torch.Generator().manual_seed(1))).to(dev)
x, y = bench.make_batch(16, 150, 512, 100, dev)
for _ in range(2):
    torch.manual_seed(1234)
    bench.sea_step(mods, model, x, y, w, args, 1)
torch.cuda.synchronize()
t0 = time.time()
torch.manual_seed(1234)
bench.sea_step(mods, model, x, y, w, args, 1)
torch.cuda.synchronize()
print(f"step wall {1e3 * (time.time() - t0):.1f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    torch.manual_seed(1234)
    bench.sea_step(mods, model, x, y, w, args, 1)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time for e in ev) / 1e3
print(f"GPU kernel+memcpy time {busy:.1f} ms over {len(ev)} device events")
# idle gaps on the device timeline
iv = sorted((e.time_range.start, e.time_range.end) for e in ev)
gap, end, big = 0.0, iv[0][1], []
for s, e in iv[1:]:
    if s > end:
        gap += s - end
        if s - end > 200:
            big.append((s - iv[0][0], s - end))
    end = max(end, e)
print(f"device timeline span {(end - iv[0][0]) / 1e3:.1f} ms, idle {gap / 1e3:.1f} ms; gaps > 0.2 ms: {len(big)}, total {sum(b for _, b in big) / 1e3:.1f} ms")
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=70))
