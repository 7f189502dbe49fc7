"""Histogram-kernel probe: device time of robseg_pixel_hist (events around the C call) for the
counters-only and the full-confusion kernels on random / coherent label maps.
  python scripts/hist_probe.py [B] [C]"""
import importlib
import statistics
import sys

import torch

sys.path.insert(0, ".")
ops = importlib.import_module("robust-segmentation_b200.ops")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C = int(sys.argv[2]) if len(sys.argv) > 2 else 150
S = 512
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
y = torch.randint(0, C, (B, S, S), device=dev, generator=g)
p = torch.where(torch.rand(B, S, S, device=dev, generator=g) < 0.5, y, torch.randint(0, C, (B, S, S), device=dev, generator=g))
blk = torch.randint(0, C, (B, S // 64, S // 64), device=dev, generator=g)
yc = blk.repeat_interleave(64, 1).repeat_interleave(64, 2).contiguous()
pblk = torch.where(torch.rand(blk.shape, device=dev, generator=g) < 0.7, blk, torch.randint_like(blk, C))
pc = pblk.repeat_interleave(64, 1).repeat_interleave(64, 2).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(name, fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(7):
        flush.zero_()
        ops.profile_start()
        fn()
        torch.cuda.synchronize()
        ts.append(sum(ms for n, _, ms in ops.profile_stop() if n == "pixel_hist"))
    ms = statistics.median(ts)
    print(f"B={B} C={C} {name:28s} {ms*1e3:8.1f} us  {16*y.numel()/ms/1e6:8.1f} GB/s", flush=True)


t("counts random", lambda: ops.pixel_hist(p, y, C))
t("counts coherent", lambda: ops.pixel_hist(pc, yc, C))
t("full random", lambda: ops.pixel_hist(p, y, C, want_hist=True, want_counts=False))
t("full coherent", lambda: ops.pixel_hist(pc, yc, C, want_hist=True, want_counts=False))
t("full+counts coherent", lambda: ops.pixel_hist(pc, yc, C, want_hist=True))
