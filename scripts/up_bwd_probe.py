"""Backward pow2 up-sampling kernels: strip / blocks-per-SM sweep (ROBSEG_UP_BWD_STRIP, ROBSEG_UP_BWD_BPS).
  python scripts/up_bwd_probe.py"""
import importlib
import os
import statistics
import sys

import torch

sys.path.insert(0, ".")
ops = importlib.import_module("robust-segmentation_b200.ops")
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, inner):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(7):
        flush.zero_()
        torch.cuda._sleep(400_000)
        ops.profile_start()
        fn()
        torch.cuda.synchronize()
        ts.append(sum(ms for n, _, ms in ops.profile_stop() if n == inner))
    return statistics.median(ts)


shapes = [(16, 150, 128, 512), (16, 512, 32, 128), (16, 512, 16, 128), (24, 21, 118, 472), (2, 21, 128, 512),
          (2, 512, 32, 128), (2, 512, 16, 128), (4, 150, 128, 512), (16, 150, 32, 512), (16, 150, 64, 512)]
for B, C, s, S in shapes:
    gup = torch.randn(B, C, S, S, device=dev, generator=g)
    nb = 4 * (gup.numel() + B * C * s * s)
    base = t(lambda: ops._upsample_bwd(gup, s, s), "upsample_bwd")
    print(f"[{B},{C},{S},{S}]->{s} ({nb/1e6:.0f} MB): default {base*1e3:6.1f} us {nb/base/1e6:5.0f} GB/s", flush=True)
    for strip in (8, 16, 32, 64):
        if strip > s:
            continue
        row = []
        for bps in (32, 48, 64, 96):
            os.environ["ROBSEG_UP_BWD_STRIP"], os.environ["ROBSEG_UP_BWD_BPS"] = str(strip), str(bps)
            ms = t(lambda: ops._upsample_bwd(gup, s, s), "upsample_bwd")
            row.append(f"bps{bps:2d} {ms*1e3:6.1f} us {nb/ms/1e6:5.0f}")
        print(f"    strip {strip:3d}: " + " | ".join(row), flush=True)
    os.environ.pop("ROBSEG_UP_BWD_STRIP"), os.environ.pop("ROBSEG_UP_BWD_BPS")
    del gup
