"""Forward pow2 up-sampling kernels: blocks-per-SM sweep (ROBSEG_UP_FWD_BPS) on the logit and decode-head shapes.
  python scripts/up_fwd_probe.py"""
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
        torch.cuda._sleep(400_000)  # keep the queue busy while the host enqueues the call
        ops.profile_start()
        fn()
        torch.cuda.synchronize()
        ts.append(sum(ms for n, _, ms in ops.profile_stop() if n == inner))
    return statistics.median(ts)


shapes = [(16, 150, 128, 512), (16, 512, 32, 128), (16, 512, 16, 128), (16, 512, 64, 256), (24, 21, 118, 472), (2, 21, 128, 512)]
for B, C, s, S in shapes:
    low = torch.randn(B, C, s, s, device=dev, generator=g)
    nb = 4 * (low.numel() + B * C * S * S)
    row = []
    for bps in (32, 24, 16, 12, 8, 6, 4):
        os.environ["ROBSEG_UP_FWD_BPS"] = str(bps)
        ms = t(lambda: ops._upsample_fwd(low, S, S), "upsample_fwd")
        row.append(f"bps{bps:2d} {ms*1e3:6.1f} us {nb/ms/1e6:5.0f}")
    os.environ.pop("ROBSEG_UP_FWD_BPS")
    print(f"[{B},{C},{s},{s}]->{S} ({nb/1e6:.0f} MB): " + " | ".join(row), flush=True)
    del low
