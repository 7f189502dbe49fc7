mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 -k "upsample or interpolate or segmenter or fast_upsample" 2>&1 | tail -12 > gpurun_out/pytest_walk.log); tail -6 gpurun_out/pytest_walk.log
python - <<'PY' 2>&1 | grep -v Warn | tee gpurun_out/upsample_probe_voc.log
import importlib, statistics, sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
ops = importlib.import_module("robust-segmentation_b200.ops")
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, inner):
    for _ in range(2): fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        if inner:
            ops.profile_start(); fn(); torch.cuda.synchronize()
            ts.append(sum(ms for n, _, ms in ops.profile_stop() if n == inner))
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
for B, C, s, S in [(24, 21, 119, 473), (24, 512, 59, 119), (24, 512, 29, 119), (24, 512, 14, 119), (24, 512, 29, 59), (24, 512, 14, 29)]:
    low = torch.randn(B, C, s, s, device=dev, generator=g); gup = torch.randn(B, C, S, S, device=dev, generator=g)
    nb = 4 * (low.numel() + gup.numel())
    lr = low.clone().requires_grad_(); up = F.interpolate(lr, size=(S, S), mode="bilinear", align_corners=False)
    r = [t(lambda: ops._upsample_fwd(low, S, S), "upsample_fwd"), t(lambda: ops._upsample_bwd(gup, s, s), "upsample_bwd"),
         t(lambda: F.interpolate(low, size=(S, S), mode="bilinear", align_corners=False), None),
         t(lambda: torch.autograd.grad(up, [lr], grad_outputs=gup, retain_graph=True), None)]
    print(f"[{B},{C},{s},{s}]->{S}: " + "  ".join(f"{n} {ms*1e3:7.1f} us {nb/ms/1e6:6.0f} GB/s" for n, ms in zip(("fwd", "bwd", "aten_fwd", "aten_bwd"), r)), flush=True)
    del up, lr, low, gup
PY
