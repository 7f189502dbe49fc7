"""Where the consumer's forward/backward time goes (torch.profiler, grouped by op and input shape).
  python scripts/consumer_profile.py [batch] [classes] [all|logits|stock] [upernet|segmenter]"""
import importlib
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
consumers = importlib.import_module("robust-segmentation_b200.consumers")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C = int(sys.argv[2]) if len(sys.argv) > 2 else 150
dev = torch.device("cuda:0")
torch.manual_seed(0)
mode = {"all": "all", "logits": True, "stock": False}[sys.argv[3] if len(sys.argv) > 3 else "logits"]
if len(sys.argv) > 4 and sys.argv[4] == "segmenter":
    model = consumers.segmenter_vit("S", C, 512, fast_upsample=bool(mode)).to(dev).eval()
else:
    model = consumers.upernet_convnext("T", C, fast_upsample=mode).to(dev).eval()
x = torch.rand(B, 3, 512, 512, device=dev, requires_grad=True)
up = torch.randn(B, C, 512, 512, device=dev)


def it():
    logits = model(x)
    torch.autograd.grad(logits, [x], grad_outputs=up)


for _ in range(3):
    it()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    it()
b.record()
torch.cuda.synchronize()
print(f"fwd+bwd: {a.elapsed_time(b)/3:.2f} ms / iteration (B={B}, C={C})")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for _ in range(3):
        it()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=45,
                                                          max_name_column_width=60, max_shapes_column_width=90))
