"""Not part of the bench contract: the UNMODIFIED reference attacker (copy under baseline/_ref) driving
its own UperNet on the SAME B200, same workload as bench.py's step.  Prints image-iterations/s."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import ref_shims
ref_shims.install()
ref_dir = os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, ref_dir)
os.chdir(ref_dir)
import torch
import semseg.attacker as RA
from semseg.models import UperNetForSemanticSegmentation
import bench
B, C, S, n_iter = 16, 150, 512, 10
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = UperNetForSemanticSegmentation("ConvNeXt-T_CVST", C, None).to(dev).eval()
x, y = bench.make_batch(B, C, S, 100, dev)
w = (0.5 + torch.rand(C, generator=torch.Generator().manual_seed(1)))  # CPU tensor, as tools/infer.py:297-301
def step():
    for loss in bench.LOSSES:
        x_adv, _, acc = RA.apgd_largereps(model, x.clone(), y, w, norm="Linf", eps=8 / 255, n_iter=n_iter, loss=loss,
                                          track_loss="ce-avg", use_rs=True, early_stop=True, num_classes=C)
        with torch.no_grad():
            model(x_adv).max(1)[1]
    torch.cuda.synchronize()
step()
ts = []
for _ in range(2):
    t0 = time.time(); step(); ts.append(time.time() - t0)
t = min(ts)
print(json.dumps({"impl": "reference-on-gpu", "image_iterations_per_s": round(B * n_iter * 3 / t, 2), "s_per_step": round(t, 3),
                  "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2**30, 1)}))
