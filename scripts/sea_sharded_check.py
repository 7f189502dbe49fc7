"""Image-sharded SEA == single-rank SEA for every integer counter and aggregate.
torchrun --nproc-per-node 2 scripts/sea_sharded_check.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.load_package()
from importlib import import_module
sea = import_module("robseg_b200.tools.sea"); cons = import_module("robseg_b200.consumers")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.deterministic = True
C = 21
torch.manual_seed(0)
# conv-only consumer: deterministic under cudnn.deterministic (UperNet's internal ATen bilinear
# backward scatters with atomics, so its attack trajectories differ from run to run)
model = cons.TinySegNet(C, hidden=16, seed=5).to(dev).eval()
g = torch.Generator().manual_seed(3)
loader = []
for _ in range(4):
    x = torch.rand(2, 3, 64, 64, generator=g)
    with torch.no_grad():
        y = model(x.to(dev)).argmax(1).cpu()
    loader.append((x, y))
sharded = sea.run_sea(model, loader, C, eps=8 / 255, n_iter=6, device=dev, seed=11)
single = sea.run_sea(model, loader, C, eps=8 / 255, n_iter=6, device=dev, seed=11, shard=False)
keys = ["clean", "mask-ce-bal", "mask-ce-avg", "js-avg", "worst_Acc", "final_miou", "n_images"]
for k in keys:
    assert sharded[k] == single[k], (rank, k, sharded[k], single[k])
assert torch.equal(sharded["worst_Acc_indiv"], single["worst_Acc_indiv"])
if rank == 0:
    print("sharded SEA == single-rank SEA:", {k: sharded[k] for k in ("worst_Acc", "final_miou", "n_images")}, flush=True)
dist.destroy_process_group()
