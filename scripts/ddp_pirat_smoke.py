"""PIR-AT training step under DDP (BASELINE config 4, scaled): torchrun --nproc-per-node 2 scripts/ddp_pirat_smoke.py"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.load_package()
from importlib import import_module
val = import_module("robseg_b200.semseg.val")
consumers = import_module("robseg_b200.consumers")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
C, B, S = 150, 4, 256
model = consumers.upernet_convnext("S", C, fast_upsample=True).to(dev)
ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4)
g = torch.Generator().manual_seed(100 + rank)
res = {}
for mode in ("reference-semantics", "input-grad-only"):
    attack = val.Pgd_Attack_1(epsilon=4 / 255, alpha=1e-2, num_iter=2, los="pgd", input_grad_only=mode != "reference-semantics")
    ts = []
    for it in range(4):
        img = torch.rand(B, 3, S, S, generator=g).to(dev)
        lbl = torch.randint(0, C, (B, S, S), generator=g).to(dev)
        torch.cuda.synchronize(); t0 = time.time()
        opt.zero_grad(set_to_none=True)
        ddp.eval()
        adv = attack.adv_attack(ddp, img, lbl)[0]      # tools/train_rob_seg.py:333-336
        ddp.train()
        loss, _ = ddp(adv, lbl)
        loss.backward()
        opt.step()
        torch.cuda.synchronize(); ts.append(time.time() - t0)
        assert torch.isfinite(loss), loss
        assert float((adv - img).abs().max()) <= 4 / 255 + 1e-6
    res[mode] = min(ts[1:])
# replicas stay in sync
chk = torch.stack([p.detach().float().sum() for p in model.parameters()]).sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
assert all(torch.equal(allc[0], c) for c in allc), allc
if rank == 0:
    print("DDP PIR-AT ok:", {k: f"{v * 1e3:.1f} ms/step" for k, v in res.items()}, f"{world} ranks, B={B}/rank, {S}x{S}, C={C}, ConvNeXt-S", flush=True)
dist.destroy_process_group()
