"""Device-resident SEA driver: the ``__main__`` loop of the reference's ``tools/infer.py``
(:332-403) for the part that sits on the hot path, with the B200 bookkeeping.

    stats = run_sea(model, loader, n_cls, eps=8/255, n_iter=300, weights=bal_weights)

For every batch: the three SEA attacks (``mask-ce-bal``, ``mask-ce-avg``, ``js-avg``) through
``apgd_largereps``; the exact int64 per-image class counters of each adversarial point come back from
the attack itself (``return_counts``: taken in the loss kernel's argmax pass, robseg_loss_fwd_bwd_counts)
instead of a D2H copy of ``x_adv``, a later re-forward (tools/infer.py:151,82-90) and 2*C masked
reductions (:94-116) -- no prediction map and no histogram launch on this route.  With
``torch.distributed`` initialised every rank attacks its own contiguous shard of the loader's batches
and ONE int64 all-reduce merges the counters.  The aggregation is the same
code path as ``evalSEA``: per-attack mAcc / aAcc / mIoU (float32 finaliser of tools/infer.py:99-116),
image-wise worst aACC (tools/worse_only.py:396-408) and the greedy worst-case mIoU (:267-334).
"""
import random

import torch
import torch.distributed as dist

from .. import dist as rdist
from .. import ops
from ..semseg import attacker
from .infer import _finalize
from .worse_only import SEED, greedy_worst_miou

LOSSES = ("mask-ce-bal", "mask-ce-avg", "js-avg")


def _own_batches(loader, rank, world, n_batches):
    """(sizes of ALL batches, index of this rank's first batch, iterable over this rank's batches only).

    A rank must not decode or hold batches it does not attack (ADE20K val at 512^2 is ~6 GB of
    host memory per copy).  For a sequential ``DataLoader`` the batch sizes follow from
    ``len(dataset)`` and the rank iterates a ``Subset`` of its own images; for any other sized
    iterable the foreign batches are skipped while iterating and only their sizes are kept."""
    import torch.utils.data as tud

    def limit(nb):
        return nb if n_batches is None or n_batches < 0 else min(nb, n_batches)

    if isinstance(loader, tud.DataLoader) and loader.batch_size and \
            isinstance(loader.sampler, tud.SequentialSampler):
        n, bs = len(loader.dataset), loader.batch_size
        nb = limit(n // bs if loader.drop_last else -(-n // bs))
        sizes = [min(bs, n - i * bs) for i in range(nb)]
        lo_b, hi_b = rdist.shard_range(nb, rank, world)
        if world == 1:
            own = (v for i, v in enumerate(loader) if i < nb)
        else:
            sub = tud.Subset(loader.dataset, range(lo_b * bs, min(hi_b * bs, n, sum(sizes))))
            own = tud.DataLoader(sub, batch_size=bs, shuffle=False, num_workers=loader.num_workers,
                                 collate_fn=loader.collate_fn, pin_memory=loader.pin_memory,
                                 worker_init_fn=loader.worker_init_fn) if hi_b > lo_b else []
        return sizes, lo_b, own
    if hasattr(loader, "__len__"):
        nb = limit(len(loader))
        lo_b, hi_b = rdist.shard_range(nb, rank, world)
        sizes, mine = [], []
        for i, vals in enumerate(loader):
            if i >= nb:
                break
            sizes.append(vals[0].shape[0])
            if lo_b <= i < hi_b:
                mine.append((vals[0], vals[1]))
        return sizes, lo_b, mine
    batches = [(v[0], v[1]) for i, v in enumerate(loader) if n_batches is None or n_batches < 0 or i < n_batches]
    lo_b, hi_b = rdist.shard_range(len(batches), rank, world)
    return [b[0].shape[0] for b in batches], lo_b, batches[lo_b:hi_b]


def run_sea(model, loader, n_cls, eps=8.0 / 255.0, n_iter=300, weights=None, losses=LOSSES,
            n_batches=-1, device="cuda", keep_adv=False, group=None, shard=True, seed=None):
    """Returns a dict: ``clean`` and per-loss ``{mAcc,aAcc,mIoU}``, ``worst_Acc``,
    ``worst_Acc_indiv`` [A], ``final_miou``, ``n_images`` (+ ``x_adv`` per loss if keep_adv).

    ``seed``: re-seed torch's generator with ``seed + batch_index`` before each batch, which makes
    the random starts -- and therefore every counter -- independent of how the batches are sharded
    over ranks.  ``shard=False`` ignores an initialised process group."""
    model.eval()
    dev = torch.device(device)
    world = dist.get_world_size(group) if shard and dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    sizes, lo_b, own = _own_batches(loader, rank, world, n_batches)
    n_total = sum(sizes)
    img_lo = sum(sizes[:lo_b])
    A = len(losses)
    per_loss = [[] for _ in range(A)]   # per batch [3, B, C] counters for every attack
    clean_cnt = []
    advs = [[] for _ in range(A)]
    for bi, vals in enumerate(own, start=lo_b):
        x, y = vals[0], vals[1]
        if seed is not None:
            torch.manual_seed(seed + bi)
        x = x.to(dev, non_blocking=True)
        y = y.to(dev, non_blocking=True)
        with torch.no_grad():
            out = model(x)
        if out.dtype not in (torch.float32, torch.bfloat16):
            out = out.float()
        c0 = ops.loss_fwd_bwd(out, y, "argmax", want_grad=False, want_stats=False, want_counts=True).counts
        clean_cnt.append(c0.permute(1, 0, 2))  # [B,3,C] -> [3,B,C]
        del out
        for a, loss in enumerate(losses):
            x_adv, _, _, cnt = attacker.apgd_largereps(
                model, x, y, weights, norm="Linf", eps=eps, n_iter=n_iter, loss=loss, track_loss="ce-avg",
                use_rs=True, early_stop=True, num_classes=n_cls, return_counts=True)
            per_loss[a].append(cnt.permute(1, 0, 2))
            if keep_adv:
                advs[a].append(x_adv)
    zero = torch.zeros((3, 0, n_cls), dtype=torch.int64, device=dev)
    local = torch.stack([torch.cat(pl, 1) if pl else zero for pl in per_loss] +
                        [torch.cat(clean_cnt, 1) if clean_cnt else zero], 1)   # [3, A+1, n_local, C]
    if world > 1:
        inter, tgt, prd, _ = rdist.allreduce_counters(n_total, img_lo, local[0], local[1], local[2], group=group)
    else:
        inter, tgt, prd = local[0], local[1], local[2]
    res = {"n_images": n_total, "clean": _finalize(inter[A].sum(0), tgt[A].sum(0), prd[A].sum(0))}
    for a, loss in enumerate(losses):
        res[loss] = _finalize(inter[a].sum(0), tgt[a].sum(0), prd[a].sum(0))
    acc_an, _ = ops.sea_worst_acc(inter[:A].contiguous(), tgt[:A].contiguous())
    acc_an = acc_an.cpu()
    res["worst_Acc"] = acc_an.min(0)[0].mean().item()
    res["worst_Acc_indiv"] = acc_an.mean(-1)
    union = (tgt + prd - inter)[:A]
    random.seed(SEED)
    res["final_miou"], res["selected_attack"] = greedy_worst_miou(inter[:A].cpu().numpy(), union.cpu().numpy())
    if keep_adv:
        res["x_adv"] = {loss: torch.cat(advs[a]) if advs[a] else None for a, loss in enumerate(losses)}
    return res
