"""Image-sharded SEA across the GPUs of one box.

Every attacked image is independent, so the validation set is split into contiguous shards,
one per rank (one process per GPU); each rank runs all attacks on its shard with its own
model replica.  The ONLY collective is one ``all_reduce(SUM)`` of a flat int64 buffer at the
end: per-attack confusion matrices ``[A,C,C]`` and the per-image counters ``[A,N,C] x 3``, each
rank writing only its own image slots (zeros elsewhere, so the sum is a gather).  NCCL over
NVLink on the GPU box; the same code runs on ``gloo`` for the CPU tests because an int64 sum
is backend-agnostic and exact.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous [lo, hi) slice of rank; sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_counters(n_total, lo, inter, tgt, prd, hist=None):
    """Place this rank's [A,n_local,C] counters into zero-filled [A,n_total,C] slots and
    flatten everything into one int64 buffer (plus the [A,C,C] confusion matrices)."""
    A, n_local, C = inter.shape
    dev = inter.device
    full = torch.zeros((3, A, n_total, C), dtype=torch.int64, device=dev)
    full[0, :, lo:lo + n_local] = inter
    full[1, :, lo:lo + n_local] = tgt
    full[2, :, lo:lo + n_local] = prd
    parts = [full.reshape(-1)]
    if hist is not None:
        parts.append(hist.to(torch.int64).reshape(-1))
    return torch.cat(parts), (A, n_total, C, hist is not None)


def unpack_counters(buf, meta):
    A, n_total, C, has_hist = meta
    n = 3 * A * n_total * C
    full = buf[:n].view(3, A, n_total, C)
    hist = buf[n:].view(A, C, C) if has_hist else None
    return full[0], full[1], full[2], hist


def allreduce_counters(n_total, lo, inter, tgt, prd, hist=None, group=None):
    """The single collective of image-sharded SEA.  Returns global (inter, tgt, prd, hist)."""
    buf, meta = pack_counters(n_total, lo, inter, tgt, prd, hist)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack_counters(buf, meta)
