"""Replica synchronisation for query-sharded multi-GPU search.

Every rank holds a full replica of the flat tree; the owner builds it and its arrays are broadcast with
torch.distributed (NCCL over NVLink on a GPU box, gloo in the CPU tests). Queries need no collective:
each rank answers its own contiguous shard. Incremental updates are kept in sync by broadcasting the
update's INPUT (points / boxes) and applying it on every replica (`broadcast_points`), which is cheaper
than shipping node deltas: the update kernels are deterministic in their effect on the point set.
"""
import numpy as np
import torch
import torch.distributed as dist


class _DevMem:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _view(ptr, nbytes, device):
    return torch.as_tensor(_DevMem(ptr, nbytes), device=device)


def shard_range(n, rank, world_size):
    """Contiguous query shard [lo, hi) of rank `rank` (equal sizes, last ranks may get one less / zero)."""
    per = (n + world_size - 1) // world_size
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def broadcast_tree(tree, src, rank, device, chunk_bytes=1 << 30):
    """Make every rank's `tree` a replica of rank `src`'s. Collective: all ranks must call it."""
    meta = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        d = tree.replica_export()
        meta[0], meta[1] = d.slots, d.npoints
    dist.broadcast(meta, src=src)
    slots, npoints = int(meta[0].item()), int(meta[1].item())
    if rank != src:
        d = tree.replica_prepare(slots, npoints)
    for ptr, nbytes in ((d.header_dev, d.header_bytes), (d.search_dev, d.search_bytes), (d.update_dev, d.update_bytes),
                        (d.points_dev, d.points_bytes)):
        off = 0
        while off < nbytes:
            m = min(chunk_bytes, nbytes - off)
            dist.broadcast(_view(ptr + off, m, device), src=src)
            off += m
    torch.cuda.synchronize(device)
    if rank != src:
        tree.replica_commit()
    return slots, npoints


def broadcast_points(arr, src, rank, device, cols=3):
    """Broadcast a float32 [n, cols] host array (an update batch) from `src`; returns it on every rank."""
    n = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        n[0] = len(arr)
    dist.broadcast(n, src=src)
    t = torch.empty((int(n.item()), cols), dtype=torch.float32, device=device)
    if rank == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()
