"""Replica synchronisation for query-sharded multi-GPU search.

Every rank holds a full replica of the flat tree; the owner builds it and its arrays are broadcast with
torch.distributed (NCCL over NVLink on a GPU box, gloo in the CPU tests). Queries need no collective:
each rank answers its own contiguous shard. Incremental updates are kept in sync by broadcasting the
update's INPUT (points / boxes) and applying it on every replica, which is cheaper than shipping node
deltas: the update kernels are deterministic in their effect on the point set, and -- given identical
replicas and identical inputs -- in the node arrays they produce (tests/test_replica_gpu.py).
`apply_delta` is the device path: the batch is broadcast as a device buffer and handed to the `_dev` entry
points of the C ABI on every rank, with no host bounce. `broadcast_points` (host arrays) remains for the
gloo tests of the host logic.
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


def replica_buffers(d):
    """(device pointer, bytes) of every array that makes up a replica, in broadcast order."""
    return ((d.header_dev, d.header_bytes), (d.search_dev, d.search_bytes), (d.update_dev, d.update_bytes),
            (d.walk_dev, d.walk_bytes), (d.points_dev, d.points_bytes))


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
    for ptr, nbytes in replica_buffers(d):
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


def apply_delta(tree, op, batch, src, rank, device, downsample_on=True):
    """Apply one update batch to every replica (collective). `op` is "add_points" | "delete_points" (batch float32
    [n, 3] or [n, 4]) or "delete_boxes" | "add_boxes" (batch [n, 6]); `batch` is only read on rank `src` (host array or
    device tensor). The batch travels as ONE device buffer (NCCL broadcast) and is applied from device memory by
    ikd_add_points_dev / ikd_delete_points_dev / ikd_delete_boxes_dev / ikd_add_boxes_dev. Returns that call's result."""
    cols = 6 if op in ("delete_boxes", "add_boxes") else 4
    n = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        b = torch.as_tensor(batch, dtype=torch.float32).to(device)
        if cols == 4 and b.shape[1] == 3:
            b = torch.cat([b, torch.zeros((b.shape[0], 1), dtype=torch.float32, device=device)], dim=1)
        b = b.contiguous()
        n[0] = b.shape[0]
    dist.broadcast(n, src=src)
    m = int(n.item())
    if rank != src:
        b = torch.empty((m, cols), dtype=torch.float32, device=device)
    if m:
        dist.broadcast(b, src=src)
    if device.type == "cuda":
        torch.cuda.synchronize(device)  # the tree enqueues on its own stream
    if m == 0:
        return 0
    if op == "add_points":
        return tree.add_points_dev(b.data_ptr(), m, downsample_on)[0]
    if op == "delete_points":
        return tree.delete_points_dev(b.data_ptr(), m)
    if op == "delete_boxes":
        return tree.delete_boxes_dev(b.data_ptr(), m)
    if op == "add_boxes":
        return tree.add_boxes_dev(b.data_ptr(), m)
    raise ValueError(op)
