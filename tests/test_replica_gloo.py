"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: query sharding and update broadcast."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200"))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import replica_sync as S
    dev = torch.device("cpu")
    n = 1003
    lo, hi = S.shard_range(n, rank, world)
    # every rank answers its own contiguous shard; gather the covered ranges and check the partition
    t = torch.tensor([lo, hi], dtype=torch.int64)
    allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, t)
    cover = sorted((int(a[0]), int(a[1])) for a in allr)
    ok = cover[0][0] == 0 and cover[-1][1] == n and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    # update batches are broadcast from the owner and must arrive bit-identical
    src = np.arange(30, dtype=np.float32).reshape(10, 3) * np.float32(1.1) if rank == 0 else None
    got = S.broadcast_points(src, 0, rank, dev)
    ok = ok and got.shape == (10, 3) and np.array_equal(got, np.arange(30, dtype=np.float32).reshape(10, 3) * np.float32(1.1))
    boxes = np.ones((2, 6), np.float32) if rank == 0 else None
    gb = S.broadcast_points(boxes, 0, rank, dev, cols=6)
    ok = ok and gb.shape == (2, 6)
    # apply_delta: the batch arrives as one [n, 4] / [n, 6] buffer on every rank and is handed to the *_dev entry points
    class FakeTree:
        def __init__(self):
            self.calls = []

        def _grab(self, name, ptr, n, cols):
            import ctypes
            buf = (ctypes.c_float * (n * cols)).from_address(ptr)
            self.calls.append((name, np.frombuffer(buf, dtype=np.float32).reshape(n, cols).copy()))

        def add_points_dev(self, ptr, n, ds):
            self._grab("add_points:%d" % int(bool(ds)), ptr, n, 4)
            return (n, 0, n)

        def delete_points_dev(self, ptr, n):
            self._grab("delete_points", ptr, n, 4)

        def delete_boxes_dev(self, ptr, n):
            self._grab("delete_boxes", ptr, n, 6)
            return 7

        def add_boxes_dev(self, ptr, n):
            self._grab("add_boxes", ptr, n, 6)

    ft = FakeTree()
    pts = np.arange(15, dtype=np.float32).reshape(5, 3)
    bxs = np.arange(12, dtype=np.float32).reshape(2, 6)
    r1 = S.apply_delta(ft, "add_points", pts if rank == 0 else None, 0, rank, dev, downsample_on=True)
    S.apply_delta(ft, "delete_points", pts if rank == 0 else None, 0, rank, dev)
    r3 = S.apply_delta(ft, "delete_boxes", bxs if rank == 0 else None, 0, rank, dev)
    S.apply_delta(ft, "add_boxes", bxs if rank == 0 else None, 0, rank, dev)
    r5 = S.apply_delta(ft, "add_points", np.zeros((0, 3), np.float32) if rank == 0 else None, 0, rank, dev)
    p4 = np.concatenate([pts, np.zeros((5, 1), np.float32)], axis=1)
    ok = ok and r1 == 5 and r3 == 7 and r5 == 0 and [c[0] for c in ft.calls] == ["add_points:1", "delete_points", "delete_boxes", "add_boxes"]
    ok = ok and np.array_equal(ft.calls[0][1], p4) and np.array_equal(ft.calls[1][1], p4)
    ok = ok and np.array_equal(ft.calls[2][1], bxs) and np.array_equal(ft.calls[3][1], bxs)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharding_and_broadcast_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_range_properties():
    import replica_sync as S
    for n in (0, 1, 7, 100, 1000003):
        for w in (1, 2, 4, 8):
            edges = [S.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            assert all(0 <= hi - lo <= (n + w - 1) // w for lo, hi in edges)
