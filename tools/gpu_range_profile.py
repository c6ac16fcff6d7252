"""c3 breakdown: device part (count + scan + fill) vs result fetch for batched box / radius search."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200"))
import ikd_ctypes as I
import bench_workloads as W
import ctypes as C
n, nq = 10_000_000, 100_000
P = W.uniform_cloud(n, -50, 50, 3)
c, rad, boxes = W.range_queries(nq, -50, 50, 0.5, 5.0, 33)
t = I.Tree(); t.build(P)
off = np.empty(nq + 1, dtype=np.int64)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = t.L.ikd_box_search_batch(t.h, boxes.ctypes.data, nq, off.ctypes.data); assert st == 0
    t1 = time.perf_counter()
    total = int(off[-1])
    ids = np.empty(total, dtype=np.int32)
    t2 = time.perf_counter()
    st = t.L.ikd_search_fetch(t.h, ids.ctypes.data, total); assert st == 0
    t3 = time.perf_counter()
    pin = torch.empty(total, dtype=torch.int32).pin_memory() if rep == 0 else pin
    t4 = time.perf_counter()
    st = t.L.ikd_search_fetch(t.h, pin.data_ptr(), total); assert st == 0
    t5 = time.perf_counter()
    print(f"box: search {1e3*(t1-t0):.1f} ms, fetch pageable {1e3*(t3-t2):.1f} ms, fetch pinned {1e3*(t5-t4):.1f} ms, total results {total}")
t.close()
