"""One pass over every major kernel of the library for an ncu --set full capture (see profiles/r01_all_kernels_ncu.txt):
Build (1M points), large and scan-sized 5-NN, 32-NN, box and radius search, box delete, three downsampled Add_Points."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200"))
import ikd_ctypes as I
import bench_workloads as W
P = W.uniform_cloud(1_000_000, -50, 50, 3)
t = I.Tree(0.5, 0.6, 0.5); t.build(P)
dev = torch.device("cuda")
def q4(n, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    q = torch.zeros((n, 4), dtype=torch.float32, device=dev); q[:, :3] = torch.rand((n, 3), generator=g, device=dev) * 100 - 50
    return q
for n, k in ((2_000_000, 5), (20_000, 5), (200_000, 32)):
    q = q4(n, n)
    oi = torch.empty((n, k), dtype=torch.int32, device=dev); od = torch.empty((n, k), dtype=torch.float32, device=dev); oc = torch.empty(n, dtype=torch.int32, device=dev)
    t.knn_dev(q.data_ptr(), n, k, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr()); t.synchronize()
c, rad, boxes = W.range_queries(20000, -50, 50, 0.5, 5.0, 33)
t.box_search(boxes); t.radius_search(c, rad)
t.delete_boxes(boxes[:8])
rng = np.random.default_rng(5)
for i in range(1):
    A = (rng.random((20000, 3), dtype=np.float32) * 60 - 30 + i).astype(np.float32)
    t.add_points(A, True)
t.close()
