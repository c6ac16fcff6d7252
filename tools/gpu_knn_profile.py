"""Large-batch kNN run for ncu: 10M-point map, 4M queries, k=5 (and k=32 with arg)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200"))
import ikd_ctypes as I
k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n, nq = 10_000_000, 4_000_000
rng = np.random.default_rng(4)
P = (rng.random((n, 3), dtype=np.float32) * 100 - 50).astype(np.float32)
t = I.Tree(); t.build(P)
g = torch.Generator(device="cuda").manual_seed(1)
q4 = torch.zeros((nq, 4), dtype=torch.float32, device="cuda"); q4[:, :3] = torch.rand((nq, 3), generator=g, device="cuda") * 100 - 50
oi = torch.empty((nq, k), dtype=torch.int32, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda"); oc = torch.empty(nq, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    t.knn_dev(q4.data_ptr(), nq, k, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr()); t.synchronize()
t.close()
