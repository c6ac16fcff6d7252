"""kNN + plane fit (ikd_knn_plane_batch_dev) for ncu and for a CUDA-event timing of the pair:
1M-point map, 20k queries (scan-sized) and 4M queries (large batch), k=5. Prints per-call milliseconds of
knn_dev alone and of knn_plane_dev, so the plane kernel's share is their difference."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200"))
import ikd_ctypes as I
k = 5
rng = np.random.default_rng(4)
P = (rng.random((1_000_000, 3), dtype=np.float32) * 100 - 50).astype(np.float32)
t = I.Tree(); t.build(P)
st = torch.cuda.ExternalStream(t.stream())
g = torch.Generator(device="cuda").manual_seed(1)
for nq in (20_000, 4_000_000):
    q4 = torch.zeros((nq, 4), dtype=torch.float32, device="cuda"); q4[:, :3] = torch.rand((nq, 3), generator=g, device="cuda") * 100 - 50
    oi = torch.empty((nq, k), dtype=torch.int32, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda"); oc = torch.empty(nq, dtype=torch.int32, device="cuda")
    pl = torch.empty((nq, 4), dtype=torch.float32, device="cuda"); rs = torch.empty(nq, dtype=torch.float32, device="cuda"); vl = torch.empty(nq, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    res = {}
    for name in ("knn", "knn_plane"):
        ts = []
        for it in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            if name == "knn":
                t.knn_dev(q4.data_ptr(), nq, k, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr())
            else:
                t.knn_plane_dev(q4.data_ptr(), nq, k, float("inf"), 5.0, 0.1, pl.data_ptr(), rs.data_ptr(), vl.data_ptr())
            b.record(st)
            t.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b))
        res[name] = float(np.median(ts))
    print(f"nq={nq}: knn {res['knn']:.4f} ms, knn+plane {res['knn_plane']:.4f} ms, plane share {res['knn_plane'] - res['knn']:.4f} ms, "
          f"gate-passing fraction {float((oc == k).float().mean()):.3f}, valid {float(vl.float().mean()):.3f}")
t.close()
