import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("IKD_DIR", "/root/repo/ikd-tree_b200")); sys.path.insert(0, "/root/repo/oracle")
import ikd_ctypes as I, ref_ctypes as R
def cloud(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32) * (hi - lo) + lo).astype(np.float32)
P = cloud(120000, -5, 5, 11)
t = I.Tree(0.5, 0.7, 0.2); o = R.OracleTree(0.5, 0.7, 0.2)
t.build(P); o.build(P)
bx = np.array([[-5, -5, -5, -1, 0, 5], [2, 2, 2, 3.5, 3.5, 3.5]], np.float32)
print("del", t.delete_boxes(bx), o.delete_boxes(bx))
A = cloud(30000, -2, 6.5, 12)
t.add_points(A, False); o.add_points(A, False)
print("valid", t.validnum(), o.validnum())
for nq in (200, 3000, 20000):
    Q = cloud(nq, -6, 7, 13 + nq)
    for k, md in ((5, np.inf), (5, 0.3), (1, np.inf), (8, 0.5)):
        idx, d, c = t.knn(Q, k, md)
        _, d2, c2 = o.knn(Q, k, md, nthreads=0, want_points=False)
        bad = np.where((d != d2).any(axis=1) | (c != c2))[0]
        print(nq, k, md, "mismatch rows", len(bad))
        for b in bad[:3]:
            print("  q", Q[b], "ours", d[b], c[b], "ref", d2[b], c2[b])
