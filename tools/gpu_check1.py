"""Scratch GPU check #1: build + kNN vs the compiled reference and brute force; first timings."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ikd_ctypes as I
import ref_ctypes as R
import torch

def bf_knn(P, Q, k):
    Pt = torch.from_numpy(P).cuda(); out = []
    for i in range(0, len(Q), 256):
        q = torch.from_numpy(Q[i:i+256]).cuda()
        dx = q[:, None, 0] - Pt[None, :, 0]; dy = q[:, None, 1] - Pt[None, :, 1]; dz = q[:, None, 2] - Pt[None, :, 2]
        d = (dx * dx + dy * dy) + dz * dz   # separate roundings (no fma in eager elementwise)
        out.append(torch.topk(d, k, dim=1, largest=False, sorted=True).values.cpu().numpy())
    return np.concatenate(out)

rng = np.random.default_rng(1)
for n, nq in [(1, 10), (2, 10), (3, 10), (7, 50), (1000, 500), (100000, 4000)]:
    P = (rng.random((n, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    Q = (rng.random((nq, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    t = I.Tree(0.3, 0.6, 0.2); t.build(P)
    r = R.RefTree(0.3, 0.6, 0.2); r.build(P)
    print("n", n, "size", t.size(), r.size(), "valid", t.validnum(), "range eq", np.array_equal(t.tree_range(), r.tree_range()), t.stats())
    D1, D2 = t.dump_tree(), r.dump_tree()
    same = D1.shape == D2.shape and np.array_equal(D1[:, :15], D2[:, :15])
    print("  structure identical:", same)
    if not same and D1.shape == D2.shape:
        bad = np.where(~np.all(D1[:, :15] == D2[:, :15], axis=1))[0]
        print("  first diffs", bad[:5], D1[bad[0]], D2[bad[0]])
    for k in (1, 5, 8, 20, 32):
        if k > n: continue
        for md in (float("inf"), 0.35):
            idx, d, c = t.knn(Q, k, md)
            _, d2, c2 = r.knn(Q, k, md)
            ok = np.array_equal(d, d2) and np.array_equal(c, c2)
            # ids consistent with distances
            pts = t.get_points(np.where(idx >= 0, idx, 0))
            dd = ((Q[:, None, 0] - pts[:, :, 0]) ** 2 + (Q[:, None, 1] - pts[:, :, 1]) ** 2) + (Q[:, None, 2] - pts[:, :, 2]) ** 2
            ok2 = np.array_equal(np.where(idx >= 0, dd.astype(np.float32), np.inf).astype(np.float32), d)
            print(f"  k={k} max_dist={md}: dist==ref {ok}, idx consistent {ok2}")
    if n >= 1000:
        b = bf_knn(P, Q[:512], 5); idx, d, c = t.knn(Q[:512], 5)
        print("  brute force (torch) equal:", np.array_equal(b, d))
    t.close(); r.close()

# timing
for n, nq in [(1_000_000, 1_000_000), (10_000_000, 4_000_000)]:
    P = (rng.random((n, 3), dtype=np.float32) * 100 - 50).astype(np.float32)
    Q = (rng.random((nq, 3), dtype=np.float32) * 100 - 50).astype(np.float32)
    t = I.Tree()
    t0 = time.time(); t.build(P); t1 = time.time(); t.build(P); t2 = time.time()
    print(f"build n={n}: first {t1-t0:.3f}s second {t2-t1:.3f}s", t.stats())
    qd = torch.zeros((nq, 4), dtype=torch.float32, device="cuda"); qd[:, :3] = torch.from_numpy(Q).cuda()
    for k in (5, 32):
        oi = torch.empty((nq, k), dtype=torch.int32, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        oc = torch.empty(nq, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for rep in range(3):
            t0 = time.time(); t.knn_dev(qd.data_ptr(), nq, k, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr()); t.synchronize(); dt = time.time() - t0
            print(f"  knn n={n} nq={nq} k={k}: {dt*1e3:.2f} ms -> {nq/dt/1e6:.1f} Mq/s")
    t.set_visit_counting(True)
    oi = torch.empty((nq, 5), dtype=torch.int32, device="cuda"); od = torch.empty((nq, 5), dtype=torch.float32, device="cuda"); oc = torch.empty(nq, dtype=torch.int32, device="cuda")
    t.knn_dev(qd.data_ptr(), nq, 5, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr()); t.synchronize()
    print("  mean visits k=5:", t.stats()["last_knn_visits"] / nq)
    t0 = time.time(); idx, d, c = t.knn(Q, 5); dt = time.time() - t0
    print(f"  e2e host knn: {dt*1e3:.1f} ms -> {nq/dt/1e6:.1f} Mq/s")
    if n == 1_000_000:
        r = R.RefTree(); t0 = time.time(); r.build(P); print("  ref build", time.time() - t0)
        t0 = time.time(); _, d2, c2 = r.knn(Q[:200000], 5, nthreads=0, want_points=False); dt = time.time() - t0
        print(f"  ref knn {r.num_threads()} thr: {200000/dt/1e6:.2f} Mq/s; equal: {np.array_equal(d[:200000], d2)}  ref visits {r.mean_visits(Q[:20000],5):.1f}")
        r.close()
    t.close()
