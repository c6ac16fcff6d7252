import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ikd_ctypes as I
import ref_ctypes as R
def cloud(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
sys.setrecursionlimit(100000)
def check(D):
    pos = [0]; bad = []
    def rec():
        i = pos[0]; pos[0] += 1
        s = 1; inv = int(D[i, 6]) & 1
        if D[i, 13]:
            a, b = rec(); s += a; inv += b
        if D[i, 14]:
            a, b = rec(); s += a; inv += b
        if s != D[i, 4] or inv != D[i, 5]:
            bad.append((i, s, D[i, 4], inv, D[i, 5]))
        return s, inv
    tot = rec()
    return tot, bad
n = 40000; params = (0.99, 0.99, 0.2)
P = cloud(n, -5, 5, 300 + n)
A = cloud(400, -5, 5, 301 + n)
box = np.array([[-5, -5, -5, -1, 5, 5]], np.float32)
for trial in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    t = I.Tree(*params); t.build(P)
    if trial % 2:
        o = R.RefTree(*params); o.build(P); o.close()
    t.delete_points(P[:50]); t.add_points(A, False)
    if trial % 3 == 0:
        time.sleep(0.01)
    nd = t.delete_boxes(box)
    a0 = t.root_alpha(); s0 = t.size()
    if trial % 4 >= 2:
        time.sleep(0.005)
    D = t.dump_tree()
    a1 = t.root_alpha()
    tot, bad = check(D)
    size, invalid = np.float32(D[0, 4]), np.float32(D[0, 5])
    son = np.float32(D[1, 4]); tb = son / (size - np.float32(1))
    exp = (float(tb) if float(tb) >= 0.5 - 1e-6 else float(np.float32(1) - tb), float(invalid / size))
    st = t.stats()
    print(trial, "deleted", nd, "size", s0, t.size(), "rows", len(D), "root", D[0, 4:6], "row1", D[1, 4:6], "alpha", a0, a1, "exp", exp, "OK" if a1 == exp else "MISMATCH",
          "recount", tot, "badnodes", len(bad), "async", st["rebuilds_async"], "partial", st["rebuilds_partial"])
    t.close()
