import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ikd_ctypes as I
import ref_ctypes as R
def cloud(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
n = 40000; params = (0.99, 0.99, 0.2)
P = cloud(n, -5, 5, 300 + n)
t = I.Tree(*params); t.build(P)
o = R.OracleTree(*params); o.build(P)
A = cloud(400, -5, 5, 301 + n)
t.delete_points(P[:50]); t.add_points(A, False)
o.delete_points(P[:50]); o.add_points(A, False)
print("before box: size", t.size(), o.size(), "valid", t.validnum(), o.validnum(), "alpha", t.root_alpha(), o.root_alpha())
box = np.array([[-5, -5, -5, -1, 5, 5]], np.float32)
print("deleted", t.delete_boxes(box), o.delete_boxes(box))
print("after box (no dump yet): size", t.size(), "valid", t.validnum(), "alpha", t.root_alpha(), "stats", {k: v for k, v in t.stats().items() if "rebuild" in k})
D = t.dump_tree()
print("after dump: size", t.size(), "valid", t.validnum(), "alpha", t.root_alpha())
print("dump rows", len(D), "root row", D[0, 3:7], D[0, 13:16], "row1", D[1, 3:7], "sum exists", len(D))
E = o.dump_tree()
print("oracle: size", o.size(), "alpha", o.root_alpha(), "root", E[0, 3:7], "row1", E[1, 3:7])
# recount sizes from the dump
pos = [0]
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
import sys as _s; _s.setrecursionlimit(100000)
bad = []
print("recount", rec(), "mismatching nodes", len(bad), bad[:5])
