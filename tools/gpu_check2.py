"""Scratch GPU check #2: box/radius search, deletes, inserts, downsample, streaming vs the compiled reference."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ikd_ctypes as I
import ref_ctypes as R

def rows(a):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
    if len(a) == 0: return a
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]

def same_set(a, b):
    a, b = rows(a), rows(b)
    return a.shape == b.shape and np.array_equal(a, b)

def check_knn(t, r, Q, tag):
    for k, md in [(5, float("inf")), (5, 0.6), (1, float("inf")), (12, 1.0)]:
        idx, d, c = t.knn(Q, k, md); _, d2, c2 = r.knn(Q, k, md, want_points=False)
        ok = np.array_equal(d, d2) and np.array_equal(c, c2)
        print(f"   [{tag}] knn k={k} md={md}: {'OK' if ok else 'MISMATCH'}")
        if not ok:
            bad = np.where(~np.all(d == d2, axis=1))[0]; print("     bad", len(bad), bad[:3], d[bad[0]], d2[bad[0]])

def check_valid(t, r, tag):
    mine = t.get_points(t.flatten()); ref = r.flatten()
    print(f"   [{tag}] valid set equal: {same_set(mine, ref)}  validnum {t.validnum()} vs {r.validnum()}  size {t.size()} vs {r.size()}")

rng = np.random.default_rng(7)
n = 100000
P = (rng.random((n, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
Q = (rng.random((3000, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
t = I.Tree(0.5, 0.6, 0.2); t.build(P)
r = R.RefTree(0.5, 0.6, 0.2); r.build(P)

# box + radius search
nb = 300
ctr = (rng.random((nb, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
half = (rng.random((nb, 1), dtype=np.float32) * 0.9 + 0.1).astype(np.float32)
boxes = np.concatenate([ctr - half, ctr + half], axis=1).astype(np.float32)
off, ids = t.box_search(boxes)
bad = 0
for i in range(nb):
    mine = t.get_points(ids[off[i]:off[i + 1]]); ref = r.box_search(boxes[i], cap=65536)
    bad += 0 if same_set(mine, ref) else 1
print("box search mismatches:", bad, "total results", off[-1])
rad = (rng.random(nb, dtype=np.float32) * 0.9 + 0.1).astype(np.float32)
off, ids = t.radius_search(ctr, rad)
bad = 0
for i in range(nb):
    mine = t.get_points(ids[off[i]:off[i + 1]]); ref = r.radius_search(ctr[i], rad[i], cap=65536)
    bad += 0 if same_set(mine, ref) else 1
print("radius search mismatches:", bad, "total results", off[-1])

# delete boxes
dboxes = boxes[:6].copy()
c1 = t.delete_boxes(dboxes); c2 = r.delete_boxes(dboxes)
print("delete_boxes count", c1, c2, "stats", t.stats())
check_valid(t, r, "after box delete"); check_knn(t, r, Q, "after box delete")
off, ids = t.box_search(boxes[:50]); bad = 0
for i in range(50):
    bad += 0 if same_set(t.get_points(ids[off[i]:off[i + 1]]), r.box_search(boxes[i], cap=65536)) else 1
print("   box search after delete mismatches:", bad)

# delete points
dp = P[rng.choice(n, 500, replace=False)]
t.delete_points(dp); r.delete_points(dp)
check_valid(t, r, "after point delete"); check_knn(t, r, Q, "after point delete")

# add without downsample
A = (rng.random((5000, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
a1 = t.add_points(A, False); a2 = r.add_points(A, False); r.wait_rebuild()
print("add (no ds) returned", a1[0], a2, "first id", a1[1], "src ok", np.array_equal(a1[2], np.arange(5000)))
check_valid(t, r, "after add"); check_knn(t, r, Q, "after add")

# add with downsample
B = (rng.random((20000, 3), dtype=np.float32) * 12 - 6).astype(np.float32)
b1 = t.add_points(B, True); b2 = r.add_points(B, True); r.wait_rebuild()
print("add (ds) returned", b1[0], b2, "ninserted", len(b1[2]), "stats", t.stats())
check_valid(t, r, "after ds add"); check_knn(t, r, Q, "after ds add")

# streaming rounds, demo style
t0 = time.time()
for it in range(30):
    A = (rng.random((2000, 3), dtype=np.float32) * 12 - 6).astype(np.float32)
    ds = bool(it % 2)
    x1 = t.add_points(A, ds)[0]; x2 = r.add_points(A, ds)
    c = (rng.random(3, dtype=np.float32) * 10 - 5); bx = np.concatenate([c - 0.75, c + 0.75]).astype(np.float32)[None]
    y1 = t.delete_boxes(bx); y2 = r.delete_boxes(bx)
    dpts = A[:50]
    t.delete_points(dpts); r.delete_points(dpts)
    if x1 != x2 or y1 != y2: print("   round", it, "add", x1, x2, "del", y1, y2)
r.wait_rebuild()
print("streaming done in", time.time() - t0, "stats", t.stats(), "depth ref", r.max_depth())
check_valid(t, r, "after streaming"); check_knn(t, r, Q, "after streaming")
D = t.dump_tree()
print("dump nodes", len(D), "size", t.size(), "root size col", D[0, 4], "invalid", D[0, 5], "sum pdel", int((D[:, 6].astype(int) & 1).sum()))
# empty-tree add + tiny cases
e = I.Tree(0.5, 0.6, 0.5)
print("empty add:", e.add_points(A[:10], True)[0], e.validnum(), e.size())
idx, d, c = e.knn(Q[:4], 3); print("  knn on tiny", c)
e.close(); t.close(); r.close()
