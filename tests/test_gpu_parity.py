"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle (C restatement),
the compiled reference when present, the committed golden vectors, and brute force.
Bar: bit-exact for distances, counts and point sets (integer / index work); see SURVEY.md 8(c)."""
import numpy as np
import pytest

from conftest import replay_delete_by_point, rows, same_set
import ref_ctypes as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def I():
    import ikd_ctypes
    ikd_ctypes.load()
    return ikd_ctypes


def cpu_trees(params):
    ts = [R.OracleTree(*params)]
    if R.available():
        ts.append(R.RefTree(*params))
    return ts


def cloud(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


# ------------------------------------------------------------------------------------------------- golden
def test_golden_vectors(I, golden, built_libs):
    """The committed outputs of the unmodified reference, reproduced by the CUDA path."""
    g = golden
    t = I.Tree(*[float(x) for x in g["params"]])
    t.build(g["P"])
    assert np.array_equal(t.dump_tree()[:, :15], g["build_dump"][:, :15]), "tree structure after Build differs from the reference"
    assert np.array_equal(t.tree_range(), g["build_range"])

    def check(tag):
        for k, md in ((5, np.inf), (1, np.inf), (12, np.inf), (5, 0.5)):
            idx, d, c = t.knn(g["Q"], k, md)
            assert np.array_equal(d, g[f"{tag}_knn_k{k}_md{md}_d"]), (tag, k, md)
            assert np.array_equal(c, g[f"{tag}_knn_k{k}_md{md}_c"])
        off, ids = t.box_search(g["boxes"])
        assert np.array_equal(off, g[f"{tag}_box_off"])
        pts = t.get_points(ids)
        for i in range(len(g["boxes"])):
            assert np.array_equal(rows(pts[off[i]:off[i + 1]]), g[f"{tag}_box_pts"][off[i]:off[i + 1]])
        off, ids = t.radius_search(g["ctr"], g["half"])
        return off, t.get_points(ids)

    off, pts = check("s0")
    # radius search is set-exact right after Build (same tree shape, same shortcut)
    assert np.array_equal(off, g["s0_rad_off"])
    for i in range(len(g["ctr"])):
        assert np.array_equal(rows(pts[off[i]:off[i + 1]]), g["s0_rad_pts"][off[i]:off[i + 1]])
    assert t.delete_boxes(g["del_boxes"]) == int(g["del_boxes_count"])
    t.delete_points(g["del_pts"])
    assert t.validnum() == int(g["s1_validnum"])
    check("s1")
    assert t.add_points(g["add1"], False)[0] == int(g["add1_ret"])
    assert t.add_points(g["add2"], True)[0] == int(g["add2_ret"])
    assert t.validnum() == int(g["s2_validnum"])
    assert np.array_equal(rows(t.get_points(t.flatten())), g["s2_valid_set"])
    check("s2")
    t.close()


# ------------------------------------------------------------------------------------------------- build
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 31, 1000, 65537])
def test_build_structure_matches_oracle(I, built_libs, n):
    P = cloud(n, -5, 5, 100 + n)
    t = I.Tree(0.5, 0.6, 0.2)
    t.build(P)
    for o in cpu_trees((0.5, 0.6, 0.2)):
        o.build(P)
        assert np.array_equal(t.dump_tree()[:, :15], o.dump_tree()[:, :15])
        assert t.size() == o.size() == n and t.validnum() == n
        assert np.array_equal(t.tree_range(), o.tree_range())
        o.close()
    st = t.stats()
    assert st["max_depth"] == int(np.ceil(np.log2(n + 1))) - 1
    t.close()


def test_build_with_duplicate_coordinates(I, built_libs):
    """Ties on the split coordinate: which tied point goes left is unspecified (nth_element, :602-611), so
    below the root the point sets of the subtrees may differ between implementations. What is defined:
    the shape (sizes depend only on n), the root's axis / AABB / split value, the k-d invariants on every
    node, and all query results."""
    rng = np.random.default_rng(5)
    P = np.round(rng.random((5000, 3)) * 8).astype(np.float32)  # heavy duplication on a 9^3 grid
    t = I.Tree()
    t.build(P)
    o = R.OracleTree()
    o.build(P)
    D, E = t.dump_tree(), o.dump_tree()
    assert D.shape == E.shape
    assert np.array_equal(D[:, 4], E[:, 4]) and np.array_equal(D[:, 13:15], E[:, 13:15])  # sizes, child pattern
    ax = int(D[0, 3])
    assert D[0, 3] == E[0, 3] and np.array_equal(D[0, 7:13], E[0, 7:13]) and D[0, ax] == E[0, ax]
    pos = 0

    def rec():
        """returns the points of the subtree; checks split invariant and tight AABB"""
        nonlocal pos
        row = D[pos]
        pos += 1
        a = int(row[3])
        pts = [row[:3]]
        if row[13]:
            L = rec()
            assert np.all(L[:, a] <= row[a])
            pts.append(L)
        if row[14]:
            Rr = rec()
            assert np.all(Rr[:, a] >= row[a])
            pts.append(Rr)
        allp = np.vstack(pts)
        assert np.array_equal(allp.min(axis=0), row[[7, 9, 11]]) and np.array_equal(allp.max(axis=0), row[[8, 10, 12]])
        rg = allp.max(axis=0) - allp.min(axis=0)
        assert a == int(np.argmax(rg))  # largest range, lowest axis on ties (:594-595)
        return allp

    allp = rec()
    assert pos == len(D) and same_set(allp, P)
    Q = cloud(300, 0, 8, 6)
    for k in (1, 5, 40):
        _, d, c = t.knn(Q, k)
        _, d2, c2 = o.knn(Q, k)
        assert np.array_equal(d, d2) and np.array_equal(c, c2)
    bx = np.array([[1, 1, 1, 3, 4, 2.5]], np.float32)
    off, ids = t.box_search(bx)
    assert same_set(t.get_points(ids), o.box_search(bx[0], cap=1 << 16))
    t_deleted_by_box = t.delete_boxes(bx)
    assert t_deleted_by_box == o.delete_boxes(bx)
    # Delete_Points descends one side only: which duplicates are reachable depends on where the build put
    # them, so the expectation is the reference's algorithm replayed on each tree's own dumped structure
    dup = P[:300]
    exp_t = t.validnum() - len(replay_delete_by_point(t.dump_tree(), dup))
    exp_o = o.validnum() - len(replay_delete_by_point(o.dump_tree(), dup))
    t.delete_points(dup)
    o.delete_points(dup)
    assert t.validnum() == exp_t and o.validnum() == exp_o
    t.close()
    o.close()


# ------------------------------------------------------------------------------------------------- kNN
@pytest.mark.parametrize("k", [1, 2, 5, 8, 9, 20, 32, 64, 128])
def test_knn_bit_exact(I, built_libs, k):
    P = cloud(100000, -5, 5, 1)
    Q = cloud(3000, -5.5, 5.5, 2)
    t = I.Tree()
    t.build(P)
    o = R.OracleTree()
    o.build(P)
    for md in (np.inf, 0.35, 0.0):
        idx, d, c = t.knn(Q, k, md)
        _, d2, c2 = o.knn(Q, k, md, nthreads=0, want_points=False)
        assert np.array_equal(d, d2) and np.array_equal(c, c2)
        # ids agree with the distances they are reported with, rows ascending, padding consistent
        pts = t.get_points(np.where(idx >= 0, idx, 0))
        dx, dy, dz = Q[:, None, 0] - pts[:, :, 0], Q[:, None, 1] - pts[:, :, 1], Q[:, None, 2] - pts[:, :, 2]
        dd = ((dx * dx + dy * dy) + dz * dz).astype(np.float32)
        assert np.array_equal(np.where(idx >= 0, dd, np.float32(np.inf)), d)
        assert np.all(np.diff(np.where(np.isinf(d), np.float32(3e38), d), axis=1) >= 0)
        assert np.array_equal((idx >= 0).sum(axis=1), c)
    t.close()
    o.close()


@pytest.mark.parametrize("nq", [200, 3000, 3001, 12000, 12001, 64000, 64001, 270000])
def test_knn_batch_size_classes(I, built_libs, nq):
    """Every launch configuration of the k <= 8 path (cooperative kernel with 32 / 16 / 4 lanes per query, persistent
    one-thread-per-query kernel; counting-sort or radix-sort query ordering) on a tree that has been unbalanced by
    deletes and inserts: distances and counts equal the oracle's bit for bit."""
    P = cloud(120000, -5, 5, 11)
    t = I.Tree(0.5, 0.7, 0.2)
    o = R.OracleTree(0.5, 0.7, 0.2)
    t.build(P)
    o.build(P)
    bx = np.array([[-5, -5, -5, -1, 0, 5], [2, 2, 2, 3.5, 3.5, 3.5]], np.float32)
    assert t.delete_boxes(bx) == o.delete_boxes(bx)
    A = cloud(30000, -2, 6.5, 12)  # partly outside the old bounding box
    t.add_points(A, False)
    o.add_points(A, False)
    assert t.validnum() == o.validnum()
    Q = cloud(nq, -6, 7, 13 + nq)
    for k, md in ((5, np.inf), (5, 0.3), (1, np.inf), (8, 0.5)):
        idx, d, c = t.knn(Q, k, md)
        _, d2, c2 = o.knn(Q, k, md, nthreads=0, want_points=False)
        assert np.array_equal(d, d2) and np.array_equal(c, c2)
        assert np.array_equal((idx >= 0).sum(axis=1), c)
    t.close()
    o.close()


def test_knn_ties_are_broken_by_point_id(I):
    """Exact distance ties at the k-th place: the documented rule is (distance, point id)."""
    P = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [2, 0, 0]], np.float32)
    t = I.Tree()
    t.build(P)
    idx, d, c = t.knn(np.zeros((1, 3), np.float32), 3)
    assert list(idx[0]) == [0, 1, 2] and np.all(d[0] == 1.0)
    idx, d, c = t.knn(np.zeros((1, 3), np.float32), 7)
    assert list(idx[0]) == [0, 1, 2, 3, 4, 5, 6]
    t.close()


def test_knn_edge_cases(I):
    t = I.Tree()
    idx, d, c = t.knn(np.zeros((4, 3), np.float32), 5)  # never built
    assert np.all(c == 0) and np.all(idx == -1) and np.all(np.isinf(d))
    t.build(np.zeros((0, 3), np.float32))
    assert t.size() == 0 and not t.has_root()
    idx, d, c = t.knn(np.zeros((0, 3), np.float32), 5)
    assert idx.shape == (0, 5)
    t.build(np.array([[1, 2, 3]], np.float32))
    idx, d, c = t.knn(np.array([[1, 2, 4], [9, 9, 9]], np.float32), 5, 2.0)
    assert list(c) == [1, 0] and d[0, 0] == 1.0 and idx[0, 0] == 0
    with pytest.raises(I.IkdError):
        t.knn(np.zeros((1, 3), np.float32), 0)
    with pytest.raises(I.IkdError):
        t.knn(np.zeros((1, 3), np.float32), 129)
    t.close()


def test_knn_large_map_properties(I):
    """BASELINE size (1M-point map): properties that do not need the oracle -- every map point is its own
    nearest neighbour at distance 0; a brute-force scan of a query sample agrees bit for bit."""
    import torch
    P = cloud(1_000_000, -50, 50, 4)
    t = I.Tree()
    t.build(P)
    assert t.size() == 1_000_000 and t.stats()["max_depth"] == 19
    sel = np.random.default_rng(1).choice(len(P), 200_000, replace=False)
    idx, d, c = t.knn(P[sel], 1)
    assert np.all(d[:, 0] == 0) and np.all(c == 1)
    assert np.array_equal(t.get_points(idx[:, 0]), P[sel])
    Q = cloud(2048, -50, 50, 5)
    idx, d, c = t.knn(Q, 5)
    Pt = torch.from_numpy(P).cuda()
    for s in range(0, len(Q), 256):
        q = torch.from_numpy(Q[s:s + 256]).cuda()
        dx, dy, dz = q[:, None, 0] - Pt[None, :, 0], q[:, None, 1] - Pt[None, :, 1], q[:, None, 2] - Pt[None, :, 2]
        dd = (dx * dx + dy * dy) + dz * dz
        ref = torch.topk(dd, 5, dim=1, largest=False, sorted=True).values.cpu().numpy()
        assert np.array_equal(ref, d[s:s + 256])
    t.close()


# ------------------------------------------------------------------------------------------------- box / radius
def test_box_and_radius_search(I, built_libs):
    P = cloud(200000, -10, 10, 7)
    t = I.Tree()
    t.build(P)
    cpus = cpu_trees((0.5, 0.6, 0.2))
    for o in cpus:
        o.build(P)
    rng = np.random.default_rng(8)
    nq = 200
    ctr = (rng.random((nq, 3), dtype=np.float32) * 22 - 11).astype(np.float32)
    half = (rng.random(nq, dtype=np.float32) * 2.5 + 0.05).astype(np.float32)
    boxes = np.concatenate([ctr - half[:, None], ctr + half[:, None]], axis=1).astype(np.float32)
    boxes[0] = [-100, -100, -100, 100, 100, 100]   # everything
    boxes[1] = [50, 50, 50, 60, 60, 60]            # nothing
    boxes[2] = [1, 1, 1, 1, 1, 1]                  # empty half-open box
    off, ids = t.box_search(boxes)
    pts = t.get_points(ids)
    assert off[1] - off[0] == len(P) and off[2] == off[1] and off[3] == off[2]
    for i in range(nq):
        m = np.all((P >= boxes[i, :3]) & (P < boxes[i, 3:]), axis=1)
        assert same_set(pts[off[i]:off[i + 1]], P[m]), i
    half[0] = 100.0
    off, ids = t.radius_search(ctr, half)
    pts = t.get_points(ids)
    for i in range(nq):
        for o in cpus:  # same tree shape right after Build -> same shortcut decisions -> same sets
            assert same_set(pts[off[i]:off[i + 1]], o.radius_search(ctr[i], half[i], cap=1 << 20)), i
    off, ids = t.box_search(np.zeros((0, 6), np.float32))
    assert list(off) == [0] and len(ids) == 0
    for o in cpus:
        o.close()
    t.close()


# ------------------------------------------------------------------------------------------------- updates
def test_updates_match_oracle(I, built_libs):
    params = (0.5, 0.6, 0.3)
    P = cloud(100000, -5, 5, 21)
    Q = cloud(2000, -5, 5, 22)
    t = I.Tree(*params)
    t.build(P)
    cpus = cpu_trees(params)
    for o in cpus:
        o.build(P)
    rng = np.random.default_rng(23)

    def agree(tag):
        for o in cpus:
            o.wait_rebuild()
            assert t.validnum() == o.validnum(), tag
            assert same_set(t.get_points(t.flatten()), o.flatten()), tag
            for k, md in ((5, np.inf), (5, 0.5), (1, np.inf), (10, 1.0)):
                _, d, c = t.knn(Q, k, md)
                _, d2, c2 = o.knn(Q, k, md, nthreads=0, want_points=False)
                assert np.array_equal(d, d2) and np.array_equal(c, c2), (tag, k, md)

    boxes = np.concatenate([cloud(6, -5, 5, 24) - 0.7, cloud(6, -5, 5, 24) + 0.7], axis=1).astype(np.float32)
    n = t.delete_boxes(boxes)
    for o in cpus:
        assert o.delete_boxes(boxes) == n
    agree("box delete")
    assert t.delete_boxes(boxes) == 0  # idempotent: nothing left to delete
    for o in cpus:
        assert o.delete_boxes(boxes) == 0
    dp = P[rng.choice(len(P), 700, replace=False)]
    dp = np.concatenate([dp, dp[:50], cloud(20, 20, 30, 25)])  # repeats and points that are not in the tree
    t.delete_points(dp)
    for o in cpus:
        o.delete_points(dp)
    agree("point delete")
    A = cloud(7000, -6, 6, 26)
    r = t.add_points(A, False)
    assert r[0] == 0 and len(r[2]) == len(A) and np.array_equal(r[2], np.arange(len(A)))
    for o in cpus:
        assert o.add_points(A, False) == 0
    agree("add")
    B = cloud(30000, -6, 6, 27)
    r = t.add_points(B, True)
    for o in cpus:
        assert o.add_points(B, True) == r[0]
    agree("downsample add")
    # payload bookkeeping: every source index is valid, re-inserted winners refer to existing ids
    src = r[2]
    assert np.all((src >= 0) & (src < len(B)) | (src < 0))
    pts_new = t.get_points(np.arange(r[1], r[1] + len(src)))
    from_batch = src >= 0
    assert np.array_equal(pts_new[from_batch], B[src[from_batch]])
    assert np.array_equal(pts_new[~from_batch], t.get_points(~src[~from_batch]))
    # the same batch again: every voxel already holds its best point
    r2 = t.add_points(B, True)
    for o in cpus:
        assert o.add_points(B, True) == r2[0]
    agree("downsample add twice")
    for o in cpus:
        o.close()
    t.close()


def test_streaming_demo_workload(I, built_libs):
    """The reference demo's loop (ikd_Tree_demo.cpp:184-286), asserting what the demo only prints."""
    params = (0.3, 0.6, 0.2)
    P = cloud(20000, -5, 5, 31)
    t = I.Tree(*params)
    t.build(P)
    cpus = cpu_trees(params)
    for o in cpus:
        o.build(P)
    rng = np.random.default_rng(32)
    alive = P.copy()
    for it in range(60):
        A = cloud(200, -5, 5, 1000 + it)
        ds = it % 3 == 2
        a = t.add_points(A, ds)[0]
        pick = rng.choice(len(alive), 100, replace=False)
        D = alive[pick]
        t.delete_points(D)
        for o in cpus:
            assert o.add_points(A, ds) == a
            o.delete_points(D)
        alive = np.delete(alive, pick, axis=0)
        if not ds:
            alive = np.concatenate([alive, A])
        else:
            alive = cpus[0].flatten()
        if it % 10 == 5:
            c = cloud(4, -5, 5, 2000 + it)
            bx = np.concatenate([c - 0.75, c + 0.75], axis=1).astype(np.float32)
            n = t.delete_boxes(bx)
            for o in cpus:
                assert o.delete_boxes(bx) == n
            alive = cpus[0].flatten()
        Q = cloud(200, -5, 5, 3000 + it)
        _, d, c = t.knn(Q, 5)
        for o in cpus:
            _, d2, c2 = o.knn(Q, 5, want_points=False)
            assert np.array_equal(d, d2), it
    for o in cpus:
        o.wait_rebuild()
        assert t.validnum() == o.validnum()
        assert same_set(t.get_points(t.flatten()), o.flatten())
    # both criteria hold at quiescence on every node (Criterion_Check, ikd_Tree.cpp:1090)
    D = t.dump_tree()
    check_criteria(D, params[0], params[1])
    for o in cpus:
        o.close()
    t.close()


def check_criteria(D, del_param, bal_param):
    """Walk the pre-order dump and test the alpha criteria and the size bookkeeping on every node."""
    pos = 0

    def rec():
        nonlocal pos
        row = D[pos]
        pos += 1
        size, invalid = int(row[4]), int(row[5])
        pdel = int(row[6]) & 1
        s_l = i_l = s_r = i_r = 0
        if row[13]:
            s_l, i_l = rec()
        if row[14]:
            s_r, i_r = rec()
        assert size == 1 + s_l + s_r and invalid == pdel + i_l + i_r
        if size > 10:
            son = s_l if row[13] else s_r
            assert np.float32(invalid) / np.float32(size) <= np.float32(del_param)
            be = np.float32(son) / np.float32(size - 1)
            assert not (be > np.float32(bal_param) or be < np.float32(1) - np.float32(bal_param))
        return size, invalid

    import sys
    sys.setrecursionlimit(10000)
    rec()
    assert pos == len(D)


def test_add_points_on_empty_tree_and_rebuild_to_empty(I):
    t = I.Tree(0.5, 0.6, 0.5)
    A = cloud(5000, -5, 5, 41)
    r = t.add_points(A, True)  # the reference would dereference null here; we build
    assert t.validnum() == len(r[2]) > 0 and t.has_root()
    n = t.delete_boxes(np.array([[-100, -100, -100, 100, 100, 100]], np.float32))
    assert n == len(r[2]) and t.validnum() == 0
    idx, d, c = t.knn(np.zeros((3, 3), np.float32), 2)
    assert np.all(c == 0)
    r = t.add_points(A[:100], False)
    assert t.validnum() == 100
    idx, d, c = t.knn(A[:100], 1)
    assert np.all(d[:, 0] == 0)
    t.close()


def test_downsample_irregular_points_take_the_exact_slow_path(I, built_libs):
    """Points sitting exactly on voxel faces (where fp32 floor(x/ds)*ds boxes can overlap by an ulp) force the
    batch to be split; the result must still equal the sequential reference."""
    params = (0.5, 0.6, 0.1)
    rng = np.random.default_rng(51)
    base = cloud(3000, -2, 2, 52)
    k = np.round(base / np.float32(0.1)).astype(np.float32)
    edge = (k * np.float32(0.1)).astype(np.float32)                 # on (or an ulp off) voxel faces
    edge2 = np.nextafter(edge, np.float32(-np.inf)).astype(np.float32)
    A = np.concatenate([edge[:1000], edge2[1000:2000], base[2000:]])
    rng.shuffle(A)
    t = I.Tree(*params)
    t.build(cloud(2000, -2, 2, 53))
    o = R.OracleTree(*params)
    o.build(cloud(2000, -2, 2, 53))
    for part in np.array_split(A, 3):
        assert t.add_points(part, True)[0] == o.add_points(part, True)
        assert t.validnum() == o.validnum()
        assert same_set(t.get_points(t.flatten()), o.flatten())
    t.close()
    o.close()


def test_acquire_removed_points(I, built_libs):
    params = (0.5, 0.6, 0.3)
    P = cloud(30000, -5, 5, 61)
    t = I.Tree(*params)
    t.build(P)
    bx = np.array([[-5, -5, -5, 0, 5, 5]], np.float32)  # half of the cube: forces rebuilds that drop the deleted points
    n = t.delete_boxes(bx)
    removed = t.get_points(t.acquire_removed())
    m = np.all((P >= bx[0, :3]) & (P < bx[0, 3:]), axis=1)
    assert n == int(m.sum())
    # every removed point was deleted by the box; rebuilds have dropped at least the fully covered subtrees
    assert len(removed) > 0 and same_set(removed, P[m][np.isin(rows_key(P[m]), rows_key(removed))])
    assert len(t.acquire_removed()) == 0  # cleared by the previous call
    t.close()


def rows_key(a):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
    return a.view([("x", np.float32), ("y", np.float32), ("z", np.float32)]).reshape(-1)


def test_add_point_boxes(I, built_libs):
    """Add_Point_Boxes (ikd_Tree.cpp:492): un-delete lazily deleted points inside boxes. Which deleted points
    still exist depends on the rebuild history, so the exact comparison uses criteria that never fire
    (structure stays the fresh Build's, identical to the reference's)."""
    params = (1.0, 1.0, 0.3)
    P = cloud(60000, -5, 5, 71)
    Q = cloud(1000, -5, 5, 72)
    t = I.Tree(*params)
    t.build(P)
    cpus = cpu_trees(params)
    for o in cpus:
        o.build(P)
    c = cloud(6, -4, 4, 73)
    dele = np.concatenate([c - 1.0, c + 1.0], axis=1).astype(np.float32)
    n = t.delete_boxes(dele)
    dp = P[np.random.default_rng(74).choice(len(P), 400, replace=False)]
    t.delete_points(dp)
    for o in cpus:
        assert o.delete_boxes(dele) == n
        o.delete_points(dp)
    assert t.stats()["rebuilds_partial"] == 0 and t.stats()["rebuilds_full"] == 0
    back = np.concatenate([c[:4] - 0.6, c[:4] + 0.9], axis=1).astype(np.float32)  # partly overlapping re-insert boxes
    t.add_boxes(back)
    for o in cpus:
        o.add_boxes(back)
        o.wait_rebuild()
        assert t.validnum() == o.validnum()
        assert same_set(t.get_points(t.flatten()), o.flatten())
        _, d, cc = t.knn(Q, 5)
        _, d2, c2 = o.knn(Q, 5, want_points=False)
        assert np.array_equal(d, d2) and np.array_equal(cc, c2)
        o.close()
    # property with the default criteria: downsample-deleted points never come back (:772, :779)
    t2 = I.Tree(0.5, 0.6, 0.5)
    t2.build(P)
    A = cloud(20000, -5, 5, 75)
    t2.add_points(A, True)
    before = rows(t2.get_points(t2.flatten()))
    t2.add_boxes(np.array([[-10, -10, -10, 10, 10, 10]], np.float32))
    assert np.array_equal(rows(t2.get_points(t2.flatten())), before)
    t2.close()
    t.close()


def test_host_knn_pipeline_large_batch(I):
    """The two-lane host path (chunks of 1M queries) returns exactly what the device path returns."""
    import torch
    P = cloud(300000, -20, 20, 81)
    Q = cloud(2_300_000, -20, 20, 82)
    t = I.Tree()
    t.build(P)
    idx, d, c = t.knn(Q, 5, 3.0)  # pageable numpy buffers -> staged through pinned lanes
    qd = torch.zeros((len(Q), 4), dtype=torch.float32, device="cuda")
    qd[:, :3] = torch.from_numpy(Q).cuda()
    oi = torch.empty((len(Q), 5), dtype=torch.int32, device="cuda")
    od = torch.empty((len(Q), 5), dtype=torch.float32, device="cuda")
    oc = torch.empty(len(Q), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    t.knn_dev(qd.data_ptr(), len(Q), 5, 3.0, oi.data_ptr(), od.data_ptr(), oc.data_ptr())
    t.synchronize()
    assert np.array_equal(idx, oi.cpu().numpy()) and np.array_equal(d, od.cpu().numpy()) and np.array_equal(c, oc.cpu().numpy())
    # pinned caller buffers are used directly
    hq = torch.from_numpy(Q).pin_memory()
    hi = torch.empty((len(Q), 5), dtype=torch.int32).pin_memory()
    hd = torch.empty((len(Q), 5), dtype=torch.float32).pin_memory()
    hc = torch.empty(len(Q), dtype=torch.int32).pin_memory()
    st = t.L.ikd_knn_batch(t.h, hq.data_ptr(), len(Q), 12, 5, 3.0, hi.data_ptr(), hd.data_ptr(), hc.data_ptr())
    assert st == 0
    assert np.array_equal(hi.numpy(), idx) and np.array_equal(hd.numpy(), d) and np.array_equal(hc.numpy(), c)
    t.close()


def test_downsample_crowded_voxels_and_far_points(I, built_libs):
    """Voxels holding hundreds of new points (the sort-free grouping hands over to the sorted path) and points
    far outside the map (voxel index beyond the packed key range -> wide grouping)."""
    params = (0.5, 0.6, 0.5)
    base = cloud(20000, -5, 5, 91)
    t = I.Tree(*params)
    t.build(base)
    o = R.OracleTree(*params)
    o.build(base)
    crowded = cloud(6000, -1, 1, 92)                       # ~90 new points per 0.5 m voxel
    far = (cloud(500, -1, 1, 93) + np.float32(3.0e6)).astype(np.float32)   # voxel index ~6e6
    for batch in (crowded, np.concatenate([cloud(3000, -5, 5, 94), far])):
        assert t.add_points(batch, True)[0] == o.add_points(batch, True)
        assert t.validnum() == o.validnum()
        assert same_set(t.get_points(t.flatten()), o.flatten())
    Q = cloud(500, -5, 5, 95)
    _, d, c = t.knn(Q, 5)
    _, d2, c2 = o.knn(Q, 5, want_points=False)
    assert np.array_equal(d, d2) and np.array_equal(c, c2)
    t.close()
    o.close()


def test_pinned_caller_buffers_take_the_in_place_path(I, built_libs):
    """Page-locked caller buffers (what FAST-LIO2-style callers would register once) are read in place by the pack
    kernels and receive the results by direct copies: same bits as the pageable path and as the oracle, for packed
    (12-byte) and strided (32-byte, payload-carrying) points."""
    import ctypes as C
    import torch
    params = (0.5, 0.6, 0.4)
    P = cloud(80000, -5, 5, 101)
    t = I.Tree(*params)
    o = R.OracleTree(*params)
    t.build(P)
    o.build(P)
    Q = cloud(20000, -5.5, 5.5, 102)
    k = 5
    ref_idx, ref_d, ref_c = t.knn(Q, k, 0.8)  # pageable numpy buffers
    _, d2, c2 = o.knn(Q, k, 0.8, nthreads=0, want_points=False)
    assert np.array_equal(ref_d, d2) and np.array_equal(ref_c, c2)
    for stride in (12, 32):
        hq = torch.zeros((len(Q), stride // 4), dtype=torch.float32).pin_memory()
        hq[:, :3] = torch.from_numpy(Q)
        hq[:, 3:] = 7.0  # payload the library must skip over
        h_idx = torch.full((len(Q), k), -7, dtype=torch.int32).pin_memory()
        h_d = torch.zeros((len(Q), k), dtype=torch.float32).pin_memory()
        h_c = torch.zeros(len(Q), dtype=torch.int32).pin_memory()
        st = t.L.ikd_knn_batch(t.h, hq.data_ptr(), len(Q), stride, k, 0.8, h_idx.data_ptr(), h_d.data_ptr(), h_c.data_ptr())
        assert st == 0, t.L.ikd_last_error()
        assert np.array_equal(h_d.numpy(), ref_d) and np.array_equal(h_c.numpy(), ref_c) and np.array_equal(h_idx.numpy(), ref_idx)
    # Add_Points from a pinned, strided buffer
    A = cloud(15000, -6, 6, 103)
    ha = torch.zeros((len(A), 8), dtype=torch.float32).pin_memory()
    ha[:, :3] = torch.from_numpy(A)
    added, first, nins = C.c_int(), C.c_int32(), C.c_int64()
    st = t.L.ikd_add_points(t.h, ha.data_ptr(), len(A), 32, 1, C.byref(added), C.byref(first), C.byref(nins), None)
    assert st == 0, t.L.ikd_last_error()
    assert added.value == o.add_points(A, True)
    assert t.validnum() == o.validnum()
    assert same_set(t.get_points(t.flatten()), o.flatten())
    _, d, c = t.knn(Q[:3000], k)
    _, d2, c2 = o.knn(Q[:3000], k, want_points=False)
    assert np.array_equal(d, d2) and np.array_equal(c, c2)
    t.close()
    o.close()


@pytest.mark.parametrize("seed", list(range(1, 25)) + [100, 101])
def test_random_operation_sequences(I, built_libs, seed):
    """Randomised sequences of every mutating call on clustered clouds (so that whole subtrees die, large subtrees go
    to the side-stream rebuild, and inserts land outside the old bounding box); after every step the valid point
    set, validnum, kNN distances and box-search sets equal the oracle's, and the criteria hold on every node at the
    end."""
    rng = np.random.default_rng(1000 + seed)
    params = (float(rng.choice([0.3, 0.5])), float(rng.choice([0.6, 0.7])), float(rng.choice([0.2, 0.5])))
    ext = float(rng.choice([3.0, 8.0]))

    def blob(n):
        c = rng.uniform(-ext, ext, 3)
        s = rng.uniform(0.05, 0.5) * ext
        return (rng.normal(0, 1, (n, 3)) * s + c).astype(np.float32)

    n0 = 400000 if seed >= 100 else int(rng.integers(2000, 30000))  # the large ones reach multi-root side-stream rebuilds
    P = np.concatenate([cloud(n0, -ext, ext, 2000 + seed), blob(int(rng.integers(100, 8000)))])
    t = I.Tree(*params)
    o = R.OracleTree(*params)
    t.build(P)
    o.build(P)
    Q = np.concatenate([cloud(700, -1.2 * ext, 1.2 * ext, 3000 + seed), blob(300)])

    def agree(tag):
        assert t.validnum() == o.validnum(), tag
        assert same_set(t.get_points(t.flatten()), o.flatten()), tag
        for k, md in ((5, np.inf), (3, 0.2 * ext)):
            _, d, c = t.knn(Q, k, md)
            _, d2, c2 = o.knn(Q, k, md, nthreads=0, want_points=False)
            assert np.array_equal(d, d2) and np.array_equal(c, c2), (tag, k, md)
        bx = np.sort(rng.uniform(-ext, ext, (2, 3)), axis=0).reshape(1, 6).astype(np.float32)
        off, ids = t.box_search(bx)
        assert same_set(t.get_points(ids), o.box_search(bx[0], cap=1 << 20)), tag

    for step in range(14):
        op = int(rng.integers(0, 5))
        if op == 0:
            lo = rng.uniform(-ext, ext, (int(rng.integers(1, 5)), 3))
            boxes = np.concatenate([lo, lo + rng.uniform(0.05, 0.8) * ext], axis=1).astype(np.float32)
            if seed >= 100:  # several disjoint, mostly emptied regions at once
                lo = np.array([[-ext, -ext, -ext], [0.1 * ext, 0.1 * ext, 0.1 * ext], [-ext, 0.2 * ext, -ext]])
                boxes = np.concatenate([lo, lo + 0.75 * ext], axis=1).astype(np.float32)
            assert t.delete_boxes(boxes) == o.delete_boxes(boxes), (step, "delete_boxes")
        elif op == 1:
            A = blob(int(rng.integers(1, 9000))) if rng.random() < 0.6 else cloud(int(rng.integers(1, 9000)), -ext, ext, step)
            assert t.add_points(A, True)[0] == o.add_points(A, True), (step, "add ds")
        elif op == 2:
            A = blob(int(rng.integers(1, 5000)))
            t.add_points(A, False)
            o.add_points(A, False)
        elif op == 3:
            cur = o.flatten()
            if len(cur):
                dp = cur[rng.choice(len(cur), min(len(cur), int(rng.integers(1, 600))), replace=False)]
                t.delete_points(dp)
                o.delete_points(dp)
        else:
            # (Add_Point_Boxes is not part of the random mix: which deleted points it can revive depends on whether a
            # rebuild has already dropped them, i.e. on rebuild timing, which differs by design -- it has its own test)
            big = np.array([[-2 * ext, -2 * ext, -2 * ext, 2 * ext, 2 * ext, rng.uniform(-ext, 0.0)]], np.float32)
            assert t.delete_boxes(big) == o.delete_boxes(big), (step, "delete slab")
        agree((seed, step, op))
    D = t.dump_tree()
    if len(D):
        check_criteria(D, params[0], params[1])
    t.close()
    o.close()


def test_degenerate_inputs_do_not_break_anything(I, built_libs):
    """All-identical points, points on a line, huge and tiny coordinates, NaN / inf queries: no hang, no error, and
    wherever the reference's result is well defined (no NaN involved) it is reproduced."""
    params = (0.5, 0.6, 0.3)
    # 1. 30000 copies of one point plus a line of points with two constant coordinates
    same = np.tile(np.array([[1.5, -2.0, 0.25]], np.float32), (30000, 1))
    line = np.stack([np.linspace(-5, 5, 20000, dtype=np.float32), np.full(20000, 3.0, np.float32), np.zeros(20000, np.float32)], axis=1)
    P = np.concatenate([same, line]).astype(np.float32)
    t = I.Tree(*params)
    o = R.OracleTree(*params)
    t.build(P)
    o.build(P)
    assert t.size() == o.size() == len(P)
    Q = np.concatenate([cloud(500, -6, 6, 301), same[:3], line[:50]])
    for k, md in ((5, np.inf), (1, 0.5), (40, np.inf)):
        _, d, c = t.knn(Q, k, md)
        _, d2, c2 = o.knn(Q, k, md, nthreads=0, want_points=False)
        assert np.array_equal(d, d2) and np.array_equal(c, c2)
    bx = np.array([[1.0, -3.0, 0.0, 2.0, -1.0, 1.0], [-1.0, 2.5, -0.5, 1.0, 3.5, 0.5]], np.float32)
    assert t.delete_boxes(bx) == o.delete_boxes(bx)
    assert t.validnum() == o.validnum()
    A = np.concatenate([same[:500] + np.float32(0.01), cloud(2000, -5, 5, 302)]).astype(np.float32)
    assert t.add_points(A, True)[0] == o.add_points(A, True)
    assert t.validnum() == o.validnum()
    assert same_set(t.get_points(t.flatten()), o.flatten())
    # 2. NaN / inf queries and a k larger than the tree: defined shapes, no hang
    bad = np.array([[np.nan, 0, 0], [np.inf, 0, 0], [0, -np.inf, 0], [1e30, 1e30, 1e30], [1e-40, 0, 0]], np.float32)
    idx, d, c = t.knn(bad, 5)
    assert idx.shape == (5, 5) and np.all(c >= 0) and np.all(c <= 5)
    small = I.Tree()
    small.build(cloud(3, -1, 1, 303))
    idx, d, c = small.knn(cloud(10, -1, 1, 304), 8)
    assert np.all(c == 3) and np.all(idx[:, 3:] == -1) and np.all(np.isinf(d[:, 3:]))
    # 3. huge / tiny magnitudes in the map
    wide = np.concatenate([cloud(5000, -1e6, 1e6, 305), cloud(5000, -1e-3, 1e-3, 306)]).astype(np.float32)
    # (checked against brute force in the reference's fp32 operation order, not against the reference: its heap compares
    # by point.x instead of distance when two squared distances differ by less than 1e-10, ikd_Tree.h:102-105, so
    # inside the millimetre-sized cluster it does not return the nearest neighbours)
    t2 = I.Tree(*params)
    t2.build(wide)
    Q2 = np.concatenate([cloud(300, -1e6, 1e6, 307), cloud(300, -2e-3, 2e-3, 308)]).astype(np.float32)
    _, d, c = t2.knn(Q2, 5)
    dx, dy, dz = Q2[:, None, 0] - wide[None, :, 0], Q2[:, None, 1] - wide[None, :, 1], Q2[:, None, 2] - wide[None, :, 2]
    dd = ((dx * dx + dy * dy) + dz * dz).astype(np.float32)
    assert np.array_equal(np.sort(dd, axis=1)[:, :5], d) and np.all(c == 5)
    for x in (t, o, small, t2):
        x.close()
