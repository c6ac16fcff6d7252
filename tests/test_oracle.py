"""CPU tests of the oracle (test infrastructure): the C restatement against the committed golden vectors
(outputs of the unmodified reference) and, where oracle/_ref is present, against the reference live."""
import numpy as np
import pytest

from conftest import rows, same_set
import make_golden
import ref_ctypes as R


def test_oracle_reproduces_golden_vectors(golden, built_libs):
    mine = make_golden.scenario(R.OracleTree)
    assert set(mine) == set(golden)
    for key in sorted(golden):
        assert np.array_equal(mine[key], golden[key]), f"oracle differs from the reference's golden vector {key}"


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (reference sources absent)")
def test_reference_still_reproduces_golden_vectors(golden):
    ref = make_golden.scenario(R.RefTree)
    for key in sorted(golden):
        assert np.array_equal(ref[key], golden[key]), key


def brute_knn(P, q, k):
    dx, dy, dz = P[:, 0] - q[0], P[:, 1] - q[1], P[:, 2] - q[2]
    d = (dx * dx + dy * dy) + dz * dz  # fp32, reference operation order (calc_dist, ikd_Tree.cpp:1374)
    return np.sort(d)[:k]


def test_oracle_knn_equals_brute_force(built_libs):
    rng = np.random.default_rng(11)
    P = (rng.random((20000, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    Q = (rng.random((200, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    t = R.OracleTree()
    t.build(P)
    for k in (1, 5, 32):
        _, d, c = t.knn(Q, k)
        assert np.all(c == k)
        for i in range(len(Q)):
            assert np.array_equal(d[i], brute_knn(P, Q[i], k))
    # box search = brute force (half-open boxes, ikd_Tree.cpp:1026)
    b = np.array([-1, -2, -0.5, 1.5, 0.25, 2], dtype=np.float32)
    m = np.all((P >= b[:3]) & (P < b[3:]), axis=1)
    assert same_set(t.box_search(b, cap=65536), P[m])
    assert t.delete_boxes(b[None]) == int(m.sum())
    assert t.validnum() == len(P) - int(m.sum())
    t.close()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_equals_reference_live(seed, built_libs):
    rng = np.random.default_rng(seed)
    P = (rng.random((30000, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    Q = (rng.random((500, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    r, o = R.RefTree(0.5, 0.6, 0.3), R.OracleTree(0.5, 0.6, 0.3)
    r.build(P)
    o.build(P)
    assert np.array_equal(r.dump_tree(), o.dump_tree())  # same structure: axis, point, size, flags, AABB
    for k, md in ((1, np.inf), (5, np.inf), (5, 0.4), (32, np.inf)):
        x1, d1, c1 = r.knn(Q, k, md)
        x2, d2, c2 = o.knn(Q, k, md)
        assert np.array_equal(d1, d2) and np.array_equal(c1, c2) and np.array_equal(np.nan_to_num(x1), np.nan_to_num(x2))
    assert r.mean_visits(Q, 5) == o.mean_visits(Q, 5)
    for _ in range(40):
        c = (rng.random(3, dtype=np.float32) * 10 - 5).astype(np.float32)
        h = np.float32(rng.random() * 0.9 + 0.1)
        b = np.concatenate([c - h, c + h]).astype(np.float32)
        assert np.array_equal(r.box_search(b, cap=65536), o.box_search(b, cap=65536))        # same pre-order
        assert np.array_equal(r.radius_search(c, h, cap=65536), o.radius_search(c, h, cap=65536))
    for it in range(8):
        A = (rng.random((1500, 3), dtype=np.float32) * 12 - 6).astype(np.float32)
        assert r.add_points(A, it % 2 == 0) == o.add_points(A, it % 2 == 0)
        c = (rng.random(3, dtype=np.float32) * 10 - 5).astype(np.float32)
        b = np.concatenate([c - 0.8, c + 0.8]).astype(np.float32)[None]
        assert r.delete_boxes(b) == o.delete_boxes(b)
        r.delete_points(A[:40])
        o.delete_points(A[:40])
    r.wait_rebuild()
    assert r.validnum() == o.validnum()
    assert same_set(r.flatten(), o.flatten())
    _, d1, _ = r.knn(Q, 5)
    _, d2, _ = o.knn(Q, 5)
    assert np.array_equal(d1, d2)
    r.close()
    o.close()


def test_oracle_edge_cases(built_libs):
    t = R.OracleTree(0.5, 0.6, 0.5)
    assert t.size() == 0 and t.validnum() == 0
    _, d, c = t.knn(np.zeros((3, 3), np.float32), 5)
    assert np.all(c == 0) and np.all(np.isinf(d))
    t.build(np.zeros((0, 3), np.float32))
    assert t.size() == 0
    one = np.array([[1, 2, 3]], np.float32)
    t.build(one)
    _, d, c = t.knn(np.array([[1, 2, 4]], np.float32), 5)
    assert c[0] == 1 and d[0, 0] == 1.0
    # duplicates and fewer points than k
    t.build(np.repeat(one, 4, axis=0))
    _, d, c = t.knn(one, 8)
    assert c[0] == 4 and np.all(d[0, :4] == 0)
    # max_dist excludes everything
    _, d, c = t.knn(np.array([[10, 10, 10]], np.float32), 2, 1.0)
    assert c[0] == 0
    t.close()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("seed", list(range(1, 9)))
def test_oracle_equals_reference_on_random_sequences(seed, built_libs):
    """The restatement against the unmodified reference on randomised operation sequences (the same generator shape the
    GPU parity suite uses against the oracle): box / slab / point deletes and plain / downsampled inserts on clustered
    clouds. Compared: return values, validnum, the valid point set, kNN distances and box-search sets after every step
    (tree structure after updates is not comparable: it differs between two runs of the reference itself)."""
    rng = np.random.default_rng(500 + seed)
    params = (float(rng.choice([0.3, 0.5])), float(rng.choice([0.6, 0.7])), float(rng.choice([0.2, 0.5])))
    ext = float(rng.choice([3.0, 8.0]))

    def blob(n):
        c = rng.uniform(-ext, ext, 3)
        s = rng.uniform(0.05, 0.5) * ext
        return (rng.normal(0, 1, (n, 3)) * s + c).astype(np.float32)

    def cloud(n):
        return (rng.random((n, 3), dtype=np.float32) * (2 * ext) - ext).astype(np.float32)

    P = np.concatenate([cloud(int(rng.integers(2000, 20000))), blob(int(rng.integers(100, 5000)))])
    r, o = R.RefTree(*params), R.OracleTree(*params)
    r.build(P)
    o.build(P)
    Q = np.concatenate([cloud(300) * np.float32(1.2), blob(100)])
    for step in range(8):
        op = int(rng.integers(0, 5))
        if op == 0:
            lo = rng.uniform(-ext, ext, (int(rng.integers(1, 5)), 3))
            boxes = np.concatenate([lo, lo + rng.uniform(0.05, 0.8) * ext], axis=1).astype(np.float32)
            assert r.delete_boxes(boxes) == o.delete_boxes(boxes), (step, "delete_boxes")
        elif op == 1:
            A = blob(int(rng.integers(1, 6000))) if rng.random() < 0.6 else cloud(int(rng.integers(1, 6000)))
            assert r.add_points(A, True) == o.add_points(A, True), (step, "add ds")
        elif op == 2:
            A = blob(int(rng.integers(1, 3000)))
            r.add_points(A, False)
            o.add_points(A, False)
        elif op == 3:
            cur = o.flatten()
            if len(cur):
                dp = cur[rng.choice(len(cur), min(len(cur), int(rng.integers(1, 400))), replace=False)]
                r.delete_points(dp)
                o.delete_points(dp)
        else:
            big = np.array([[-2 * ext, -2 * ext, -2 * ext, 2 * ext, 2 * ext, rng.uniform(-ext, 0.0)]], np.float32)
            assert r.delete_boxes(big) == o.delete_boxes(big), (step, "delete slab")
        r.wait_rebuild()
        assert r.validnum() == o.validnum(), (seed, step, op)
        assert same_set(r.flatten(), o.flatten()), (seed, step, op)
        for k, md in ((5, np.inf), (3, 0.2 * ext)):
            _, d, c = r.knn(Q, k, md, nthreads=0, want_points=False)
            _, d2, c2 = o.knn(Q, k, md, nthreads=0, want_points=False)
            assert np.array_equal(d, d2) and np.array_equal(c, c2), (seed, step, op, k)
        bx = np.sort(rng.uniform(-ext, ext, (2, 3)), axis=0).reshape(6).astype(np.float32)
        assert same_set(r.box_search(bx, cap=1 << 20), o.box_search(bx, cap=1 << 20)), (seed, step, op)
    r.close()
    o.close()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_oracle_add_point_boxes_and_removed_points_vs_reference(seed, built_libs):
    """Add_Point_Boxes (:492 -> Add_by_range :763) after box deletes, and acquire_removed_points (:559), restatement
    against the unmodified reference. Box re-insertion is compared exactly (delete counts, validnum, valid set, kNN).
    Removed-point lists are compared exactly on a tree below Multi_Thread_Rebuild_Point_Num (every rebuild is inline in
    both); on larger trees the reference's background thread decides WHEN a subtree is rebuilt, so there the lists are
    only checked to hold points that were really deleted."""
    rng = np.random.default_rng(900 + seed)
    params = (0.3, 0.6, 0.2)
    for n in (1200, int(rng.integers(5000, 30000))):
        P = (rng.random((n, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
        Q = (rng.random((200, 3), dtype=np.float32) * 12 - 6).astype(np.float32)
        r, o = R.RefTree(*params), R.OracleTree(*params)
        r.build(P)
        o.build(P)
        deleted = np.zeros((0, 3), np.float32)
        for step in range(5):
            lo = rng.uniform(-5, 4, (3, 3))
            boxes = np.concatenate([lo, lo + rng.uniform(0.5, 4)], axis=1).astype(np.float32)
            before = o.flatten()
            # one box per call and a wait after each: an operation that runs WHILE the reference's background thread
            # rebuilds a subtree goes through its operation log, and the outcome of that race is not reproducible
            # between two runs of the reference itself (seen: validnum off by one, once in ~40 runs)
            for b in boxes:
                assert r.delete_boxes(b[None]) == o.delete_boxes(b[None])
                r.wait_rebuild()
            for b in boxes[:int(rng.integers(1, 4))]:
                r.add_boxes(b[None])
                o.add_boxes(b[None])
                r.wait_rebuild()
            assert r.validnum() == o.validnum(), (seed, n, step)
            after = o.flatten()
            assert same_set(r.flatten(), after), (seed, n, step)
            _, d, c = r.knn(Q, 5, np.inf, nthreads=0, want_points=False)
            _, d2, c2 = o.knn(Q, 5, np.inf, nthreads=0, want_points=False)
            assert np.array_equal(d, d2) and np.array_equal(c, c2)
            gone = set(map(tuple, rows(before))) - set(map(tuple, rows(after)))
            deleted = np.concatenate([deleted, np.array(sorted(gone), np.float32).reshape(-1, 3)])
            ra, oa = r.acquire_removed(), o.acquire_removed()
            if n < 1500:
                assert same_set(ra, oa), (seed, n, step)
            dset = set(map(tuple, deleted))
            assert set(map(tuple, rows(ra))) <= dset and set(map(tuple, rows(oa))) <= dset
        r.close()
        o.close()


def test_delete_by_point_replay_describes_oracle_and_reference(built_libs):
    """conftest.replay_delete_by_point (used by the GPU tests to state the one-sided-descent expectation) is
    pinned here: replayed on the oracle's / the reference's own structure dump it predicts exactly which nodes
    Delete_Points flags, on a cloud with heavy coordinate duplication."""
    from conftest import replay_delete_by_point
    rng = np.random.default_rng(5)
    P = np.round(rng.random((5000, 3)) * 8).astype(np.float32)
    dup = P[:300]
    trees = [R.OracleTree()] + ([R.RefTree()] if R.available() else [])
    for o in trees:
        o.build(P)
        D = o.dump_tree()
        hit = replay_delete_by_point(D, dup)
        o.delete_points(dup)
        o.wait_rebuild()
        assert o.validnum() == 5000 - len(hit)
        D2 = o.dump_tree()
        if len(D2) == len(D):
            newly = np.nonzero((D2[:, 6].astype(int) & 1) & ~(D[:, 6].astype(int) & 1))[0]
            assert sorted(newly.tolist()) == sorted(hit)
        assert 0 < len(hit) < 300  # some duplicates are unreachable: the quirk is exercised
        o.close()
