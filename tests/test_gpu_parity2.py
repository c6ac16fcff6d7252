"""GPU parity tests, second set (-m gpu): the gaps the round-1 review listed.

  * the bench's own configs[1] generator (synthetic LiDAR world, 1M-point map, max_dist 5 m, downsample 0.5 m) against the
    unmodified reference and the C restatement, every kNN distance and the valid set after every scan;
  * acquire_removed_points exactly against the oracle where the reference is deterministic (< 1500 points: every rebuild
    inline, ikd_Tree.h:15), subset / superset invariants on large trees;
  * Radius_Search after updates (tree shape differs from the reference's): superset of brute force inside the radius,
    subset of brute force with the shortcut's slack;
  * root_alpha against the reference's Update (:1315-1321);
  * Delete_Points' one-sided descent replayed on the dumped structure;
  * 10M-point box / radius samples and a 100M-point kNN sample against brute force (torch, chunked).
All comparisons are exact unless a tolerance is written next to them."""
import os
import sys

import numpy as np
import pytest

from conftest import replay_delete_by_point, rows, same_set
import ref_ctypes as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def I():
    import ikd_ctypes
    ikd_ctypes.load()
    return ikd_ctypes


def cloud(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def cpu_trees(params, serial=True):
    ts = [R.OracleTree(*params)]
    if R.available():
        ts.append(R.RefTree(*params, serial=serial))
    return ts


# ------------------------------------------------------------------------------------------------- (i) configs[1]
def test_scan_loop_generator_matches_reference(I, built_libs):
    """bench.py's configs[1] inputs (bench_workloads.LidarWorld seed 2: ground plane + buildings, 64 beams, 1M-point map,
    scans filtered at 0.25 m): Build + 5 scans of [5-NN max_dist 5 m for every scan point, Add_Points(downsample 0.5 m)].
    All squared distances, counts, Add_Points return values, validnum after every scan and the final valid set are
    compared with the reference (one element per call, so that its rebuild thread cannot race) and the oracle."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    dev = torch.device("cuda")
    pmap, steps = bench.make_scanloop_inputs(dev, 1_000_000, 5)
    assert len(pmap) == 1_000_000
    t = I.Tree(*bench.PARAMS)
    t.build(pmap)
    cpus = cpu_trees(bench.PARAMS)
    for o in cpus:
        o.build(pmap)
    for i, (q, a) in enumerate(steps):
        _, d, c = t.knn(q, 5, bench.MAX_DIST)
        added = t.add_points(a, True)[0]
        for o in cpus:
            _, d2, c2 = o.knn(q, 5, bench.MAX_DIST, nthreads=0, want_points=False)
            assert np.array_equal(d.view(np.uint32), d2.view(np.uint32)) and np.array_equal(c, c2), (i, type(o).__name__)
            assert added == o.add_points(a, True), (i, type(o).__name__)
            o.wait_rebuild()
            assert t.validnum() == o.validnum(), (i, type(o).__name__)
    valid = t.get_points(t.flatten())
    for o in cpus:
        assert same_set(valid, o.flatten()), type(o).__name__
        o.close()
    t.close()


# ------------------------------------------------------------------------------------------------- (ii) removed points
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_acquire_removed_points_exact_vs_oracle(I, built_libs, seed):
    """Box deletes / box re-inserts, one box per call, on a 1200-point tree (below Multi_Thread_Rebuild_Point_Num, so every
    rebuild of the reference is inline and its removed list is reproducible): the removed-point list must equal the
    oracle's after every step (recording rule ikd_Tree.cpp:1338-1347: lazily deleted, not downsample-deleted points of
    a rebuilt subtree). On a larger tree WHEN a subtree is rebuilt differs by design; there the list must hold only
    points that really were deleted, none twice, and everything that is no longer stored in the tree."""
    rng = np.random.default_rng(900 + seed)
    params = (0.3, 0.6, 0.2)
    for n in (1200, int(rng.integers(5000, 30000))):
        P = cloud(n, -5, 5, 950 + seed)
        t = I.Tree(*params)
        t.build(P)
        cpus = cpu_trees(params) if n < 1500 else [R.OracleTree(*params)]
        for o in cpus:
            o.build(P)
        o = cpus[0]
        deleted = set()
        reported = []
        for step in range(6):
            lo = rng.uniform(-5, 4, (3, 3))
            boxes = np.concatenate([lo, lo + rng.uniform(0.5, 4)], axis=1).astype(np.float32)
            before = set(map(tuple, rows(t.get_points(t.flatten()))))
            for b in boxes:
                nd = t.delete_boxes(b[None])
                for c in cpus:
                    assert nd == c.delete_boxes(b[None])
                    c.wait_rebuild()
            for b in boxes[:int(rng.integers(0, 3))]:
                t.add_boxes(b[None])
                for c in cpus:
                    c.add_boxes(b[None])
                    c.wait_rebuild()
            assert t.validnum() == o.validnum()
            after_pts = t.get_points(t.flatten())
            assert same_set(after_pts, o.flatten())
            after = set(map(tuple, rows(after_pts)))
            deleted |= before - after
            got = t.get_points(t.acquire_removed())
            if n < 1500:
                for c in cpus:
                    assert same_set(got, c.acquire_removed()), (seed, n, step, type(c).__name__)
            else:
                o.acquire_removed()
            reported.extend(map(tuple, rows(got)))
            assert set(map(tuple, rows(got))) <= deleted, (seed, n, step)
        assert len(reported) == len(set(reported)), "a removed point was reported twice"
        # what the tree no longer stores at all (size counts deleted nodes that are still there) has been reported
        assert len(reported) >= n - t.size()
        for c in cpus:
            c.close()
        t.close()


# ------------------------------------------------------------------------------------------------- (iii) radius after updates
def test_radius_search_after_updates_brackets_brute_force(I, built_libs):
    """After updates the tree shape is not the reference's, and Radius_Search's whole-subtree shortcut (:1053-1063) decides
    at ULP level from the shape. Contract: every valid point with squared distance <= r^2 (1 - 1e-5) is
    returned, nothing beyond r^2 (1 + 1e-5) is; box results stay exact."""
    rng = np.random.default_rng(77)
    params = (0.5, 0.6, 0.25)
    P = cloud(150000, -10, 10, 78)
    t = I.Tree(*params)
    t.build(P)
    o = R.OracleTree(*params)
    o.build(P)
    for rnd in range(4):
        add = (cloud(20000, -10, 10, 100 + rnd) * np.float32(0.7 + 0.1 * rnd)).astype(np.float32)
        assert t.add_points(add, bool(rnd % 2))[0] == o.add_points(add, bool(rnd % 2))
        lo = rng.uniform(-10, 6, (3, 3))
        boxes = np.concatenate([lo, lo + rng.uniform(1, 4, (3, 1))], axis=1).astype(np.float32)
        assert t.delete_boxes(boxes) == o.delete_boxes(boxes)
        dp = P[rng.choice(len(P), 3000, replace=False)]
        t.delete_points(dp)
        o.delete_points(dp)
        assert t.validnum() == o.validnum()
        V = t.get_points(t.flatten())
        assert same_set(V, o.flatten())
        V64 = V.astype(np.float64)
        nq = 150
        ctr = (rng.random((nq, 3), dtype=np.float32) * 22 - 11).astype(np.float32)
        rad = (rng.random(nq, dtype=np.float32) * 3.0 + 0.05).astype(np.float32)
        rad[0] = 50.0  # everything
        off, ids = t.radius_search(ctr, rad)
        pts = t.get_points(ids)
        for i in range(nq):
            got = set(map(tuple, rows(pts[off[i]:off[i + 1]])))
            assert len(got) == off[i + 1] - off[i], "duplicate in a radius result"
            d2 = ((V64 - ctr[i].astype(np.float64)) ** 2).sum(axis=1)
            r2 = float(rad[i]) ** 2
            inner = set(map(tuple, rows(V[d2 <= r2 * (1 - 1e-5)])))
            outer = set(map(tuple, rows(V[d2 <= r2 * (1 + 1e-5)])))
            assert inner <= got <= outer, (rnd, i, len(inner), len(got), len(outer))
        bx = np.concatenate([ctr - rad[:, None], ctr + rad[:, None]], axis=1).astype(np.float32)[:60]
        off, ids = t.box_search(bx)
        pts = t.get_points(ids)
        for i in range(len(bx)):
            m = np.all((V >= bx[i, :3]) & (V < bx[i, 3:]), axis=1)
            assert same_set(pts[off[i]:off[i + 1]], V[m]), (rnd, i)
    t.close()
    o.close()


# ------------------------------------------------------------------------------------------------- (iv) root_alpha
def test_root_alpha_matches_reference(I, built_libs):
    """root_alpha (ikd_Tree.cpp:148, values from Update :1315-1321) after lazy deletes and after inserts, with criteria loose
    enough that nothing is rebuilt (so TreeSize / invalid_point_num of the root's children are defined by the operations
    alone), and after inline rebuilds on a small tree.
    Right after Build the reference's values are UNINITIALISED memory: Build runs Update on the new root before
    Root_Node points to it (:360-363), so the `root == Root_Node` branch at :1315 is skipped and InitTreeNode (:52-76)
    never sets the two fields. This implementation returns what Update would have computed; that state is therefore
    checked against the formula, not against the reference."""
    for n, params in ((5, (0.5, 0.6, 0.2)), (4, (0.5, 0.6, 0.2)), (1000, (0.99, 0.99, 0.2)), (40000, (0.99, 0.99, 0.2)),
                      (1300, (0.3, 0.6, 0.2))):
        P = cloud(n, -5, 5, 300 + n)
        t = I.Tree(*params)
        t.build(P)
        nleft = (n - 1) // 2  # lower median: the left subtree gets floor((n-1)/2) points
        tb = np.float32(nleft) / np.float32(n - 1)
        exp_bal = float(tb) if float(tb) >= 0.5 - 1e-6 else float(np.float32(1) - tb)
        assert t.root_alpha() == (exp_bal, 0.0), (n, "build")
        cpus = cpu_trees(params)
        for o in cpus:
            o.build(P)
        if n < 100:
            t.delete_points(P[:1])
            for o in cpus:
                o.delete_points(P[:1])
                o.wait_rebuild()
                assert t.root_alpha() == o.root_alpha(), (n, "delete one", type(o).__name__)
                o.close()
            t.close()
            continue
        # point deletes and plain inserts with loose criteria: no rebuild in either implementation, so TreeSize and
        # invalid_point_num of the root and its children are defined by the operations alone
        if params[0] > 0.9:
            A = cloud(min(n // 3, 400), -5, 5, 301 + n)
            t.delete_points(P[:50])
            t.add_points(A, False)
            for o in cpus:
                o.delete_points(P[:50])
                o.add_points(A, False)
                o.wait_rebuild()
                assert t.size() == o.size() and t.validnum() == o.validnum()
                assert t.root_alpha() == o.root_alpha(), (n, "point delete + insert", type(o).__name__)
        # Box deletes: the alphas are ratios of TreeSize values, which count lazily deleted nodes that are still stored.
        # WHEN such nodes are dropped differs by design: the reference flags a fully covered subtree at its root and does
        # not run Criterion_Check inside it (:656-669), this implementation deletes eagerly and its refit evaluates the
        # criteria on every touched node, so an entirely deleted subtree of more than 10 nodes is dropped at once
        # (DESIGN 3). Equal sizes => equal history => the alphas must be equal bit for bit; in every case they must be
        # Update's formula (:1315-1321) applied to this tree's own root and son.
        box = np.array([[-5, -5, -5, -1, 5, 5]], np.float32)
        nd = t.delete_boxes(box)
        D = t.dump_tree()
        size, invalid = np.float32(D[0, 4]), np.float32(D[0, 5])
        assert D[0, 13] or D[0, 14]
        son = np.float32(D[1, 4])  # pre-order: row 1 is the left child, or the right one when there is no left child
        tb = son / (size - np.float32(1))
        exp = (float(tb) if float(tb) >= 0.5 - 1e-6 else float(np.float32(1) - tb), float(invalid / size))
        assert t.root_alpha() == exp, (n, "box delete: formula on own structure")
        for o in cpus:
            assert nd == o.delete_boxes(box)
            o.wait_rebuild()
            assert t.validnum() == o.validnum()
            if t.size() == o.size():
                assert t.root_alpha() == o.root_alpha(), (n, "box delete", type(o).__name__)
            o.close()
        t.close()


# ------------------------------------------------------------------------------------------------- (v) one-sided point delete
def test_delete_points_one_sided_descent_with_duplicates(I, built_libs):
    """Heavy duplication: a stored point whose split coordinate equals an ancestor's and that the build placed in the
    left subtree cannot be reached by Delete_by_point. Which duplicates sit left is build-specific
    (nth_element ties), so the expectation is the reference's algorithm replayed on THIS tree's dumped structure; the
    same replay is first pinned on the oracle's own structure and result."""
    rng = np.random.default_rng(5)
    P = np.round(rng.random((5000, 3)) * 8).astype(np.float32)
    dup = P[:300]
    o = R.OracleTree()
    o.build(P)
    Do = o.dump_tree()
    hit_o = replay_delete_by_point(Do, dup)
    o.delete_points(dup)
    assert o.validnum() == 5000 - len(hit_o), "the replay does not describe the oracle"
    assert 0 < len(hit_o) <= 300
    t = I.Tree()
    t.build(P)
    D = t.dump_tree()
    hit = replay_delete_by_point(D, dup)
    t.delete_points(dup)
    assert t.validnum() == 5000 - len(hit)
    D2 = t.dump_tree()
    if len(D2) == len(D):  # no rebuild moved nodes: the very nodes the replay names carry the flag
        newly = np.nonzero((D2[:, 6].astype(int) & 1) & ~(D[:, 6].astype(int) & 1))[0]
        assert sorted(newly.tolist()) == sorted(hit)
    expect = np.delete(D[:, :3], hit, axis=0)
    assert same_set(t.get_points(t.flatten()), expect)
    o.close()
    t.close()


# ------------------------------------------------------------------------------------------------- (vi) big maps vs brute force
def brute_knn_torch(P_dev, Q, k, chunk=4_000_000):
    """k smallest fp32 squared distances per query with the reference's operation order (calc_dist :1374), chunked."""
    import torch
    q = torch.from_numpy(Q).to(P_dev.device)
    best = torch.full((len(Q), k), float("inf"), device=P_dev.device)
    for s in range(0, P_dev.shape[0], chunk):
        p = P_dev[s:s + chunk]
        dx = q[:, 0:1] - p[:, 0][None, :]
        d = dx * dx
        del dx
        dy = q[:, 1:2] - p[:, 1][None, :]
        d = d + dy * dy
        del dy
        dz = q[:, 2:3] - p[:, 2][None, :]
        d = d + dz * dz
        del dz
        cand = torch.cat([best, torch.topk(d, k, dim=1, largest=False).values], dim=1)
        best = torch.topk(cand, k, dim=1, largest=False).values
        del d, cand
    return torch.sort(best, dim=1).values.cpu().numpy()


def test_knn_100M_point_map_sample_vs_brute_force(I):
    """configs[3] size: 100M uniform points in [-100,100)^3 (bench.py's generator and seed); 5-NN and 32-NN of a query
    sample against brute force over all 100M points (bit-exact squared distances), through the host-buffer and the
    device-resident entry points; self-queries return distance 0."""
    import torch
    sys.path.insert(0, ROOT)
    import bench_workloads as W
    n = 100_000_000
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~45 GB of device memory")
    P = W.uniform_cloud(n, -100, 100, 4)
    t = I.Tree()
    t.build(P)
    assert t.size() == n and t.validnum() == n
    dev = torch.device("cuda")
    P_dev = torch.from_numpy(P).to(dev)
    Q = cloud(256, -100, 100, 4001)
    Q[:8] = P[::12_345_678][:8]  # self queries
    for k in (5, 32):
        _, d, c = t.knn(Q, k)
        ref = brute_knn_torch(P_dev, Q, k)
        assert np.array_equal(d.view(np.uint32), ref.view(np.uint32)), k
        assert np.all(c == k) and np.all(d[:8, 0] == 0)
    # device-resident entry point, a batch big enough for the persistent kernel + radix-sort ordering
    m = 300_000
    q4 = torch.zeros((m, 4), dtype=torch.float32, device=dev)
    q4[:, :3] = torch.rand((m, 3), device=dev, generator=torch.Generator(device=dev).manual_seed(9)) * 200 - 100
    q4[:256, :3] = torch.from_numpy(Q).to(dev)
    oi = torch.empty((m, 5), dtype=torch.int32, device=dev)
    od = torch.empty((m, 5), dtype=torch.float32, device=dev)
    oc = torch.empty(m, dtype=torch.int32, device=dev)
    t.knn_dev(q4.data_ptr(), m, 5, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr())
    t.synchronize()
    _, d5, _ = t.knn(Q, 5)
    assert np.array_equal(od[:256].cpu().numpy().view(np.uint32), d5.view(np.uint32))
    # ids are consistent with distances
    ids = oi[:256].cpu().numpy()
    nb = t.get_points(ids.ravel()).reshape(256, 5, 3)
    dd = ((Q[:, None, 0] - nb[:, :, 0]) ** 2 + (Q[:, None, 1] - nb[:, :, 1]) ** 2) + (Q[:, None, 2] - nb[:, :, 2]) ** 2
    assert np.array_equal(dd.astype(np.float32).view(np.uint32), d5.view(np.uint32))
    t.close()


def test_box_radius_10M_point_map_sample_vs_brute_force(I):
    """configs[2] size: 10M uniform points in [-50,50)^3, boxes / radii 0.5-5 m (bench_workloads.range_queries): box results
    exact against brute force; radius results bracket brute force (fresh build: also equal to the reference's, which
    tests on smaller maps hold exactly)."""
    import torch
    sys.path.insert(0, ROOT)
    import bench_workloads as W
    P = W.uniform_cloud(10_000_000, -50, 50, 3)
    c, rad, boxes = W.range_queries(400, -50, 50, 0.5, 5.0, 33)
    t = I.Tree()
    t.build(P)
    dev = torch.device("cuda")
    P_dev = torch.from_numpy(P).to(dev)
    off, ids = t.box_search(boxes)
    pts = t.get_points(ids)
    offr, idsr = t.radius_search(c, rad)
    ptsr = t.get_points(idsr)
    for i in range(len(boxes)):
        lo = torch.from_numpy(boxes[i, :3]).to(dev)
        hi = torch.from_numpy(boxes[i, 3:]).to(dev)
        m = ((P_dev >= lo) & (P_dev < hi)).all(dim=1)
        assert same_set(pts[off[i]:off[i + 1]], P_dev[m].cpu().numpy()), i
        cc = torch.from_numpy(c[i]).to(dev).double()
        d2 = ((P_dev.double() - cc) ** 2).sum(dim=1)
        r2 = float(rad[i]) ** 2
        inner = set(map(tuple, rows(P_dev[d2 <= r2 * (1 - 1e-5)].cpu().numpy())))
        outer = set(map(tuple, rows(P_dev[d2 <= r2 * (1 + 1e-5)].cpu().numpy())))
        got = set(map(tuple, rows(ptsr[offr[i]:offr[i + 1]])))
        assert inner <= got <= outer, i
    t.close()


# ------------------------------------------------------------------------------------------------- ids
def test_id_compaction_keeps_results(I, built_libs):
    """ikd_compact_ids: valid points renumbered 0..M-1 in increasing old-id order, removed log drained with old ids,
    every query result unchanged (as coordinates), Build resets the removed log and the id numbering."""
    params = (0.5, 0.6, 0.3)
    P = cloud(60000, -5, 5, 71)
    t = I.Tree(*params)
    t.build(P)
    o = R.OracleTree(*params)
    o.build(P)
    A = cloud(30000, -5, 5, 72)
    assert t.add_points(A, True)[0] == o.add_points(A, True)
    bx = np.array([[-5, -5, -5, 0.5, 5, 5]], np.float32)
    assert t.delete_boxes(bx) == o.delete_boxes(bx)
    Q = cloud(500, -5, 5, 73)
    _, d0, c0 = t.knn(Q, 5)
    valid_before = t.get_points(t.flatten())
    ids_before = t.flatten()
    xyz_by_old = {int(i): tuple(p) for i, p in zip(ids_before, valid_before)}
    next_before, epoch = t.next_id(), t.id_epoch()
    # (the removed-point log is deliberately left in place: the compaction must drain it)
    old_of_new, removed_old = t.compact_ids()
    assert t.id_epoch() == epoch + 1
    assert t.next_id() == len(old_of_new) == t.validnum() == o.validnum() < next_before
    assert np.all(np.diff(old_of_new) > 0)
    new_xyz = t.get_points(np.arange(len(old_of_new), dtype=np.int32))
    assert all(tuple(new_xyz[i]) == xyz_by_old[int(old_of_new[i])] for i in range(0, len(old_of_new), 97))
    assert len(set(removed_old.tolist())) == len(removed_old) and not (set(removed_old.tolist()) & set(old_of_new.tolist()))
    assert len(t.acquire_removed()) == 0
    assert t.size() == t.validnum()
    _, d1, c1 = t.knn(Q, 5)
    assert np.array_equal(d0, d1) and np.array_equal(c0, c1)
    assert same_set(t.get_points(t.flatten()), valid_before) and same_set(valid_before, o.flatten())
    # updates keep working in the new numbering
    B = cloud(5000, -5, 5, 74)
    added, first, src = t.add_points(B, True)
    assert added == o.add_points(B, True) and first == len(old_of_new)
    assert same_set(t.get_points(t.flatten()), o.flatten())
    # a second Build voids the removed log of the first tree
    assert t.delete_boxes(np.array([[-5, -5, -5, 5, 5, 0]], np.float32)) > 0
    t.build(P[:1000])
    assert len(t.acquire_removed()) == 0 and t.next_id() == 1000
    with pytest.raises(I.IkdError):
        t.get_points(np.array([1000], np.int32))
    t.close()
    o.close()


# --------------------------------------------------------------------------- leaf ids in the walk records (round 2)
@pytest.mark.parametrize("params", [(1.0, 1.0, 0.2), (0.5, 0.6, 0.2)])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 1000, 30000])
def test_range_search_leaf_words_follow_updates(I, built_libs, n, params):
    """Box / radius search report a single-node child from its PARENT's walk record (ikd_node.cuh, WalkRec z / w) and never
    visit it. Every transition of such a leaf -- deleted by point, deleted by box, re-inserted by Add_Point_Boxes, turned into
    an inner node by an insert below it, removed by a rebuild -- must show up in the parent's record: the result set is held
    to the oracle (Search_by_range :1016-1044, Search_by_radius :1047-1087 on a fresh build) and to brute force after every step.
    Which deleted points Add_Point_Boxes can bring back depends on the rebuild history (a rebuild drops them), so that step
    runs only with criteria that never fire (1.0, 1.0); the default criteria cover the rebuilds instead."""
    rng = np.random.default_rng(100 + n)
    P = cloud(n, -4, 4, 200 + n)
    t = I.Tree(*params)
    o = R.OracleTree(*params)
    t.build(P)
    o.build(P)
    boxes = np.array([[-5, -5, -5, 5, 5, 5], [-1, -2, -1.5, 2.5, 1, 3], [0, 0, 0, 4.1, 4.1, 4.1], [-4.1, -4.1, -4.1, 0, 0.5, 0]], np.float32)
    ctr = np.array([[0, 0, 0], [1, -1, 2], [-3, 3, 0]], np.float32)
    rad = np.array([9.0, 2.5, 3.0], np.float32)

    def check(tag):
        assert t.validnum() == o.validnum(), tag
        valid = o.flatten()
        off, ids = t.box_search(boxes)
        pts = t.get_points(ids)
        for i, b in enumerate(boxes):
            got = pts[off[i]:off[i + 1]]
            assert same_set(got, o.box_search(b, cap=1 << 20)), (tag, "box", i)
            inside = np.all((valid >= b[:3]) & (valid < b[3:]), axis=1)
            assert same_set(got, valid[inside]), (tag, "box vs brute force", i)
        off, ids = t.radius_search(ctr, rad)
        pts = t.get_points(ids)
        for i in range(len(ctr)):
            got = pts[off[i]:off[i + 1]]
            d = np.sqrt(((valid.astype(np.float64) - ctr[i].astype(np.float64)) ** 2).sum(axis=1))
            # the bounding-sphere shortcut may admit points a hair outside the radius: bracket
            as_set = lambda a: set(map(tuple, rows(a)))
            assert not (as_set(valid[d <= rad[i] * (1 - 1e-5)]) - as_set(got)), (tag, "radius misses a point", i)
            assert not (as_set(got) - as_set(valid[d <= rad[i] * (1 + 1e-5) + 1e-5])), (tag, "radius reports a far point", i)

    check("fresh")
    m = max(1, n // 3)
    victims = P[rng.choice(n, size=m, replace=False)]
    t.delete_points(victims); o.delete_points(victims)
    check("after point deletes")
    db = np.array([[-1, -1, -1, 1.5, 1, 2]], np.float32)
    assert t.delete_boxes(db) == o.delete_boxes(db)
    check("after box delete")
    if params[0] >= 1.0:
        assert t.stats()["rebuilds_partial"] == 0 and t.stats()["rebuilds_full"] == 0
        t.add_boxes(db); o.add_boxes(db)
        check("after Add_Point_Boxes")
    A = cloud(max(2, n // 2), -4.5, 4.5, 300 + n)
    t.add_points(A, False); o.add_points(A, False)
    check("after plain inserts (leaves become inner nodes)")
    A2 = cloud(max(2, n // 2), -4.5, 4.5, 400 + n)
    assert t.add_points(A2, True)[0] == o.add_points(A2, True)
    check("after downsampled inserts")
    wide = np.array([[-5, -5, -5, 5, 5, 0.2]], np.float32)
    assert t.delete_boxes(wide) == o.delete_boxes(wide)
    check("after a slab delete (rebuilds)")
    t.close()
    o.close()
