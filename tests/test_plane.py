"""Plane fit behind kNN (ikd_knn_plane_batch). CPU part: the oracle's fp32 column-pivoted Householder
QR against float64 least squares within a stated tolerance. GPU part (-m gpu): the CUDA kernel against the oracle,
bit-exact, through the C ABI."""
import numpy as np
import pytest

import ref_ctypes as R

EPS32 = float(np.finfo(np.float32).eps)


def planar_patches(nq, k, seed, spread=0.4, noise=0.01, extent=50.0):
    """nq noisy planar patches of k points each around random centres, plus one query point near each."""
    rng = np.random.default_rng(seed)
    n = rng.normal(size=(nq, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    c = rng.uniform(-extent, extent, size=(nq, 3))
    nbr = np.empty((nq, k, 3))
    for j in range(k):
        v = rng.normal(size=(nq, 3)) * spread
        v -= (v * n).sum(1, keepdims=True) * n
        nbr[:, j] = c + v + n * rng.normal(size=(nq, 1)) * noise
    q = c + rng.normal(size=(nq, 3)) * 0.1
    nbr = nbr.astype(np.float32)
    q = q.astype(np.float32)
    d = ((nbr - q[:, None]) ** 2).sum(2).astype(np.float32)
    o = np.argsort(d, axis=1, kind="stable")
    nbr = np.take_along_axis(nbr, o[:, :, None], axis=1)
    d = np.take_along_axis(d, o, axis=1)
    return q, np.ascontiguousarray(nbr), np.ascontiguousarray(d)


def lstsq64(nbr):
    """float64 reference: solve A n = -1 per patch; returns plane[n,4] and the condition number of A."""
    A = nbr.astype(np.float64)
    k = A.shape[1]
    out = np.empty((len(A), 4))
    cond = np.empty(len(A))
    for i in range(len(A)):
        x = np.linalg.lstsq(A[i], -np.ones(k), rcond=None)[0]
        nn = np.linalg.norm(x)
        out[i, :3] = x / nn
        out[i, 3] = 1.0 / nn
        s = np.linalg.svd(A[i], compute_uv=False)
        cond[i] = s[0] / s[2]
    return out, cond


@pytest.mark.parametrize("k", [3, 5, 8])
def test_oracle_plane_fit_vs_float64_least_squares(k, built_libs):
    """Tolerance: |normal - normal64| <= 8 * cond(A) * eps_fp32 per component (Householder QR is backward stable, so the
    forward error scales with the condition number of the k x 3 system), and the same bound times |q| for the residual."""
    q, nbr, d = planar_patches(3000, k, seed=10 + k)
    cnt = np.full(len(q), k, np.int32)
    pl, res, val = R.plane_batch(q, nbr, d, cnt, max_kth_sqdist=5.0, threshold=0.1)
    pl64, cond = lstsq64(nbr)
    if k == 3:
        cond = np.maximum(cond, 1.0)
    tol = 8.0 * cond * EPS32
    err = np.abs(pl[:, :3] - pl64[:, :3]).max(1)
    assert np.all(err <= tol), (err / tol).max()
    r64 = (pl64[:, :3] * q.astype(np.float64)).sum(1) + pl64[:, 3]
    qn = np.linalg.norm(q.astype(np.float64), axis=1) + 1.0
    assert np.all(np.abs(res - r64) <= 3.0 * tol * qn)
    # validity = every neighbour within the threshold of the fitted plane (checked with the float64 plane, away from the edge)
    dist64 = np.abs((pl64[:, None, :3] * nbr.astype(np.float64)).sum(2) + pl64[:, None, 3]).max(1)
    clear = np.abs(dist64 - 0.1) > 0.01
    assert np.array_equal(val[clear] == 1, dist64[clear] <= 0.1)
    assert val.mean() > 0.9


def test_oracle_plane_fit_vs_lapack_fp32_pivoted_qr(built_libs):
    """The same algorithm from an independent fp32 implementation: LAPACK sgeqp3 (column-pivoted Householder QR) through
    scipy, Q^T b, triangular solve. Two fp32 runs of a backward-stable method differ by at most a few cond(A) * eps."""
    import scipy.linalg as sl
    q, nbr, d = planar_patches(1500, 5, seed=77)
    pl, res, val = R.plane_batch(q, nbr, d, np.full(len(q), 5, np.int32), 5.0, 1e9)
    _, cond = lstsq64(nbr)
    worst = 0.0
    for i in range(len(q)):
        Q, Rm, P = sl.qr(nbr[i], mode="economic", pivoting=True)
        assert Q.dtype == np.float32
        x = sl.solve_triangular(Rm, Q.T @ (-np.ones(5, np.float32)))
        n = np.empty(3, np.float32)
        n[P] = x
        nn = np.linalg.norm(n)
        worst = max(worst, float(np.abs(n / nn - pl[i, :3]).max() / (cond[i] * EPS32)))
        assert abs(1.0 / nn - pl[i, 3]) <= 4.0 * cond[i] * EPS32 * max(1.0, abs(pl[i, 3]))
    assert worst <= 4.0, worst


def test_oracle_plane_gates_and_degenerate_inputs(built_libs):
    q, nbr, d = planar_patches(64, 5, seed=3)
    cnt = np.full(64, 5, np.int32)
    cnt[:8] = 4                       # fewer than k neighbours found
    d2 = d.copy()
    d2[8:16, 4] = 6.0                 # k-th neighbour too far
    nbr2 = nbr.copy()
    nbr2[16:24] = nbr2[16:24, :1]     # five copies of one point: rank 1
    line = np.linspace(0, 1, 5, dtype=np.float32)[None, :, None] * np.float32([1, 2, 3])[None, None, :]
    nbr2[24:32] = nbr2[24:32, :1] + line  # collinear: rank 2 up to rounding -- any plane through the line passes
    pl, res, val = R.plane_batch(q, nbr2, d2, cnt, 5.0, 0.1)
    assert not val[:24].any()
    for i in np.flatnonzero(val[24:32]) + 24:
        assert np.abs((pl[i, :3] * nbr2[i].astype(np.float64)).sum(1) + pl[i, 3]).max() <= 0.1 + 1e-4
    assert np.all(pl[:16] == 0) and np.all(res[:16] == 0)
    assert np.all(np.isfinite(pl)) and np.all(np.isfinite(res))
    assert val[32:].mean() > 0.9
    # an exact plane z = 2 through axis-aligned points: normal (0, 0, -1) or (0, 0, 1), |d| = 2, residual = distance
    P = np.float32([[0, 0, 2], [1, 0, 2], [0, 1, 2], [1, 1, 2], [3, 5, 2]])[None]
    pl, res, val = R.plane_batch(np.float32([[0.5, 0.5, 2.5]]), P, np.zeros((1, 5), np.float32), np.int32([5]), 5.0, 0.1)
    assert val[0] == 1
    assert np.allclose(np.abs(pl[0]), [0, 0, 1, 2], atol=1e-6) and abs(abs(res[0]) - 0.5) < 1e-6


# ------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def I():
    import ikd_ctypes
    ikd_ctypes.load()
    return ikd_ctypes


def world(seed, n_planes=60, per=1500, extent=20.0):
    """Planar patches (walls / floors with 1 cm noise) plus uniform clutter."""
    rng = np.random.default_rng(seed)
    pts = []
    for _ in range(n_planes):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        c = rng.uniform(-extent, extent, size=3)
        u = np.cross(n, [1.0, 0.3, 0.2])
        u /= np.linalg.norm(u)
        v = np.cross(n, u)
        ab = rng.uniform(-3, 3, size=(per, 2))
        pts.append(c + ab[:, :1] * u + ab[:, 1:] * v + n * rng.normal(size=(per, 1)) * 0.01)
    pts.append(rng.uniform(-extent, extent, size=(20000, 3)))
    P = np.concatenate(pts).astype(np.float32)
    Q = (P[rng.choice(len(P), 6000, replace=False)] + rng.normal(size=(6000, 3)) * 0.05).astype(np.float32)
    return P, Q


def check_against_oracle(t, Q, k, max_dist, max_kth, thr):
    pl, res, val, idx = t.knn_plane(Q, k, max_dist, max_kth, thr, want_idx=True)
    idx2, d, cnt = t.knn(Q, k, max_dist)
    assert np.array_equal(idx, idx2)
    nbr = t.get_points(np.where(idx.ravel() < 0, 0, idx.ravel())).reshape(len(Q), k, 3)
    opl, ores, oval = R.plane_batch(Q, nbr, d, cnt, max_kth, thr)
    assert np.array_equal(val, oval)
    assert np.array_equal(pl.view(np.uint32), opl.view(np.uint32)), "plane parameters differ from the oracle (bitwise)"
    assert np.array_equal(res.view(np.uint32), ores.view(np.uint32)), "residuals differ from the oracle (bitwise)"
    return pl, res, val


@pytest.mark.gpu
@pytest.mark.parametrize("k", [3, 4, 5, 6, 7, 8])
def test_gpu_plane_fit_bit_exact(I, k, built_libs):
    P, Q = world(100 + k)
    t = I.Tree(0.5, 0.6, 0.2)
    t.build(P)
    pl, res, val = check_against_oracle(t, Q, k, 5.0, 5.0, 0.1)
    assert 0.3 < val.mean() <= 1.0
    # gates: small search radius (some queries find fewer than k), small k-th distance limit, tight threshold
    check_against_oracle(t, Q, k, 0.08, 5.0, 0.1)
    check_against_oracle(t, Q, k, float("inf"), 0.004, 0.02)
    t.close()


@pytest.mark.gpu
def test_gpu_plane_fit_full_oracle_pipeline_and_updates(I, built_libs):
    """kNN + fit on both sides (oracle tree -> oracle fit), before and after a box delete and a downsampled insert."""
    P, Q = world(7)
    t = I.Tree(0.5, 0.6, 0.3)
    o = R.OracleTree(0.5, 0.6, 0.3)
    t.build(P)
    o.build(P)

    def both():
        pl, res, val = t.knn_plane(Q, 5, 5.0, 5.0, 0.1)
        xyz, d, cnt = o.knn(Q, 5, 5.0, nthreads=0)
        strict = np.all(np.diff(d, axis=1) > 0, axis=1)   # no distance ties: the neighbour order is unique
        opl, ores, oval = R.plane_batch(Q, np.nan_to_num(xyz), d, cnt, 5.0, 0.1)
        assert strict.mean() > 0.99
        assert np.array_equal(val[strict], oval[strict])
        assert np.array_equal(pl[strict].view(np.uint32), opl[strict].view(np.uint32))
        assert np.array_equal(res[strict].view(np.uint32), ores[strict].view(np.uint32))

    both()
    box = np.array([[-5, -5, -5, 6, 4, 5]], np.float32)
    assert t.delete_boxes(box) == o.delete_boxes(box)
    rng = np.random.default_rng(1)
    A = (Q[:3000] + rng.normal(size=(3000, 3)).astype(np.float32) * 0.02).astype(np.float32)
    assert t.add_points(A, True)[0] == o.add_points(A, True)
    both()
    t.close()
    o.close()


@pytest.mark.gpu
def test_gpu_plane_fit_degenerate_and_argument_errors(I, built_libs):
    line = np.linspace(0, 1, 50, dtype=np.float32)[:, None] * np.float32([1, 2, 3])[None, :]
    dup = np.repeat(np.float32([[4, 4, 4]]), 10, axis=0)
    grid = np.stack(np.meshgrid(np.arange(6), np.arange(6), [9.0]), -1).reshape(-1, 3).astype(np.float32)  # exact plane z = 9
    P = np.concatenate([line, dup, grid]).astype(np.float32)
    Q = np.float32([[0.5, 1.0, 1.5], [4, 4, 4.1], [2.2, 2.7, 9.3], [100, 100, 100]])
    t = I.Tree(0.5, 0.6, 0.2)
    t.build(P)
    pl, res, val = check_against_oracle(t, Q, 5, 2.0, 5.0, 0.1)
    assert val[1] == 0 and val[2] == 1 and val[3] == 0   # (the collinear query may get a plane through its line)
    assert np.allclose(np.abs(pl[2]), [0, 0, 1, 9], atol=1e-5) and abs(abs(res[2]) - 0.3) < 1e-5
    assert np.all(np.isfinite(pl)) and np.all(np.isfinite(res))
    for bad_k in (2, 9):
        with pytest.raises(RuntimeError):
            t.knn_plane(Q, bad_k)
    pl, res, val = t.knn_plane(np.zeros((0, 3), np.float32), 5)
    assert pl.shape == (0, 4)
    t.close()


@pytest.mark.gpu
def test_gpu_plane_fit_device_variant_and_float64_tolerance(I, built_libs):
    import torch
    P, Q = world(21)
    t = I.Tree(0.5, 0.6, 0.2)
    t.build(P)
    pl, res, val, idx = t.knn_plane(Q, 5, 5.0, 5.0, 0.1, want_idx=True)
    # device-pointer variant on the tree's stream
    dev = torch.device("cuda:0")
    q4 = torch.zeros((len(Q), 4), dtype=torch.float32, device=dev)
    q4[:, :3] = torch.from_numpy(Q).to(dev)
    dpl = torch.empty((len(Q), 4), dtype=torch.float32, device=dev)
    dres = torch.empty(len(Q), dtype=torch.float32, device=dev)
    dval = torch.empty(len(Q), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    t.knn_plane_dev(q4.data_ptr(), len(Q), 5, 5.0, 5.0, 0.1, dpl.data_ptr(), dres.data_ptr(), dval.data_ptr())
    t.synchronize()
    assert np.array_equal(dpl.cpu().numpy().view(np.uint32), pl.view(np.uint32))
    assert np.array_equal(dres.cpu().numpy().view(np.uint32), res.view(np.uint32))
    assert np.array_equal(dval.cpu().numpy(), val)
    # float64 least squares on the same neighbours: 8 * cond * eps_fp32 per normal component
    sel = np.flatnonzero(val)[:1500]
    nbr = t.get_points(idx[sel].ravel()).reshape(len(sel), 5, 3)
    pl64, cond = lstsq64(nbr)
    assert np.all(np.abs(pl[sel, :3] - pl64[:, :3]).max(1) <= 8.0 * cond * EPS32)
    # more than one chunk through the host path (1M-query chunks)
    rng = np.random.default_rng(5)
    big = P[rng.integers(0, len(P), (1 << 20) + 777)] + np.float32(0.01)
    bpl, bres, bval = t.knn_plane(big, 5, 5.0, 5.0, 0.1)
    tail = big[-2000:]
    tpl, tres, tval = t.knn_plane(tail, 5, 5.0, 5.0, 0.1)
    assert np.array_equal(bpl[-2000:].view(np.uint32), tpl.view(np.uint32)) and np.array_equal(bval[-2000:], tval)
    # page-locked caller buffers: scan-sized calls read / write them in place, large calls copy straight into them
    for qs, ref in ((Q, (pl, res, val, idx)), (big, (bpl, bres, bval, None))):
        n = len(qs)
        hq = torch.from_numpy(np.ascontiguousarray(qs)).pin_memory()
        hpl = torch.empty((n, 4), dtype=torch.float32).pin_memory()
        hrs = torch.empty(n, dtype=torch.float32).pin_memory()
        hvl = torch.empty(n, dtype=torch.uint8).pin_memory()
        hid = torch.empty((n, 5), dtype=torch.int32).pin_memory()
        st = t.L.ikd_knn_plane_batch(t.h, hq.data_ptr(), n, 12, 5, 5.0, 5.0, 0.1, hpl.data_ptr(), hrs.data_ptr(),
                                     hvl.data_ptr(), hid.data_ptr())
        assert st == 0, t.L.ikd_last_error()
        assert np.array_equal(hpl.numpy().view(np.uint32), ref[0].view(np.uint32))
        assert np.array_equal(hrs.numpy().view(np.uint32), ref[1].view(np.uint32))
        assert np.array_equal(hvl.numpy(), ref[2])
        if ref[3] is not None:
            assert np.array_equal(hid.numpy(), ref[3])
        else:
            assert np.array_equal(hid.numpy()[-2000:], t.knn(tail, 5, 5.0)[0])
    t.close()
