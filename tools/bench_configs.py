"""Measurements of the BASELINE.json configs other than the default bench workload (c1, c3, c4, c5).
Prints one JSON line per measurement; results are copied into profiles/ and quoted in DESIGN.md.
Usage: python tools/bench_configs.py [c1] [c3] [c4] [c5] [--c4-points N]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench_workloads as W  # noqa: E402
import ikd_ctypes as I  # noqa: E402
import ref_ctypes as R  # noqa: E402


def out(**kw):
    print(json.dumps(kw), flush=True)


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def c1():
    P = W.uniform_cloud(100_000, -5, 5, 1)
    Q = W.uniform_cloud(10_000, -5, 5, 11)
    boxes = W.demo_boxes(4, 1.5, -5, 5, 12)
    t = I.Tree(0.3, 0.6, 0.2)
    tb = timed(lambda: t.build(P))
    tk = timed(lambda: t.knn(Q, 5))
    t0 = time.perf_counter()
    nd = t.delete_boxes(boxes)
    td = time.perf_counter() - t0
    r = R.RefTree(0.3, 0.6, 0.2, serial=False)
    t0 = time.perf_counter(); r.build(P); rb = time.perf_counter() - t0
    t0 = time.perf_counter(); r.knn(Q, 5, nthreads=0, want_points=False); rk = time.perf_counter() - t0
    t0 = time.perf_counter(); rd_n = r.delete_boxes(boxes); rd = time.perf_counter() - t0
    out(config="c1 demo-style 100k/10k", build_ms=tb * 1e3, knn10k_ms=tk * 1e3, delete4boxes_ms=td * 1e3, deleted=nd,
        ref_build_ms=rb * 1e3, ref_knn10k_ms=rk * 1e3, ref_delete_ms=rd * 1e3, ref_deleted=rd_n, ref_threads=r.num_threads())
    t.close(); r.close()


def c3(n=10_000_000, nq=100_000):
    P = W.uniform_cloud(n, -50, 50, 3)
    c, rad, boxes = W.range_queries(nq, -50, 50, 0.5, 5.0, 33)
    t = I.Tree()
    tb = timed(lambda: t.build(P), reps=1)
    res = {}
    def box():
        off, ids = t.box_search(boxes)
        res["box_total"] = int(off[-1])
    def radius():
        off, ids = t.radius_search(c, rad)
        res["rad_total"] = int(off[-1])
    tbx = timed(box, reps=2)
    trd = timed(radius, reps=2)
    # device part alone (count + scan + fill; offsets come back, ids stay on the GPU) and fetch into page-locked memory
    off = np.empty(nq + 1, dtype=np.int64)
    radf = np.ascontiguousarray(rad, dtype=np.float32)
    def box_dev():
        assert t.L.ikd_box_search_batch(t.h, boxes.ctypes.data, nq, off.ctypes.data) == 0
    def rad_dev():
        assert t.L.ikd_radius_search_batch(t.h, c.ctypes.data, radf.ctypes.data, nq, off.ctypes.data) == 0
    tbx_dev = timed(box_dev, reps=3)
    pin = torch.empty(res["box_total"], dtype=torch.int32).pin_memory()
    def box_pinned():
        box_dev()
        assert t.L.ikd_search_fetch(t.h, pin.data_ptr(), res["box_total"]) == 0
    tbx_pin = timed(box_pinned, reps=3)
    trd_dev = timed(rad_dev, reps=3)
    # reference on a subsample of the queries
    r = R.RefTree(serial=False)
    t0 = time.perf_counter(); r.build(P); rb = time.perf_counter() - t0
    m = 300
    t0 = time.perf_counter()
    tot = 0
    for i in range(m):
        tot += len(r.box_search(boxes[i], cap=1 << 16))
    rbox = (time.perf_counter() - t0) / m
    t0 = time.perf_counter()
    for i in range(m):
        r.radius_search(c[i], rad[i], cap=1 << 16)
    rrad = (time.perf_counter() - t0) / m
    out(config=f"c3 box/radius {n} pts, {nq} queries", build_s=tb, box_s=tbx, box_results=res["box_total"],
        box_queries_per_s=nq / tbx, box_points_per_s=res["box_total"] / tbx, radius_s=trd, radius_results=res["rad_total"],
        radius_queries_per_s=nq / trd, ref_build_s=rb, ref_box_queries_per_s_1thread=1 / rbox,
        ref_radius_queries_per_s_1thread=1 / rrad,
        box_device_s=tbx_dev, box_device_queries_per_s=nq / tbx_dev, box_device_points_per_s=res["box_total"] / tbx_dev,
        box_pinned_fetch_s=tbx_pin, box_pinned_queries_per_s=nq / tbx_pin, radius_device_s=trd_dev,
        radius_device_queries_per_s=nq / trd_dev,
        note="box_s / radius_s include the copy of all result ids into pageable numpy arrays; *_device_* = search only "
             "(ids stay on the GPU); box_pinned_* = search + fetch into page-locked memory; reference single thread, 300-query sample")
    t.close(); r.close()


def c4(n=100_000_000, nq=100_000_000):
    dev = torch.device("cuda")
    P = W.uniform_cloud(n, -100, 100, 4)
    t = I.Tree()
    t0 = time.perf_counter(); t.build(P); tb = time.perf_counter() - t0
    del P
    st = t.stats()
    CH = 25_000_000
    g = torch.Generator(device=dev).manual_seed(4000)
    q4 = torch.zeros((CH, 4), dtype=torch.float32, device=dev)
    q4[:, :3] = torch.rand((CH, 3), generator=g, device=dev) * 200 - 100
    for k in (5, 32):
        oi = torch.empty((CH, k), dtype=torch.int32, device=dev)
        od = torch.empty((CH, k), dtype=torch.float32, device=dev)
        oc = torch.empty(CH, dtype=torch.int32, device=dev)
        t.set_kernel_timing(True)
        t.knn_dev(q4.data_ptr(), CH, k, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr()); t.synchronize()
        t.kernel_time()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        reps = nq // CH
        for _ in range(reps):
            t.knn_dev(q4.data_ptr(), CH, k, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr())
        t.synchronize(); dt = time.perf_counter() - t0
        kms, kn = t.kernel_time()
        t.set_kernel_timing(False)
        t.set_visit_counting(True)
        m = 2_000_000
        t.knn_dev(q4.data_ptr(), m, k, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr()); t.synchronize()
        V = t.stats()["last_knn_visits"] / m
        t.set_visit_counting(False)
        bpq = 12 + 8 * k + 64 * V
        out(config=f"c4 large batch {n} pts, {reps * CH} queries (4 x 25M device batches), k={k}", build_s=tb,
            queries_per_s=reps * CH / dt, kernel_only_queries_per_s=reps * CH / (kms * 1e-3), visits_per_query=V,
            algorithmic_bytes_per_query=bpq, achieved_GBs=bpq * reps * CH / (kms * 1e-3) / 1e9,
            frac_of_measured_6551=bpq * reps * CH / (kms * 1e-3) / 1e9 / 6551.4, node_slots=st["node_slots_used"], depth=st["max_depth"])
        del oi, od, oc
    t.close()


def c5(scans=300):
    """Streaming moving map: per scan Delete_Point_Boxes (local map cube) + Add_Points + 5-NN batch."""
    dev = torch.device("cuda")
    world = W.LidarWorld(seed=5, device=dev)
    half = 100.0
    t = I.Tree(0.5, 0.6, 0.5)
    lat_knn, lat_upd, nq_tot = [], [], 0
    prev = None
    built = False
    t_start = time.perf_counter()
    for i in range(scans):
        o, yaw = world.pose(i, 2.0)
        pts = world.voxel_filter(world.scan(o, yaw), 0.25).cpu().numpy().astype(np.float32)
        if not built:
            t.build(pts)
            built = True
            prev = o
            continue
        torch.cuda.synchronize(); t0 = time.perf_counter()
        t.knn(pts, 5, 5.0)
        t1 = time.perf_counter()
        boxes = W.local_map_boxes(o, half, prev)
        if len(boxes):
            t.delete_boxes(boxes)
        t.add_points(pts, True)
        t2 = time.perf_counter()
        prev = o
        lat_knn.append(t1 - t0); lat_upd.append(t2 - t1); nq_tot += len(pts)
    st = t.stats()
    out(config=f"c5 streaming {scans} scans, 200 m local map cube, ds 0.5", scans=len(lat_knn), knn_p50_ms=1e3 * float(np.median(lat_knn)),
        knn_p99_ms=1e3 * float(np.percentile(lat_knn, 99)), update_p50_ms=1e3 * float(np.median(lat_upd)),
        update_p99_ms=1e3 * float(np.percentile(lat_upd, 99)), update_max_ms=1e3 * float(np.max(lat_upd)),
        mean_queries_per_scan=nq_tot / max(len(lat_knn), 1), validnum=t.validnum(), size=t.size(),
        rebuilds_partial=st["rebuilds_partial"], rebuilds_full=st["rebuilds_full"], rebuilt_points=st["rebuilt_points"],
        node_slots_used=st["node_slots_used"], max_depth=st["max_depth"], wall_s=time.perf_counter() - t_start,
        note="host-buffer API (ikd_knn_batch / ikd_delete_boxes / ikd_add_points), scan generation excluded from latencies")
    t.close()


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c3", "c5"]
    npts = 100_000_000
    if "--c4-points" in sys.argv:
        npts = int(sys.argv[sys.argv.index("--c4-points") + 1])
    for w in which:
        {"c1": c1, "c3": c3, "c4": lambda: c4(npts, npts), "c5": c5}[w]()
