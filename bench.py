#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native ikd-Tree hot path.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload largebatch|scanloop]`
prints ONE JSON line on rank 0 (N > 1: launched with torchrun, one rank per GPU).

Default workload `largebatch` = BASELINE.json configs[3], the configuration the metric is quoted on: a 100M-point
uniform map built on rank 0 and broadcast to every rank with NCCL (timed: `replica_broadcast_s`), 100M 5-NN queries
split into contiguous shards, one per rank ("strong" scaling: the total work is fixed). A step = one pass over all
100M queries. The tree (6.4 GB of search records) and the query set are far larger than L2.
  value        : queries/s with the queries already in HBM (CUDA events on the tree's stream, max over ranks)
  e2e          : the same pass through ikd_knn_batch with HOST (page-locked) shards: H2D of the queries and D2H of
                 ids / squared distances / counts inside the timed region
  roofline     : knn_reg_persist_kernel<5>, algorithmic bytes (12 + 8k + 64 V, V = visits of the reference traversal)
                 / kernel time measured live with CUDA events, against MEASURED_PEAKS.json
  cpu_baseline : (N = 1) the UNMODIFIED reference (oracle/_ref) on the host cores: Build of the same 100M-point map,
                 OpenMP Nearest_Search over a 1M-query sample of the same query set; its distances are compared bit for
                 bit with the GPU's for the same queries (a mismatch fails the run)
  scan_loop    : (N = 1) configs[1], the FAST-LIO2 per-scan loop on a 1M-point LiDAR map, as a nested object with its own
                 value / e2e / scan_p50_ms / rooflines (kNN kernel and Add_Points) / cpu_baseline / parity check
  extras       : (N = 1) k = 32 on the same map, configs[2] (box / radius search, 10M points), configs[4] (1000-scan stream)
`--workload scanloop` runs configs[1] alone (N > 1: independent replicas, the loop does not shard).
`--impl reference` times the reference's own CPU implementation on the same config (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

K_NN = 5
MAX_DIST = 5.0
DS = 0.5
SCAN_LEAF = 0.25
MAP_STRIDE = 4.0
PARAMS = (0.5, 0.6, DS)
LB_EXT = 100.0          # configs[3]: uniform map and queries in [-100, 100)^3
LB_MAP_SEED = 4
LB_QUERY_SEED = 4000
LB_SAMPLE = 1_000_000   # queries of the CPU sample = the first LB_SAMPLE queries of the global query set


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="largebatch", choices=["largebatch", "scanloop"])
    ap.add_argument("--map-points", type=int, default=None)
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--k", type=int, default=K_NN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", default="default",
                    help="comma list of scan_loop,k32,c3,c5 | none | default (= all four at N=1 on the full-size workload)")
    ap.add_argument("--c5-scans", type=int, default=1000)
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 5 if a.workload == "largebatch" else 20
    return a


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe). The timed region of the
    default bench is ~10 ms, shorter than one `nvidia-smi -lms` period, so the samples come from NVML (same counters,
    ~0.1 ms per query) on a helper thread; nvidia-smi is the fallback when the NVML binding is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.p = None
        self.nvml = None
        self.samples = []   # (sm_mhz, reasons bitmask)
        self.sm_max = None
        self.stop_flag = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            self.nvml, self.h = self._nvml_handle()
            self.sm_max = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, rs))
            except Exception:
                pass
            time.sleep(0.0005)

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            n = self.nvml
            masks = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, m in masks.items() if any(rs & m for _, rs in self.samples))
            sm = [x for x, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def ncu_traffic(name):
    """DRAM bytes per launch of a kernel from the committed ncu capture summary, if any."""
    d = profile_json(name)
    return d.get("dram_bytes_per_launch") if d else None


class stdout_to_stderr:
    """The reference prints from C (`printf("Multi thread started")`, ikd_Tree.cpp:176); keep stdout to the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def host_mem_available_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2**30
    except Exception:
        return None


class ParityError(RuntimeError):
    pass


def dist_max(x, dev, world_size):
    if world_size == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dist_sum(x, dev, world_size):
    if world_size == 1:
        return int(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(x)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def barrier_sync(world_size):
    import torch
    if world_size > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------------
# configs[1]: FAST-LIO2 scan loop
# ---------------------------------------------------------------------------------------------------
def make_scanloop_inputs(device, n_map, n_steps):
    import bench_workloads as W
    world = W.LidarWorld(seed=2, device=device)
    pmap, nxt = world.build_map(n_map, leaf=DS, scan_stride=MAP_STRIDE, max_scans=6000)
    steps = [world.scan_step(nxt + i, leaf=SCAN_LEAF, scan_stride=MAP_STRIDE, seed=2) for i in range(n_steps)]
    return pmap, steps


def scanloop_config(n_map, k):
    return {"workload": "configs[1] FAST-LIO2 scan loop", "map_points": int(n_map), "k": k, "max_dist_m": MAX_DIST,
            "downsample_m": DS, "scan_filter_leaf_m": SCAN_LEAF, "params": list(PARAMS),
            "l2": "flushed between timed steps (256 MB write, outside the timer)"}


def run_reference_scanloop(args, pmap, steps, warmup, timed):
    """The unmodified reference on the host cores: per step OpenMP Nearest_Search over all queries + Add_Points."""
    with stdout_to_stderr():
        return _run_reference_scanloop(args, pmap, steps, warmup, timed)


def _run_reference_scanloop(args, pmap, steps, warmup, timed):
    import ref_ctypes as R
    t = R.RefTree(*PARAMS, serial=False)
    t.build(pmap)
    nthr = host_threads()  # torchrun exports OMP_NUM_THREADS=1; the baseline uses every core it is allowed to
    times, nq_tot = [], 0
    for i in range(warmup + timed):
        q, a = steps[i]
        t0 = time.perf_counter()
        t.knn(q, args.k, MAX_DIST, nthreads=nthr, want_points=False)
        t.add_points(a, True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            nq_tot += len(q)
    t.wait_rebuild()
    vis = t.mean_visits(steps[warmup][0][:4000], args.k, MAX_DIST)
    t.close()
    return {"qps": nq_tot / sum(times), "ms_per_step": 1e3 * sum(times) / len(times), "threads": nthr,
            "p50_ms": 1e3 * float(np.median(times)), "mean_visits": vis, "steps": len(times)}


def oracle_scanloop_check(pmap, steps, k):
    """Parity checker for the scan loop (outside every timed region): the deterministic C restatement of the reference
    (oracle/ikd_oracle.c, pinned to the reference by tests/test_oracle.py) runs the same steps; returns validnum after
    every step, Add_Points' return values, and the kNN squared distances of the LAST step."""
    import ref_ctypes as R
    if not R.oracle_available():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    o = R.OracleTree(*PARAMS)
    o.build(pmap)
    valid, added, d_last, c_last = [], [], None, None
    for i, (q, a) in enumerate(steps):
        if i == len(steps) - 1:
            _, d_last, c_last = o.knn(q, k, MAX_DIST, nthreads=0, want_points=False)
        added.append(int(o.add_points(a, True)))
        valid.append(int(o.validnum()))
    o.close()
    return valid, added, d_last, c_last


def scanloop_ours(args, rank, world_size, local_rank, W_, K_, n_map=None, with_cpu=True, with_plane=True):
    """configs[1] on this rank's GPU. Returns the result dict (every rank; the caller prints rank 0's)."""
    import ctypes as C

    import torch
    import ikd_ctypes as I
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    n_map = n_map or 1_000_000
    pmap, steps = make_scanloop_inputs(dev, n_map, W_ + K_)
    k = args.k

    def new_tree():
        t = I.Tree(*PARAMS, device=local_rank)
        t.build(pmap)
        return t

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---------------- value: inputs resident in HBM
    tree = new_tree()
    tstream = torch.cuda.ExternalStream(tree.stream(), device=dev)
    qd, ad = [], []
    for q, a in steps:
        q4 = torch.zeros((len(q), 4), dtype=torch.float32, device=dev)
        q4[:, :3] = torch.from_numpy(q).to(dev)
        a4 = torch.zeros((len(a), 4), dtype=torch.float32, device=dev)
        a4[:, :3] = torch.from_numpy(a).to(dev)
        qd.append(q4)
        ad.append(a4)
    nq_max = max(len(q) for q, _ in steps)
    oi = torch.empty((nq_max, k), dtype=torch.int32, device=dev)
    od = torch.empty((nq_max, k), dtype=torch.float32, device=dev)
    oc = torch.empty(nq_max, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    tree.set_kernel_timing(True)
    for i in range(W_):
        tree.knn_dev(qd[i].data_ptr(), qd[i].shape[0], k, MAX_DIST, oi.data_ptr(), od.data_ptr(), oc.data_ptr())
        tree.add_points_dev(ad[i].data_ptr(), ad[i].shape[0], True)
    tree.synchronize()
    tree.kernel_time()  # reset
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier_sync(world_size)
    launches0 = I.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K_)]
    knn_ev = [torch.cuda.Event(enable_timing=True) for _ in range(K_)]
    nq_tot, na_tot = 0, 0
    for j in range(K_):
        i = W_ + j
        with torch.cuda.stream(tstream):
            flush.zero_()  # L2 flush between timed iterations (untimed)
            ev[j][0].record(tstream)
            tree.knn_dev(qd[i].data_ptr(), qd[i].shape[0], k, MAX_DIST, oi.data_ptr(), od.data_ptr(), oc.data_ptr())
            knn_ev[j].record(tstream)
            tree.add_points_dev(ad[i].data_ptr(), ad[i].shape[0], True)
            ev[j][1].record(tstream)
        nq_tot += qd[i].shape[0]
        na_tot += ad[i].shape[0]
    tree.synchronize()
    barrier_sync(world_size)
    launches = I.launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    knn_ms = [ev[j][0].elapsed_time(knn_ev[j]) for j in range(K_)]
    add_ms = [knn_ev[j].elapsed_time(ev[j][1]) for j in range(K_)]
    kern_ms, kern_n = tree.kernel_time()
    total_ms = dist_max(float(sum(step_ms)), dev, world_size)
    nq_all = dist_sum(nq_tot, dev, world_size)
    value = nq_all / (total_ms * 1e-3)
    stats = tree.stats()
    valid_dev = tree.validnum()
    tree.close()

    # ---------------- e2e: host buffers through the public C ABI, H2D + D2H inside the timed region
    tree2 = new_tree()
    hq = [torch.from_numpy(q).pin_memory() for q, _ in steps]
    ha = [torch.from_numpy(a).pin_memory() for _, a in steps]
    h_idx = torch.empty((nq_max, k), dtype=torch.int32).pin_memory()
    h_d = torch.empty((nq_max, k), dtype=torch.float32).pin_memory()
    h_c = torch.empty(nq_max, dtype=torch.int32).pin_memory()
    added_e2e, valid_e2e = [], []

    def host_step(i):
        n = hq[i].shape[0]
        st = tree2.L.ikd_knn_batch(tree2.h, hq[i].data_ptr(), n, 12, k, MAX_DIST, h_idx.data_ptr(), h_d.data_ptr(),
                                   h_c.data_ptr())
        assert st == 0, tree2.L.ikd_last_error()
        added, first, nins = C.c_int(), C.c_int32(), C.c_int64()
        st = tree2.L.ikd_add_points(tree2.h, ha[i].data_ptr(), ha[i].shape[0], 12, 1, C.byref(added), C.byref(first),
                                    C.byref(nins), None)
        assert st == 0, tree2.L.ikd_last_error()
        return added.value

    for i in range(W_):
        added_e2e.append(host_step(i))
        valid_e2e.append(tree2.validnum())
    tree2.synchronize()
    barrier_sync(world_size)
    e2e_t, h2d, d2h = [], 0, 0
    d_last = c_last = None
    main2 = torch.cuda.ExternalStream(tree2.stream(), device=dev)
    for j in range(K_):
        i = W_ + j
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a = host_step(i)
        # everything the step enqueued on the tree's stream (inline rebuilds included) is waited for; a rebuild handed to
        # the side stream is background work by design, like the reference's rebuild thread, which its arm does not wait
        # for per step either -- it is finished by the device-wide synchronize of the (untimed) L2 flush that follows
        main2.synchronize()
        e2e_t.append(time.perf_counter() - t0)
        added_e2e.append(a)
        valid_e2e.append(tree2.validnum())
        h2d += hq[i].shape[0] * 12 + ha[i].shape[0] * 12
        d2h += hq[i].shape[0] * (8 * k + 4) + 64
        if j == K_ - 1:
            n = hq[i].shape[0]
            d_last, c_last = h_d[:n].numpy().copy(), h_c[:n].numpy().copy()
    e2e_total = dist_max(float(sum(e2e_t)), dev, world_size)
    e2e_value = nq_all / e2e_total
    # extra (not part of `value` / `e2e`): the same queries through ikd_knn_plane_batch (kNN + the caller's plane fit on the
    # device, 21 B/query back) next to plain ikd_knn_batch (8k+4 B/query back), host buffers, queries only
    plane_extra = None
    if with_plane and 3 <= k <= 8:
        h_pl = torch.empty((nq_max, 4), dtype=torch.float32).pin_memory()
        h_rs = torch.empty(nq_max, dtype=torch.float32).pin_memory()
        h_vl = torch.empty(nq_max, dtype=torch.uint8).pin_memory()

        def plane_call(i):
            st = tree2.L.ikd_knn_plane_batch(tree2.h, hq[i].data_ptr(), hq[i].shape[0], 12, k, MAX_DIST, 5.0, 0.1,
                                             h_pl.data_ptr(), h_rs.data_ptr(), h_vl.data_ptr(), None)
            assert st == 0, tree2.L.ikd_last_error()

        def knn_call(i):
            st = tree2.L.ikd_knn_batch(tree2.h, hq[i].data_ptr(), hq[i].shape[0], 12, k, MAX_DIST, h_idx.data_ptr(),
                                       h_d.data_ptr(), h_c.data_ptr())
            assert st == 0, tree2.L.ikd_last_error()

        res = {}
        for name, fn in (("knn_plane", plane_call), ("knn", knn_call)):
            for i in range(min(W_, 3)):
                fn(i)
            ts, nqs, nvalid = [], 0, 0
            for j in range(K_):
                i = W_ + j
                flush.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn(i)
                ts.append(time.perf_counter() - t0)
                nqs += hq[i].shape[0]
                if name == "knn_plane":
                    nvalid += int(h_vl[:hq[i].shape[0]].sum())
            res[name] = {"qps": nqs / sum(ts), "p50_ms": 1e3 * float(np.median(ts))}
            if name == "knn_plane":
                res[name]["valid_fraction"] = nvalid / max(nqs, 1)
        plane_extra = {"host_buffers_queries_only": res, "d2h_bytes_per_query": {"knn_plane": 21, "knn": 8 * k + 4},
                       "max_kth_sqdist": 5.0, "plane_threshold": 0.1}
    tree2.close()

    # ---------------- work counters (untimed replay on a fresh tree with visit counting on): V of OUR kNN traversal on one
    # timed step, and V_box / descent depth of Add_Points for its algorithmic-bytes figure
    tree3 = new_tree()
    tree3.set_visit_counting(True)
    our_visits = None
    for i in range(W_ + K_):
        tree3.knn_dev(qd[i].data_ptr(), qd[i].shape[0], k, MAX_DIST, oi.data_ptr(), od.data_ptr(), oc.data_ptr())
        if i == W_:
            tree3.synchronize()
            our_visits = tree3.stats()["last_knn_visits"] / qd[i].shape[0]
        if i == W_:
            st0 = tree3.stats()
        tree3.add_points_dev(ad[i].data_ptr(), ad[i].shape[0], True)
    tree3.synchronize()
    st1 = tree3.stats()
    valid_replay = tree3.validnum()
    tree3.close()
    add_in = st1["add_points_in"] - st0["add_points_in"]
    add_ins = st1["add_points_inserted"] - st0["add_points_inserted"]
    add_vox = st1["add_vox_visits"] - st0["add_vox_visits"]
    add_desc = st1["add_descend_levels"] - st0["add_descend_levels"]

    if rank != 0:
        return None
    peak, peak_kind = measured_peak()
    out = {
        "metric": "5-NN queries/s (FAST-LIO2 scan loop: per-scan 5-NN max_dist 5 m + Add_Points downsample 0.5 m, 1M-point map)",
        "value": value, "unit": "queries/s", "n_gpus": world_size, "steps": K_, "warmup": W_,
        "ms_per_step": total_ms / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": scanloop_config(len(pmap), k),
        "parallelism": "replicas only (the per-scan loop stays on one GPU)" if world_size > 1 else "single GPU",
        "queries_per_scan_mean": nq_tot / K_, "add_points_per_scan_mean": na_tot / K_,
        "scan_p50_ms": float(np.median(step_ms)), "scan_mean_ms": float(np.mean(step_ms)),
        "step_ms": [round(float(x), 3) for x in step_ms],
        "knn_ms_per_step": float(np.mean(knn_ms)), "add_points_ms_per_step": float(np.mean(add_ms)),
        "knn_p50_ms": float(np.median(knn_ms)), "add_points_p50_ms": float(np.median(add_ms)),
        "knn_only_qps": nq_tot / (sum(knn_ms) * 1e-3),
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d // K_, "d2h_bytes_per_step": d2h // K_,
                "scan_p50_ms": 1e3 * float(np.median(e2e_t)), "scan_mean_ms": 1e3 * float(np.mean(e2e_t)),
                "note": "pinned host buffers through ikd_knn_batch + ikd_add_points, then a wait for the tree's stream; the L2 "
                        "flush and its device-wide sync (which also ends side-stream rebuilds) sit outside the timer, so the "
                        "first kernel of every step starts on a cold L2"},
        "gpu_launches": int(launches), "clocks": clocks,
        "tree_stats": {k_: (int(v) if isinstance(v, int) else float(v)) for k_, v in stats.items()},
    }
    if plane_extra:
        out["knn_plane_fit"] = plane_extra
    # parity: the same steps on the CPU checker (outside the timed regions); a mismatch fails the run
    parity = {"checker": "oracle/ikd_oracle.c (C restatement pinned to the reference)", "steps": W_ + K_}
    try:
        ov, oa, od_last, oc_last = oracle_scanloop_check(pmap, steps, k)
        parity["validnum_equal_after_every_step"] = bool(ov == valid_e2e)
        parity["add_points_return_equal"] = bool(oa == added_e2e)
        parity["knn_distances_bitexact_last_step"] = bool(np.array_equal(od_last.view(np.uint32), d_last.view(np.uint32)) and
                                                          np.array_equal(oc_last, c_last))
        parity["validnum_device_path_equal"] = bool(valid_dev == ov[-1] and valid_replay == ov[-1])
        parity["validnum_final"] = int(ov[-1])
        parity["ok"] = all(v for k_, v in parity.items() if k_.endswith(("equal", "step", "every_step")) and isinstance(v, bool))
    except ParityError:
        raise
    except Exception as e:  # the checker is test infrastructure; say so instead of hiding the measurement
        parity["ok"] = None
        parity["unavailable"] = str(e)
    out["parity"] = parity
    if parity["ok"] is False:
        raise ParityError("scan loop: GPU results differ from the CPU checker: " + json.dumps(parity))
    # cpu baseline + roofline (rank 0, N=1)
    V = None
    if with_cpu and world_size == 1 and not args.no_cpu_baseline:
        try:
            nb = min(W_ + K_, 3 + 10)
            rb = run_reference_scanloop(args, pmap, steps, min(W_, 3), nb - min(W_, 3))
            out["cpu_baseline"] = {"value": rb["qps"], "unit": "queries/s", "cores": rb["threads"], "kind": "reference",
                                   "sample": f"{rb['steps']} full scan steps (OpenMP Nearest_Search over all queries + Add_Points) "
                                             f"on the same 1M-point map, after Build; p50 {rb['p50_ms']:.1f} ms/scan",
                                   "scan_p50_ms": rb["p50_ms"]}
            V = rb["mean_visits"]
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    v_src = "reference traversal (ref_mean_visits on the same map and queries)"
    if V is None:
        V, v_src = our_visits, "this implementation's own visit count (no CPU leg in this run)"
    bytes_per_q = 12 + 8 * k + 64 * V
    achieved = (bytes_per_q * nq_tot) / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
    out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                       "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic("knn_ncu_summary.json"),
                       "peak_kind": peak_kind, "kernel": "knn_coop_kernel<%d, 4> (4 lanes per query)" % k, "launches": int(kern_n),
                       "kernel_ms_mean": kern_ms / max(kern_n, 1), "algorithmic_bytes_per_query": bytes_per_q,
                       "visits_per_query": V, "visits_source": v_src, "visits_per_query_ours": our_visits,
                       "note": "1M-point tree (64 B search records) fits the 126 MB L2, so HBM is not the binding roof here: "
                               "achieved is algorithmic bytes / kernel time; see profiles/ for the L2 throughput from ncu"}
    # Add_Points: 12 B per input point + 64 B per node visited by the voxel box searches (every input point of the
    # reference runs one) + 2 x 64 B per level descended by the points that are inserted (delete walk + insert walk)
    if add_in > 0:
        nsc = K_
        a_bytes = 12.0 * add_in + 64.0 * (add_vox + 2.0 * add_desc)
        a_ms = float(np.sum(add_ms))
        out["roofline_add_points"] = {
            "bound": "hbm", "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": a_bytes / (a_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_kind": peak_kind,
            "kernel": "Add_Points launch chain (voxel decision, box delete, insert, refit, rebuild)",
            "formula": "12*n_in + 64*(sum V_box over input points + 2*sum descent levels over inserted points)  (SURVEY 8d)",
            "points_in_per_scan": add_in / nsc, "points_inserted_per_scan": add_ins / nsc,
            "v_box_mean": add_vox / add_in, "descent_depth_mean": add_desc / max(add_ins, 1),
            "algorithmic_bytes_per_scan": a_bytes / nsc, "ms_per_scan": a_ms / nsc,
            "note": "latency-bound launch chain on an L2-resident working set, not a bandwidth kernel"}
    return out


# ---------------------------------------------------------------------------------------------------
# configs[3]: large-batch kNN, map replicated per GPU, queries sharded
# ---------------------------------------------------------------------------------------------------
def largebatch_config(args):
    n_map = args.map_points or 100_000_000
    nq = args.queries or 100_000_000
    return {"workload": "configs[3] large-batch kNN: map replicated per GPU, queries sharded", "map_points": n_map,
            "queries": nq, "k": args.k, "max_dist": "inf", "extent_m": [-LB_EXT, LB_EXT],
            "map_seed": LB_MAP_SEED, "query_seed": LB_QUERY_SEED, "gpus": args.gpus,
            "parallelism": f"query-sharded x{args.gpus} (contiguous shards), replica broadcast from rank 0",
            "l2": "tree and query set far larger than L2; no flush needed"}


def largebatch_metric(k):
    return f"{k}-NN queries/s (large batch: map replicated per GPU, queries sharded)"


def sample_queries(n, ext):
    """The first n queries of the global query set (CPU generator, so that both arms and the parity check see the same bits)."""
    import torch
    g = torch.Generator().manual_seed(LB_QUERY_SEED)
    return (torch.rand((n, 3), generator=g) * (2 * ext) - ext).numpy().astype(np.float32)


def run_reference_largebatch(n_map, k, ext, sample, steps, warmup):
    """The unmodified reference on the host cores for configs[3]: Build of the same map, then `steps` timed OpenMP passes of
    Nearest_Search over the query sample (after `warmup` untimed ones). Returns rate, distances (for the parity check), V."""
    import bench_workloads as W
    import ref_ctypes as R
    with stdout_to_stderr():
        pm = W.uniform_cloud(n_map, -ext, ext, LB_MAP_SEED)
        t = R.RefTree(*PARAMS, serial=False)
        t0 = time.perf_counter()
        t.build(pm)
        tb = time.perf_counter() - t0
        del pm
        nthr = host_threads()
        t.knn(sample[:20000], k, float("inf"), nthreads=nthr, want_points=False)  # page in
        times = []
        d = c = None
        for i in range(max(warmup, 0) + max(steps, 1)):
            t0 = time.perf_counter()
            _, d, c = t.knn(sample, k, float("inf"), nthreads=nthr, want_points=False)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        t1 = time.perf_counter()
        t.knn(sample[:50000], k, float("inf"), nthreads=1, want_points=False)
        single = 50000 / (time.perf_counter() - t1)
        vis = t.mean_visits(sample[:20000], k, float("inf"))
        depth = t.max_depth()
        t.close()
    return {"qps": len(sample) * len(times) / sum(times), "threads": nthr, "build_s": tb, "sample_q": len(sample),
            "ms_per_step": 1e3 * sum(times) / len(times), "p50_ms": 1e3 * float(np.median(times)), "steps": len(times),
            "d": d, "c": c, "mean_visits": vis, "single_thread_qps": single, "depth": depth}


def largebatch_ours(args, rank, world_size, local_rank):
    import torch
    import ikd_ctypes as I
    import bench_workloads as W
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg = largebatch_config(args)
    n_map, nq_all, k = cfg["map_points"], cfg["queries"], args.k
    ext = LB_EXT
    W_, K_ = args.warmup, args.steps
    tree = I.Tree(*PARAMS, device=local_rank)
    t_build = t_bcast = 0.0
    bcast_bytes = 0
    if rank == 0:
        pm = W.uniform_cloud(n_map, -ext, ext, LB_MAP_SEED)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tree.build(pm)
        tree.synchronize()
        t_build = time.perf_counter() - t0
        del pm
    if world_size > 1:
        import torch.distributed as dist
        from replica_sync import broadcast_tree
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        slots, npoints = broadcast_tree(tree, src=0, rank=rank, device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        t_bcast = time.perf_counter() - t0
        bcast_bytes = slots * 144 + npoints * 16
    # this rank's contiguous shard of the global query set; the set's first LB_SAMPLE queries are the CPU sample
    per = (nq_all + world_size - 1) // world_size
    lo = min(nq_all, rank * per)
    n = max(0, min(per, nq_all - lo))
    n_sample = min(LB_SAMPLE, n) if rank == 0 else 0
    g = torch.Generator(device=dev).manual_seed(LB_QUERY_SEED + 1 + rank)
    q4 = torch.zeros((n, 4), dtype=torch.float32, device=dev)
    q4[:, :3] = torch.rand((n, 3), generator=g, device=dev) * (2 * ext) - ext
    sample = None
    if n_sample:
        sample = sample_queries(n_sample, ext)
        q4[:n_sample, :3] = torch.from_numpy(sample).to(dev)
    oi = torch.empty((n, k), dtype=torch.int32, device=dev)
    od = torch.empty((n, k), dtype=torch.float32, device=dev)
    oc = torch.empty(n, dtype=torch.int32, device=dev)
    tstream = torch.cuda.ExternalStream(tree.stream(), device=dev)
    torch.cuda.synchronize()

    def dev_pass(m=n, kk=k, o=(oi, od, oc)):
        tree.knn_dev(q4.data_ptr(), m, kk, float("inf"), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())

    for _ in range(W_):
        dev_pass()
    tree.synchronize()
    tree.set_kernel_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier_sync(world_size)
    l0 = I.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K_ + 1)]
    with torch.cuda.stream(tstream):
        evs[0].record(tstream)
        for j in range(K_):
            dev_pass()
            evs[j + 1].record(tstream)
    tree.synchronize()
    barrier_sync(world_size)
    launches = dist_sum(I.launch_count() - l0, dev, world_size)
    clocks = sampler.stop()
    step_ms = [evs[j].elapsed_time(evs[j + 1]) for j in range(K_)]
    ms = dist_max(float(sum(step_ms)), dev, world_size)
    kern_ms, kern_n = tree.kernel_time()
    value = nq_all * K_ / (ms * 1e-3)
    tree.set_kernel_timing(False)
    d_dev_sample = od[:n_sample].cpu().numpy() if n_sample else None
    c_dev_sample = oc[:n_sample].cpu().numpy() if n_sample else None
    tree.set_visit_counting(True)
    m = min(n, 2_000_000)
    dev_pass(m)
    tree.synchronize()
    V_ours = tree.stats()["last_knn_visits"] / max(m, 1)
    tree.set_visit_counting(False)
    depth = tree.stats()["max_depth"]

    # ---- extra: k = 32 on the same replica (N=1): 25M-query passes (the 32-wide result rows of 100M queries do not fit)
    extras = want_extras(args, world_size, full=(n_map >= 50_000_000))
    k32 = None
    if "k32" in extras and k != 32:
        m32 = min(n, 25_000_000)
        oi32 = torch.empty((m32, 32), dtype=torch.int32, device=dev)
        od32 = torch.empty((m32, 32), dtype=torch.float32, device=dev)
        oc32 = torch.empty(m32, dtype=torch.int32, device=dev)
        dev_pass(m32, 32, (oi32, od32, oc32))
        tree.synchronize()
        tree.set_kernel_timing(True)
        tree.kernel_time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(tstream):
            e0.record(tstream)
            for _ in range(2):
                dev_pass(m32, 32, (oi32, od32, oc32))
            e1.record(tstream)
        tree.synchronize()
        ms32 = e0.elapsed_time(e1)
        kms32, kn32 = tree.kernel_time()
        tree.set_kernel_timing(False)
        tree.set_visit_counting(True)
        dev_pass(min(m32, 1_000_000), 32, (oi32, od32, oc32))
        tree.synchronize()
        V32 = tree.stats()["last_knn_visits"] / min(m32, 1_000_000)
        tree.set_visit_counting(False)
        bpq32 = 12 + 8 * 32 + 64 * V32
        peak, peak_kind = measured_peak()
        k32 = {"k": 32, "queries_per_pass": m32, "value": 2 * m32 / (ms32 * 1e-3), "unit": "queries/s",
               "roofline": {"bound": "hbm", "achieved": bpq32 * m32 * kn32 / (kms32 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": bpq32 * m32 * kn32 / (kms32 * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "knn_heap_kernel",
                            "visits_per_query": V32, "visits_source": "this implementation's own visit count",
                            "algorithmic_bytes_per_query": bpq32}}
        del oi32, od32, oc32

    # ---- e2e: host shard through the C ABI (page-locked buffers), H2D + D2H inside the timed region
    del oi, od, oc
    hq = torch.empty((n, 3), dtype=torch.float32, pin_memory=True)
    hq.copy_(q4[:, :3])
    h_idx = torch.empty((n, k), dtype=torch.int32, pin_memory=True)
    h_d = torch.empty((n, k), dtype=torch.float32, pin_memory=True)
    h_c = torch.empty(n, dtype=torch.int32, pin_memory=True)
    del q4
    torch.cuda.empty_cache()

    def host_pass():
        st = tree.L.ikd_knn_batch(tree.h, hq.data_ptr(), n, 12, k, float("inf"), h_idx.data_ptr(), h_d.data_ptr(), h_c.data_ptr())
        assert st == 0, tree.L.ikd_last_error()
        tree.synchronize()

    host_pass()  # warm-up: lane buffers are allocated on first use
    barrier_sync(world_size)
    e2e_runs = []
    for _ in range(max(1, min(K_, 3))):
        barrier_sync(world_size)
        t0 = time.perf_counter()
        host_pass()
        e2e_runs.append(dist_max(time.perf_counter() - t0, dev, world_size))
    e2e_s = float(np.median(e2e_runs))
    d_e2e_sample = h_d[:n_sample].numpy().copy() if n_sample else None
    c_e2e_sample = h_c[:n_sample].numpy().copy() if n_sample else None
    # extra: the same pass through ikd_knn_plane_batch (kNN + plane fit on the device; 21 B/query come back instead of 8k+4)
    plane_extra = None
    if 3 <= k <= 8 and world_size == 1 and "k32" in extras:
        del h_idx, h_d, h_c
        h_pl = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
        h_rs = torch.empty(n, dtype=torch.float32, pin_memory=True)
        h_vl = torch.empty(n, dtype=torch.uint8, pin_memory=True)

        def plane_pass():
            st = tree.L.ikd_knn_plane_batch(tree.h, hq.data_ptr(), n, 12, k, float("inf"), 5.0, 0.1, h_pl.data_ptr(),
                                            h_rs.data_ptr(), h_vl.data_ptr(), None)
            assert st == 0, tree.L.ikd_last_error()
            tree.synchronize()

        plane_pass()
        runs = []
        for _ in range(2):
            t0 = time.perf_counter()
            plane_pass()
            runs.append(time.perf_counter() - t0)
        plane_extra = {"e2e_qps": n / float(np.median(runs)), "d2h_bytes_per_query": 21, "h2d_bytes_per_query": 12,
                       "valid_fraction": float(h_vl.float().mean())}
        del h_pl, h_rs, h_vl
    tree.close()
    del hq
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    peak, peak_kind = measured_peak()
    out = {
        "metric": largebatch_metric(k), "value": value, "unit": "queries/s", "n_gpus": world_size, "steps": K_, "warmup": W_,
        "ms_per_step": ms / K_, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "step_ms": [round(float(x), 3) for x in step_ms],
        "build_s": t_build, "tree_depth": int(depth),
        "replica_broadcast_s": t_bcast, "replica_broadcast_bytes": int(bcast_bytes),
        "replica_broadcast_GBps": (bcast_bytes / t_bcast / 1e9) if t_bcast > 0 else None,
        "shard_queries_rank0": int(n),
        "e2e": {"value": nq_all / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": n * 12, "d2h_bytes_per_step": n * (8 * k + 4),
                "ms_per_step": 1e3 * e2e_s, "note": "per rank: its shard through ikd_knn_batch with page-locked host buffers "
                                                    "(byte counts are rank 0's); time = max over ranks"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    # ---- cpu baseline (N=1): the reference itself on the same map, 1M-query sample; bit-exact parity check on that sample
    V_ref = None
    cpu = None
    parity = {"checker": "oracle/_ref (the unmodified reference)", "queries": int(n_sample)}
    if world_size == 1 and not args.no_cpu_baseline:
        import ref_ctypes as R
        need_gb = 40.0 * n_map / 100e6 + 4
        avail = host_mem_available_gb()
        if not R.available():
            cpu = {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference", "sample": "unavailable: oracle/_ref not built"}
        elif avail is not None and avail < need_gb:
            cpu = {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference",
                   "sample": f"unavailable: {avail:.0f} GB of host memory free, the reference's {n_map}-point tree needs ~{need_gb:.0f} GB"}
        else:
            rb = run_reference_largebatch(n_map, k, ext, sample, steps=2, warmup=1)
            cpu = {"value": rb["qps"], "unit": "queries/s", "cores": rb["threads"], "kind": "reference",
                   "sample": f"Build of the same {n_map}-point map ({rb['build_s']:.1f} s, depth {rb['depth']}), then OpenMP Nearest_Search "
                             f"({rb['threads']} threads, schedule(static)) over the first {rb['sample_q']} queries of the same query set, "
                             f"{rb['steps']} timed passes after 1 warm-up pass; one thread alone: {rb['single_thread_qps']:.0f} q/s",
                   "build_s": rb["build_s"], "single_thread_qps": rb["single_thread_qps"]}
            V_ref = rb["mean_visits"]
            parity["knn_distances_bitexact_device_path"] = bool(np.array_equal(rb["d"].view(np.uint32), d_dev_sample.view(np.uint32)) and
                                                                np.array_equal(rb["c"], c_dev_sample))
            parity["knn_distances_bitexact_host_path"] = bool(np.array_equal(rb["d"].view(np.uint32), d_e2e_sample.view(np.uint32)) and
                                                              np.array_equal(rb["c"], c_e2e_sample))
            parity["ok"] = parity["knn_distances_bitexact_device_path"] and parity["knn_distances_bitexact_host_path"]
            if not parity["ok"]:
                raise ParityError("large batch: GPU kNN distances differ from the reference's on the sample: " + json.dumps(parity))
    if cpu is not None:
        out["cpu_baseline"] = cpu
    if "ok" not in parity:
        # no CPU leg in this run: the two GPU paths (device-resident and host-buffer) must at least agree with each other
        if n_sample:
            parity["checker"] = "device-resident path vs host-buffer path (no CPU leg in this run; tests/ hold both to the oracle)"
            parity["ok"] = bool(np.array_equal(d_dev_sample.view(np.uint32), d_e2e_sample.view(np.uint32)))
            if not parity["ok"]:
                raise ParityError("large batch: device-resident and host-buffer results differ")
    out["parity"] = parity
    # V of the reference traversal: measured live at N=1; at N>1 from the committed N=1 measurement of this exact workload
    v_src = "reference traversal, measured in this run (ref_mean_visits on the same map, 20k queries of the sample)"
    V = V_ref
    if V is None:
        pj = profile_json("c4_reference_visits.json") or {}
        key = f"map{n_map}_k{k}"
        if key in pj:
            V, v_src = float(pj[key]), "reference traversal, from profiles/c4_reference_visits.json (measured by the N=1 run of this workload)"
        else:
            V, v_src = V_ours, "this implementation's own visit count (no reference measurement for this map size)"
    bytes_per_q = 12 + 8 * k + 64 * V
    achieved = bytes_per_q * n * kern_n / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
    full = (n_map == 100_000_000 and nq_all == 100_000_000 and k == 5 and world_size == 1)
    out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                       # the committed capture is of this configuration at N=1 (100M-point map, 100M queries, one launch, k=5)
                       "traffic": ncu_traffic("knn_large_ncu_summary.json") if full else None,
                       "kernel": "knn_reg_persist_kernel<%d>" % k if k <= 8 else "knn_heap_kernel", "launches": int(kern_n),
                       "kernel_ms_mean": kern_ms / max(kern_n, 1), "queries_per_launch": int(n), "per_gpu": True,
                       "peak_kind": peak_kind, "visits_per_query": V, "visits_source": v_src, "visits_per_query_ours": V_ours,
                       "algorithmic_bytes_per_query": bytes_per_q,
                       "nominal_8TBps_frac": achieved / 8000.0 if achieved else None}
    if k32:
        out["k32"] = k32
    if plane_extra:
        out["knn_plane_fit"] = plane_extra
    return out


def want_extras(args, world_size, full):
    if args.extras == "none" or world_size > 1 or args.impl != "ours":
        return set()
    if args.extras == "default":
        return {"scan_loop", "k32", "c3", "c5"} if full else set()
    return set(x for x in args.extras.split(",") if x)


# ---------------------------------------------------------------------------------------------------
# extras (N=1): configs[2] box / radius search, configs[4] streaming
# ---------------------------------------------------------------------------------------------------
def c3_extra(args, local_rank, n=10_000_000, nq=100_000):
    """configs[2]: batched Box_Search / Radius_Search on a 10M-point map, 100k queries with half-extent / radius 0.5-5 m.
    Device time = count + scan + fill (ids stay in HBM); e2e adds the fetch of all ids into page-locked memory."""
    import torch
    import ikd_ctypes as I
    import bench_workloads as W
    dev = torch.device("cuda", local_rank)
    P = W.uniform_cloud(n, -50, 50, 3)
    c, rad, boxes = W.range_queries(nq, -50, 50, 0.5, 5.0, 33)
    t = I.Tree(device=local_rank)
    t.build(P)
    off = np.empty(nq + 1, dtype=np.int64)
    radf = np.ascontiguousarray(rad, dtype=np.float32)
    peak, peak_kind = measured_peak()
    res = {"config": {"workload": "configs[2] batched Box_Search / Radius_Search", "map_points": n, "queries": nq,
                      "half_extent_or_radius_m": [0.5, 5.0]}}

    def run(kind):
        if kind == "box":
            st = t.L.ikd_box_search_batch(t.h, boxes.ctypes.data, nq, off.ctypes.data)
        else:
            st = t.L.ikd_radius_search_batch(t.h, c.ctypes.data, radf.ctypes.data, nq, off.ctypes.data)
        assert st == 0, t.L.ikd_last_error()
        return int(off[-1])

    for kind in ("box", "radius"):
        total = run(kind)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(kind)
            t.synchronize()  # the call returns with the offsets; the kernel that lays the ids out contiguously is still running
            ts.append(time.perf_counter() - t0)
        dt = float(np.median(ts))
        pin = torch.empty(max(total, 1), dtype=torch.int32, pin_memory=True)
        te = []
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(kind)
            assert t.L.ikd_search_fetch(t.h, pin.data_ptr(), total) == 0
            te.append(time.perf_counter() - t0)
        # algorithmic bytes: 24 + 64 V_partial + 32 M per query; V_partial is not counted here, so the figure
        # below is the 32 M part alone (a lower bound of the algorithmic bytes)
        a_bytes = 32.0 * total + 24.0 * nq
        res[kind] = {"results": total, "device_s": dt, "queries_per_s": nq / dt, "points_per_s": total / dt,
                     "e2e_pinned_fetch_s": float(np.median(te)), "e2e_queries_per_s": nq / float(np.median(te)),
                     "roofline": {"bound": "hbm", "achieved": a_bytes / dt / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": a_bytes / dt / 1e9 / peak, "traffic": ncu_traffic(f"range_{kind}_ncu_summary.json"),
                                  "formula": "24*nq + 32*M (reported points; the 64*V_partial term is left out: lower bound)",
                                  "note": "wall time of the call + stream sync: H2D of the queries, traversal, scan, D2H of the offsets, gather of the ids into contiguous ranges"}}
        del pin
    # the reference, one thread, on a query sample (its range searches are not thread-safe)
    if not args.no_cpu_baseline:
        import ref_ctypes as R
        if R.available():
            with stdout_to_stderr():
                r = R.RefTree(serial=False)
                t0 = time.perf_counter()
                r.build(P)
                rb = time.perf_counter() - t0
                m = 300
                t0 = time.perf_counter()
                nb = sum(len(r.box_search(boxes[i], cap=1 << 16)) for i in range(m))
                rbox = (time.perf_counter() - t0) / m
                t0 = time.perf_counter()
                nr = sum(len(r.radius_search(c[i], rad[i], cap=1 << 16)) for i in range(m))
                rrad = (time.perf_counter() - t0) / m
                r.close()
            # parity on the sample: counts per query equal (box: exact; radius: fresh build, same shape -> exact)
            res["cpu_baseline"] = {"kind": "reference", "cores": 1, "build_s": rb, "box_queries_per_s": 1 / rbox,
                                   "radius_queries_per_s": 1 / rrad, "sample": f"{m} queries of each kind, one thread"}
            run("box")
            ok_box = int(off[m]) == nb
            run("radius")
            ok_rad = int(off[m]) == nr
            res["parity"] = {"checker": "oracle/_ref", "box_result_count_equal_on_sample": ok_box,
                             "radius_result_count_equal_on_sample": ok_rad, "ok": bool(ok_box and ok_rad)}
            if not (ok_box and ok_rad):
                raise ParityError("c3: result counts differ from the reference on the sample: " + json.dumps(res["parity"]))
    t.close()
    del P
    return res


def c5_extra(args, local_rank, scans=1000):
    """configs[4]: streaming moving map. Per scan: 5-NN batch for the voxel-filtered scan, Delete_Point_Boxes with the
    boxes that leave a 200 m local-map cube, Add_Points(downsample 0.5 m). Host-buffer API; scan generation is outside the
    latencies. Rebuild counts and device times come from the library's rebuild timing."""
    import torch
    import ikd_ctypes as I
    import bench_workloads as W
    dev = torch.device("cuda", local_rank)
    world = W.LidarWorld(seed=5, device=dev)
    half = 100.0
    t = I.Tree(0.5, 0.6, 0.5, device=local_rank)
    t.set_rebuild_timing(True)
    lat_knn, lat_del, lat_add, nq_tot, ndel_tot = [], [], [], 0, 0
    scans_pts, poses = [], []
    for i in range(scans):
        o, yaw = world.pose(i, 2.0)
        scans_pts.append(world.voxel_filter(world.scan(o, yaw), 0.25).cpu().numpy().astype(np.float32))
        poses.append(o)
    torch.cuda.synchronize()
    t.build(scans_pts[0])
    prev = poses[0]
    valid_trace = []
    t_start = time.perf_counter()
    for i in range(1, scans):
        pts, o = scans_pts[i], poses[i]
        t0 = time.perf_counter()
        t.knn(pts, 5, 5.0)
        t1 = time.perf_counter()
        boxes = W.local_map_boxes(o, half, prev)
        if len(boxes):
            ndel_tot += t.delete_boxes(boxes)
        t2 = time.perf_counter()
        t.add_points(pts, True)
        t3 = time.perf_counter()
        prev = o
        lat_knn.append(t1 - t0); lat_del.append(t2 - t1); lat_add.append(t3 - t2); nq_tot += len(pts)
        valid_trace.append(t.validnum())
    wall = time.perf_counter() - t_start
    t.synchronize()
    st = t.stats()
    upd = np.asarray(lat_del) + np.asarray(lat_add)
    worst = int(np.argmax(upd))
    res = {"config": {"workload": "configs[4] streaming incremental map", "scans": scans, "local_map_cube_m": 2 * half,
                      "downsample_m": 0.5, "scan_filter_leaf_m": 0.25, "params": [0.5, 0.6, 0.5]},
           "scans_timed": len(lat_knn), "mean_queries_per_scan": nq_tot / max(len(lat_knn), 1),
           "knn_ms": {"p50": 1e3 * float(np.median(lat_knn)), "p99": 1e3 * float(np.percentile(lat_knn, 99)), "max": 1e3 * float(np.max(lat_knn))},
           "update_ms": {"p50": 1e3 * float(np.median(upd)), "p99": 1e3 * float(np.percentile(upd, 99)), "max": 1e3 * float(np.max(upd)),
                         "max_at_scan": worst + 1},
           "add_points_ms": {"p50": 1e3 * float(np.median(lat_add)), "p99": 1e3 * float(np.percentile(lat_add, 99)), "max": 1e3 * float(np.max(lat_add))},
           "delete_boxes_ms": {"p50": 1e3 * float(np.median(lat_del)), "max": 1e3 * float(np.max(lat_del))},
           "points_deleted_by_boxes": int(ndel_tot), "validnum_final": int(t.validnum()), "size_final": int(t.size()),
           "rebuilds": {"subtrees": int(st["rebuilds_partial"]), "whole_tree": int(st["rebuilds_full"]),
                        "subtrees_on_side_stream": int(st["rebuilds_async"]), "rebuilt_points": int(st["rebuilt_points"]),
                        "inline_batches": int(st["rebuild_inline_n"]), "inline_ms_mean": st["rebuild_inline_ms"] / max(st["rebuild_inline_n"], 1),
                        "side_stream_batches": int(st["rebuild_async_n"]), "side_stream_ms_mean": st["rebuild_async_ms"] / max(st["rebuild_async_n"], 1),
                        "whole_tree_ms_mean": st["rebuild_full_ms"] / max(st["rebuild_full_n"], 1),
                        "longest_ms": st["rebuild_max_ms"]},
           "node_slots_used": int(st["node_slots_used"]), "max_depth": int(st["max_depth"]), "loop_wall_s": wall,
           "note": "host-buffer API (ikd_knn_batch / ikd_delete_boxes / ikd_add_points), pageable numpy inputs; scan generation excluded"}
    t.close()
    # the reference: one caller thread + its own rebuild thread (how it runs), same scans; kNN with all host cores
    if not args.no_cpu_baseline:
        import ref_ctypes as R
        if R.available():
            with stdout_to_stderr():
                r = R.RefTree(0.5, 0.6, 0.5, serial=False)
                r.build(scans_pts[0])
                prev = poses[0]
                nthr = host_threads()
                m = min(scans, 300)
                rk, ru = [], []
                rvalid = []
                for i in range(1, m):
                    pts, o = scans_pts[i], poses[i]
                    t0 = time.perf_counter()
                    r.knn(pts, 5, 5.0, nthreads=nthr, want_points=False)
                    t1 = time.perf_counter()
                    boxes = W.local_map_boxes(o, half, prev)
                    if len(boxes):
                        r.delete_boxes(boxes)
                    r.add_points(pts, True)
                    t2 = time.perf_counter()
                    prev = o
                    rk.append(t1 - t0); ru.append(t2 - t1)
                    rvalid.append(int(r.validnum()))
                r.wait_rebuild()
                r.close()
            res["cpu_baseline"] = {"kind": "reference", "cores": nthr, "sample": f"the first {m} scans",
                                   "knn_ms": {"p50": 1e3 * float(np.median(rk)), "p99": 1e3 * float(np.percentile(rk, 99))},
                                   "update_ms": {"p50": 1e3 * float(np.median(ru)), "p99": 1e3 * float(np.percentile(ru, 99)),
                                                 "max": 1e3 * float(np.max(ru))}}
            # validnum trace: the reference reports -1 / stale values while its root is being rebuilt (ikd_Tree.cpp:129-146)
            # and its batch calls race its rebuild thread, so only settled values are compared, with the count
            neq = sum(1 for a, b in zip(valid_trace[:m - 1], rvalid) if b >= 0 and a != b)
            res["parity"] = {"checker": "oracle/_ref (batch calls, background thread running: not bit-reproducible, see DESIGN 5)",
                             "validnum_trace_mismatches": int(neq), "scans_compared": m - 1}
    return res


# ---------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world_size, local_rank):
    """--impl reference: the unmodified reference CPU implementation on the same config."""
    if rank != 0:
        return None
    import ref_ctypes as R
    if not R.available():
        return {"impl": "reference", "unavailable": "oracle/_ref/libikd_ref.so not built (run make -C oracle ref where /root/reference exists)"}
    if args.workload == "largebatch":
        cfg = largebatch_config(args)
        n_map = cfg["map_points"]
        sample = sample_queries(min(LB_SAMPLE, cfg["queries"]), LB_EXT)
        rb = run_reference_largebatch(n_map, args.k, LB_EXT, sample, steps=args.steps, warmup=args.warmup)
        return {
            "impl": "reference", "metric": largebatch_metric(args.k),
            "value": rb["qps"], "unit": "queries/s", "n_gpus": world_size, "steps": rb["steps"], "warmup": args.warmup,
            "ms_per_step": rb["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rb["qps"], "unit": "queries/s", "cores": rb["threads"], "kind": "reference",
                             "sample": f"Build of the {n_map}-point map ({rb['build_s']:.1f} s) + OpenMP Nearest_Search ({rb['threads']} threads) "
                                       f"per step over the first {rb['sample_q']} queries of the query set (bounded sample of the "
                                       f"{cfg['queries']}-query workload)"},
            "build_s": rb["build_s"], "visits_per_query": rb["mean_visits"], "single_thread_qps": rb["single_thread_qps"],
            "e2e": {"value": rb["qps"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
    import torch
    dev = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
    n_map = args.map_points or 1_000_000
    pmap, steps = make_scanloop_inputs(dev, n_map, args.warmup + args.steps)
    rb = run_reference_scanloop(args, pmap, steps, args.warmup, args.steps)
    return {
        "impl": "reference",
        "metric": "5-NN queries/s (FAST-LIO2 scan loop: per-scan 5-NN max_dist 5 m + Add_Points downsample 0.5 m, 1M-point map)",
        "value": rb["qps"], "unit": "queries/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": rb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": scanloop_config(len(pmap), args.k),
        "scan_p50_ms": rb["p50_ms"],
        "cpu_baseline": {"value": rb["qps"], "unit": "queries/s", "cores": rb["threads"], "kind": "reference",
                         "sample": f"{rb['steps']} full scan steps, OpenMP Nearest_Search ({rb['threads']} threads) + Add_Points"},
        "e2e": {"value": rb["qps"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        out = reference_arm(args, rank, world_size, local_rank)
        print(json.dumps(out), flush=True)
        return 0
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: bench.py measures the B200 path and has no CPU fallback"}), flush=True)
        return 2
    if world_size > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        with stdout_to_stderr():  # NCCL announces its version on stdout when the communicator is created
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
    rc = 0
    try:
        if args.workload == "scanloop":
            out = scanloop_ours(args, rank, world_size, local_rank, args.warmup, args.steps, n_map=args.map_points)
        else:
            out = largebatch_ours(args, rank, world_size, local_rank)
            if rank == 0:
                full = (args.map_points or 100_000_000) >= 50_000_000
                extras = want_extras(args, world_size, full)
                if "scan_loop" in extras:
                    out["scan_loop"] = scanloop_ours(args, 0, 1, local_rank, 3, 20, with_plane=False)
                if "c3" in extras:
                    out["c3_range_search"] = c3_extra(args, local_rank)
                if "c5" in extras:
                    out["c5_streaming"] = c5_extra(args, local_rank, scans=args.c5_scans)
        if rank == 0:
            print(json.dumps(out), flush=True)
    except ParityError as e:
        if rank == 0:
            print(json.dumps({"error": "parity check failed", "detail": str(e)}), file=sys.stderr, flush=True)
        rc = 3
    finally:
        if world_size > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
