import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200"))
import bench_workloads as W, ikd_ctypes as I
world = W.LidarWorld(seed=5, device="cuda"); half = 100.0
t = I.Tree(0.5, 0.6, 0.5); prev = None
for i in range(300):
    o, yaw = world.pose(i, 2.0)
    pts = world.voxel_filter(world.scan(o, yaw), 0.25).cpu().numpy().astype(np.float32)
    if i == 0: t.build(pts); prev = o; continue
    torch.cuda.synchronize(); s0 = t.stats(); t0 = time.perf_counter()
    t.knn(pts, 5, 5.0); t1 = time.perf_counter()
    boxes = W.local_map_boxes(o, half, prev)
    nd = t.delete_boxes(boxes) if len(boxes) else 0
    t2 = time.perf_counter()
    a = t.add_points(pts, True); t3 = time.perf_counter(); prev = o
    s1 = t.stats()
    if (t3 - t0) > 0.004:
        print(i, f"knn {1e3*(t1-t0):.2f} del {1e3*(t2-t1):.2f} add {1e3*(t3-t2):.2f} ms n={len(pts)} boxes={len(boxes)} deleted={nd} added={a[0]} valid={t.validnum()} size={t.size()}",
              {k: s1[k] - s0[k] for k in ("rebuilds_partial", "rebuilds_full", "rebuilds_async", "rebuilt_points", "node_slots_used", "node_slots_cap")}, flush=True)
t.close()
