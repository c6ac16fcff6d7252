import time, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench_workloads as W
w = W.LidarWorld(seed=2)
torch.cuda.synchronize(); t0 = time.time()
p = w.scan(*w.pose(0)); torch.cuda.synchronize(); print("scan", p.shape, time.time() - t0)
for leaf in (0.5, 0.4, 0.3, 0.25):
    print("leaf", leaf, "filtered", w.voxel_filter(p, leaf).shape[0])
for stride in (4.0,):
    acc = torch.empty((0, 3), device=w.device); t0 = time.time()
    for i in range(0, 3700):
        o, yaw = w.pose(i, stride)
        acc = torch.cat([acc, w.voxel_filter(w.scan(o, yaw), 0.5)])
        if i % 8 == 7: acc = w.voxel_filter(acc, 0.5)
        if i % 200 == 199: torch.cuda.synchronize(); print(stride, i + 1, acc.shape[0], round(time.time() - t0, 2), flush=True)
        if acc.shape[0] > 1_300_000: break
