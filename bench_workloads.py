"""Seeded synthetic inputs for the five BASELINE.json configs.

Data generation only (torch/numpy, CPU or GPU); nothing here is on the measured path. Both bench arms
(ours and --impl reference) and the tests call the same functions with the same seeds, so they see the
same inputs.
"""
import math

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# c1 / c3 / c4: uniform cubes
# ------------------------------------------------------------------------------------------------
def uniform_cloud(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def demo_boxes(nb, side, lo, hi, seed):
    """Axis-aligned cubes of the given side at uniform centres (ikd_Tree_demo.cpp:140-155)."""
    rng = np.random.default_rng(seed)
    c = (rng.random((nb, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
    h = np.float32(side / 2)
    return np.concatenate([c - h, c + h], axis=1).astype(np.float32)


def range_queries(nq, lo, hi, rmin, rmax, seed):
    """Centres uniform in the cube, half-extent / radius uniform in [rmin, rmax] (config c3)."""
    rng = np.random.default_rng(seed)
    c = (rng.random((nq, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
    r = (rng.random(nq, dtype=np.float32) * np.float32(rmax - rmin) + np.float32(rmin)).astype(np.float32)
    boxes = np.concatenate([c - r[:, None], c + r[:, None]], axis=1).astype(np.float32)
    return c, r, boxes


# ------------------------------------------------------------------------------------------------
# c2 / c5: synthetic spinning LiDAR in a block city
# ------------------------------------------------------------------------------------------------
class LidarWorld:
    """Ground plane + axis-aligned box buildings; a 64-beam spinning LiDAR ray-cast against them.

    beams: 64 between -24.8 and +2 degrees; azimuth step 0.17 degrees (~2118 columns, ~135k rays);
    range noise sigma 2 cm; max range 100 m.
    """

    def __init__(self, seed=2, extent=400.0, n_buildings=360, device=None, beams=64, az_step_deg=0.17,
                 max_range=100.0, noise_sigma=0.02, sensor_height=1.8):
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        g = torch.Generator().manual_seed(seed)
        half = extent / 2
        ctr = (torch.rand((n_buildings, 2), generator=g) * 2 - 1) * half
        size = torch.rand((n_buildings, 2), generator=g) * 22.0 + 6.0
        height = torch.rand((n_buildings, 1), generator=g) * 26.0 + 4.0
        bmin = torch.cat([ctr - size / 2, torch.zeros(n_buildings, 1)], dim=1)
        bmax = torch.cat([ctr + size / 2, height], dim=1)
        self.bmin = bmin.to(self.device)
        self.bmax = bmax.to(self.device)
        self.extent = extent
        self.max_range = max_range
        self.noise_sigma = noise_sigma
        self.sensor_height = sensor_height
        el = torch.deg2rad(torch.linspace(-24.8, 2.0, beams))
        az = torch.deg2rad(torch.arange(0.0, 360.0, az_step_deg))
        el_g, az_g = torch.meshgrid(el, az, indexing="ij")
        d = torch.stack([torch.cos(el_g) * torch.cos(az_g), torch.cos(el_g) * torch.sin(az_g), torch.sin(el_g)], dim=-1)
        self.dirs = d.reshape(-1, 3).to(self.device).float()
        self.gen = torch.Generator(device=self.device).manual_seed(seed + 1000)

    def pose(self, i, step=1.0):
        """Straight lanes joined by U-turns (a lawnmower sweep of the block city along x, then along y,
        repeated): sensor origin and yaw of scan i, `step` metres apart along the path."""
        lanes, span, gap = 10, 360.0, 40.0
        lane_len = span + gap  # one lane plus the hop to the next
        sweep = lanes * lane_len
        s = (i * step) % (2 * sweep)
        along_y = s >= sweep
        s = s % sweep
        k = int(s // lane_len)
        t = s - k * lane_len
        fwd = (k % 2) == 0
        c = -180.0 + gap * k
        if t <= span:
            a = -180.0 + t if fwd else 180.0 - t
            yaw = 0.0 if fwd else math.pi
            x, y = a, c
        else:  # hop to the next lane
            a = 180.0 if fwd else -180.0
            x, y = a, c + (t - span)
            yaw = math.pi / 2
        if along_y:
            x, y, yaw = y, x, math.pi / 2 - yaw
        return np.array([x, y, self.sensor_height]), yaw

    def scan(self, origin, yaw=0.0):
        """World-frame hit points (float32 [m,3]) of one revolution from `origin`."""
        dev = self.device
        o = torch.tensor(origin, dtype=torch.float32, device=dev)
        cy, sy = math.cos(yaw), math.sin(yaw)
        rot = torch.tensor([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32, device=dev)
        d = self.dirs @ rot.T
        tmax = torch.full((d.shape[0],), float("inf"), device=dev)
        # ground z = 0
        tg = torch.where(d[:, 2] < -1e-6, -o[2] / d[:, 2], tmax)
        t_hit = torch.minimum(tmax, tg)
        # buildings (slab test), chunked over rays
        inv = 1.0 / torch.where(d.abs() < 1e-9, torch.full_like(d, 1e-9), d)
        out = []
        CH = 32768
        for s in range(0, d.shape[0], CH):
            iv = inv[s:s + CH]
            t0 = (self.bmin[None, :, :] - o[None, None, :]) * iv[:, None, :]
            t1 = (self.bmax[None, :, :] - o[None, None, :]) * iv[:, None, :]
            tn = torch.minimum(t0, t1).amax(dim=2)
            tf = torch.maximum(t0, t1).amin(dim=2)
            hit = (tf >= tn) & (tf > 0)
            tb = torch.where(hit, torch.where(tn > 0, tn, tf), torch.full_like(tn, float("inf"))).amin(dim=1)
            out.append(tb)
        t_hit = torch.minimum(t_hit, torch.cat(out))
        ok = t_hit < self.max_range
        t_ok = t_hit[ok]
        t_ok = t_ok + torch.randn(t_ok.shape, generator=self.gen, device=dev) * self.noise_sigma
        pts = o[None, :] + d[ok] * t_ok[:, None]
        return pts.float()

    @staticmethod
    def voxel_filter(pts, leaf):
        """One point per voxel (the one nearest to the voxel centre), like FAST-LIO2's scan downsampling."""
        if pts.shape[0] == 0:
            return pts
        key = torch.floor(pts / leaf).to(torch.int64)
        key = key - key.amin(dim=0, keepdim=True)
        dims = key.amax(dim=0) + 1
        lin = (key[:, 0] * dims[1] + key[:, 1]) * dims[2] + key[:, 2]
        centre = (torch.floor(pts / leaf) + 0.5) * leaf
        dist = ((pts - centre) ** 2).sum(dim=1)
        order = torch.argsort(dist, stable=True)
        lin_o = lin[order]
        order2 = torch.argsort(lin_o, stable=True)
        lin_s = lin_o[order2]
        first = torch.ones_like(lin_s, dtype=torch.bool)
        first[1:] = lin_s[1:] != lin_s[:-1]
        return pts[order[order2[first]]]

    def build_map(self, target_points, leaf=0.5, scan_stride=2.0, max_scans=4000):
        """Accumulate voxel-filtered scans along the trajectory until the map holds `target_points` points.
        Returns (map float32 [n,3] on CPU, next scan index)."""
        acc = torch.empty((0, 3), dtype=torch.float32, device=self.device)
        i = 0
        pending = []
        while i < max_scans:
            o, yaw = self.pose(i, scan_stride)
            pending.append(self.voxel_filter(self.scan(o, yaw), leaf))
            i += 1
            if len(pending) == 8:
                acc = self.voxel_filter(torch.cat([acc] + pending), leaf)
                pending = []
                if acc.shape[0] >= target_points:
                    break
        if pending:
            acc = self.voxel_filter(torch.cat([acc] + pending), leaf)
        acc = acc[:target_points] if acc.shape[0] > target_points else acc
        return acc.cpu().numpy().astype(np.float32), i

    def scan_step(self, i, leaf=0.5, scan_stride=2.0, seed=0):
        """Inputs of one measured scan: (queries [m,3] = voxel-filtered scan under a perturbed pose,
        adds [m,3] = the same points at the true pose)."""
        o, yaw = self.pose(i, scan_stride)
        pts = self.voxel_filter(self.scan(o, yaw), leaf)
        g = torch.Generator().manual_seed(seed * 100003 + i)
        dt = ((torch.rand(3, generator=g) * 2 - 1) * 0.05).to(self.device)      # <= 5 cm
        dyaw = float((torch.rand(1, generator=g) * 2 - 1) * math.radians(0.5))   # <= 0.5 deg
        c, s = math.cos(dyaw), math.sin(dyaw)
        rot = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32, device=self.device)
        oo = torch.tensor(o, dtype=torch.float32, device=self.device)
        q = (pts - oo) @ rot.T + oo + dt
        return q.float().cpu().numpy().astype(np.float32), pts.cpu().numpy().astype(np.float32)


def local_map_boxes(center, half, prev_center):
    """Boxes that fall out of a cube local map of half-size `half` when it moves from prev_center to center
    (config c5; FAST-LIO2's lasermap_fov_segment idea). Returns [nb,6] float32, nb in 0..3."""
    boxes = []
    big = 1.0e4
    lo_prev = np.asarray(prev_center, dtype=np.float64) - half
    hi_prev = np.asarray(prev_center, dtype=np.float64) + half
    lo_new = np.asarray(center, dtype=np.float64) - half
    hi_new = np.asarray(center, dtype=np.float64) + half
    for a in range(3):
        if lo_new[a] > lo_prev[a]:
            b = [-big, -big, -big, big, big, big]
            b[a] = lo_prev[a] - 1.0
            b[3 + a] = lo_new[a]
            boxes.append(b)
        elif hi_new[a] < hi_prev[a]:
            b = [-big, -big, -big, big, big, big]
            b[a] = hi_new[a]
            b[3 + a] = hi_prev[a] + 1.0
            boxes.append(b)
    return np.asarray(boxes, dtype=np.float32).reshape(-1, 6)
