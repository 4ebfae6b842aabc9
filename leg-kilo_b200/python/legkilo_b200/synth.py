"""Seeded synthetic scenes for the BASELINE.json configs.

Everything here produces plain numpy arrays that are handed, unchanged, to both the CUDA library
and (in tests / the cpu_baseline leg of bench.py) the CPU oracle — so bit-identical inputs need no
cross-language RNG. Base seed 0x4C4B494C4F ("LKILO").

Point layout everywhere: float32 [n, 4] = (x, y, z, curvature) in the LiDAR body frame, the four
fields of the reference's 48-byte pcl::PointXYZINormal that the hot path reads
(legkilo/src/core/slam/KILO.cc:123-127; curvature = per-point time offset in seconds,
legkilo/src/preprocess/lidar_processing.cc:48).
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 0x4C4B494C4F


def rng(stream: int, seed: int = BASE_SEED) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, stream])))


def exp_so3(v) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def world_to_body(pw: np.ndarray, R: np.ndarray, p: np.ndarray, ext_R: np.ndarray, ext_t: np.ndarray) -> np.ndarray:
    """Inverse of pw = R (ext_R pb + ext_t) + p  (KILO.cc:127-129)."""
    pi = (pw - p) @ R  # R^T (pw - p), row-vector form
    return (pi - ext_t) @ ext_R


# ---------------------------------------------------------------------------------------------
# Config 1: planar scene
# ---------------------------------------------------------------------------------------------

def planar_map_points(half_extent: float = 20.0, z: float = -0.75, voxel: float = 0.5, pts_per_voxel: int = 8,
                      sigma: float = 0.01, ext_R=None, ext_t=None, stream: int = 1):
    """Ground plane z=const sampled `pts_per_voxel` per voxel column (first-frame cloud).
    Returns (xyz_world float32 [n,3], xyz_body float32 [n,3]) for BuildVoxelMap with R=I, p=0."""
    ext_R = np.eye(3) if ext_R is None else np.asarray(ext_R, float)
    ext_t = np.zeros(3) if ext_t is None else np.asarray(ext_t, float)
    g = rng(stream)
    nv = int(round(2 * half_extent / voxel))
    ix, iy = np.meshgrid(np.arange(nv), np.arange(nv), indexing="ij")
    base = np.stack([ix.ravel(), iy.ravel()], 1).astype(np.float64) * voxel - half_extent
    base = np.repeat(base, pts_per_voxel, axis=0)
    xy = base + g.uniform(0.02, voxel - 0.02, size=base.shape)
    zz = z + sigma * g.standard_normal(len(xy))
    pw = np.concatenate([xy, zz[:, None]], 1)
    pb = world_to_body(pw, np.eye(3), np.zeros(3), ext_R, ext_t)
    pb32 = pb.astype(np.float32)
    # the reference stores the world cloud as float (KILO.cc:101-103): world = f32(R(ext pb)+p)
    pw32 = ((pb32.astype(np.float64) @ ext_R.T) + ext_t).astype(np.float32)
    return pw32, pb32


def planar_scan(n: int = 2048, radius: float = 15.0, z: float = -0.75, sigma: float = 0.01,
                rotvec=(2e-3, -1e-3, 3e-3), trans=(0.02, -0.01, 0.03), ext_R=None, ext_t=None, stream: int = 2,
                blind: float = 0.0):
    """n points uniform in a disc on the plane as seen from the TRUE pose (Exp(rotvec), trans)."""
    ext_R = np.eye(3) if ext_R is None else np.asarray(ext_R, float)
    ext_t = np.zeros(3) if ext_t is None else np.asarray(ext_t, float)
    g = rng(stream)
    R = exp_so3(rotvec)
    p = np.asarray(trans, float)
    out = np.zeros((0, 3))
    while len(out) < n:
        m = 2 * (n - len(out)) + 16
        r = radius * np.sqrt(g.uniform(0, 1, m))
        a = g.uniform(0, 2 * np.pi, m)
        pw = np.stack([r * np.cos(a), r * np.sin(a), z + sigma * g.standard_normal(m)], 1)
        pb = world_to_body(pw, R, p, ext_R, ext_t)
        keep = np.linalg.norm(pb, axis=1) >= blind
        out = np.concatenate([out, pb[keep]], 0)
    pts = np.zeros((n, 4), np.float32)
    pts[:, :3] = out[:n].astype(np.float32)
    return pts


# ---------------------------------------------------------------------------------------------
# Configs 2-5: box room (ground + 4 walls) ray-cast by a spinning LiDAR
# ---------------------------------------------------------------------------------------------

class BoxScene:
    """Ground plane z=zg over [-E,E]^2 and four walls x=+-W, y=+-W from zg up to z_top.
    Plane offsets sit mid-voxel (…25) so that noisy samples stay inside one root voxel."""

    def __init__(self, ground_half_extent: float = 250.0, wall: float = 15.25, zg: float = -0.75, z_top: float = 6.25,
                 voxel: float = 0.5, rooms=None):
        self.E = float(ground_half_extent)
        self.W = float(wall)
        self.zg = float(zg)
        self.z_top = float(z_top)
        self.voxel = float(voxel)
        # room centres (multiples of the voxel size so every room looks the same on the voxel grid)
        self.rooms = [(0.0, 0.0)] if rooms is None else [(float(a), float(b)) for a, b in rooms]

    @staticmethod
    def room_grid(n_side: int, spacing: float = 62.0):
        """n_side x n_side room centres, `spacing` metres apart, centred on the origin."""
        c = (np.arange(n_side) - (n_side - 1) / 2.0) * spacing
        c = np.round(c / 0.5) * 0.5
        return [(float(a), float(b)) for a in c for b in c]

    # -- first-frame style dense cloud for the map --------------------------------------------
    def map_points(self, pts_per_voxel: int = 8, sigma: float = 0.01, ext_R=None, ext_t=None, stream: int = 11,
                   ground_half_extent: float | None = None):
        ext_R = np.eye(3) if ext_R is None else np.asarray(ext_R, float)
        ext_t = np.zeros(3) if ext_t is None else np.asarray(ext_t, float)
        g = rng(stream)
        v = self.voxel
        E = self.E if ground_half_extent is None else float(ground_half_extent)
        chunks = []
        # ground
        nv = int(round(2 * E / v))
        ix, iy = np.meshgrid(np.arange(nv, dtype=np.int32), np.arange(nv, dtype=np.int32), indexing="ij")
        base = np.stack([ix.ravel(), iy.ravel()], 1).astype(np.float64) * v - E
        base = np.repeat(base, pts_per_voxel, axis=0)
        xy = base + g.uniform(0.02, v - 0.02, size=base.shape)
        zz = self.zg + sigma * g.standard_normal(len(xy))
        chunks.append(np.concatenate([xy, zz[:, None]], 1))
        # walls: cells over (along, height)
        nl = int(round(2 * self.W / v)) + 1
        nh = int(np.ceil((self.z_top - self.zg) / v))
        il, ih = np.meshgrid(np.arange(nl), np.arange(nh), indexing="ij")
        cell = np.stack([il.ravel(), ih.ravel()], 1).astype(np.float64)
        cell = np.repeat(cell, pts_per_voxel, axis=0)
        for (rcx, rcy) in self.rooms:
            rc = (rcx, rcy)
            for axis, sign in ((0, 1), (0, -1), (1, 1), (1, -1)):
                u = cell + g.uniform(0.04, 0.96, size=cell.shape)
                along = -self.W - 0.25 + u[:, 0] * v
                hz = np.floor(self.zg / v) * v + u[:, 1] * v
                off = sign * self.W + sigma * g.standard_normal(len(u))
                ok = (np.abs(along) < self.W) & (hz > self.zg + 0.05) & (hz < self.z_top)
                p = np.zeros((ok.sum(), 3))
                p[:, axis] = rc[axis] + off[ok]
                p[:, 1 - axis] = rc[1 - axis] + along[ok]
                p[:, 2] = hz[ok]
                chunks.append(p)
        pw = np.concatenate(chunks, 0)
        # every point is "seen" from the centre of its nearest room (identity attitude), the way a
        # first frame taken there would see it: body = ext^-1 (world - room centre)
        rc = np.asarray(self.rooms, float)
        near = np.zeros(len(pw), np.int64)
        best = np.full(len(pw), np.inf)
        for i, c in enumerate(rc):
            d2 = (pw[:, 0] - c[0]) ** 2 + (pw[:, 1] - c[1]) ** 2
            m = d2 < best
            near[m] = i
            best[m] = d2[m]
        origin = np.concatenate([rc[near], np.zeros((len(pw), 1))], 1)
        pb = world_to_body(pw - origin, np.eye(3), np.zeros(3), ext_R, ext_t)
        pb32 = pb.astype(np.float32)
        pw32 = (((pb32.astype(np.float64) @ ext_R.T) + ext_t) + origin).astype(np.float32)
        return pw32, pb32

    # -- one LiDAR revolution -------------------------------------------------------------------
    def scan(self, n_rings: int, n_az: int, fov_deg: tuple[float, float], rotvec, trans, ext_R=None, ext_t=None,
             sigma: float = 0.01, blind: float = 1.5, stream: int = 21, scan_period: float = 0.1,
             time_quantum: float = 0.002, streaming: bool = False, room: int = 0):
        """Ray-cast n_rings x n_az rays from the true pose. Returns float32 [n,4]; curvature is
        the 2 ms-quantised time offset (lidar_processing.cc:48) when streaming, else 0."""
        ext_R = np.eye(3) if ext_R is None else np.asarray(ext_R, float)
        ext_t = np.zeros(3) if ext_t is None else np.asarray(ext_t, float)
        g = rng(stream)
        R = exp_so3(rotvec)
        sensor_xy = self.rooms[room]
        p = np.asarray(trans, float) + np.array([sensor_xy[0], sensor_xy[1], 0.0])
        el = np.deg2rad(np.linspace(fov_deg[0], fov_deg[1], n_rings))
        az = (np.arange(n_az) + 0.5) * (2 * np.pi / n_az)
        A, EL = np.meshgrid(az, el, indexing="ij")  # azimuth-major = time order
        d_l = np.stack([np.cos(EL) * np.cos(A), np.cos(EL) * np.sin(A), np.sin(EL)], -1).reshape(-1, 3)
        t_off = np.repeat(np.arange(n_az) / n_az * scan_period, n_rings)
        M = R @ ext_R
        o = R @ ext_t + p
        d = d_l @ M.T
        best = np.full(len(d), np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            tg = (self.zg - o[2]) / d[:, 2]
            hit = o[None, :] + tg[:, None] * d
            ok = (tg > 0) & (np.abs(hit[:, 0] - sensor_xy[0]) <= self.W) & (np.abs(hit[:, 1] - sensor_xy[1]) <= self.W)
            best = np.where(ok, np.minimum(best, tg), best)
            for axis, sign in ((0, 1), (0, -1), (1, 1), (1, -1)):
                c = sensor_xy[axis] + sign * self.W
                tw = (c - o[axis]) / d[:, axis]
                hit = o[None, :] + tw[:, None] * d
                ok = (tw > 0) & (np.abs(hit[:, 1 - axis] - sensor_xy[1 - axis]) <= self.W) & (hit[:, 2] >= self.zg) & (
                    hit[:, 2] <= self.z_top)
                best = np.where(ok, np.minimum(best, tw), best)
        valid = np.isfinite(best)
        rng_m = best + sigma * g.standard_normal(len(best))
        valid &= rng_m >= blind
        pb = d_l[valid] * rng_m[valid, None]
        pts = np.zeros((int(valid.sum()), 4), np.float32)
        pts[:, :3] = pb.astype(np.float32)
        if streaming:
            # curvature = round(t / quantum) * quantum as float (lidar_processing.cc:48)
            pts[:, 3] = (np.round(t_off[valid] / time_quantum) * time_quantum).astype(np.float32)
        return pts


VLP16 = dict(n_rings=16, n_az=1800, fov_deg=(-15.0, 15.0))
OS64 = dict(n_rings=64, n_az=2048, fov_deg=(-16.6, 16.6))


def random_poses(batch: int, rot_sigma: float, trans_sigma: float, stream: int = 31):
    g = rng(stream)
    return rot_sigma * g.standard_normal((batch, 3)), trans_sigma * g.standard_normal((batch, 3))


def bucketize(pts: np.ndarray, begin_time: float = 0.0):
    """Canonical a1 ordering: stable sort by curvature, then maximal equal runs
    (KILO.cc:370-378). Returns (sorted pts, bucket_offsets uint32 [nb+1], bucket_times f64 [nb])."""
    order = np.argsort(pts[:, 3], kind="stable")
    s = np.ascontiguousarray(pts[order])
    if len(s) == 0:
        return s, np.zeros(1, np.uint32), np.zeros(0, np.float64)
    brk = np.flatnonzero(s[1:, 3] != s[:-1, 3]) + 1
    offs = np.concatenate([[0], brk, [len(s)]]).astype(np.uint32)
    times = begin_time + s[offs[:-1], 3].astype(np.float64)
    return s, offs, times


# ---------------------------------------------------------------------------------------------
# Inertial / kinematic sample streams
# ---------------------------------------------------------------------------------------------

def imu_stream(t0: float, t1: float, rate_hz: float = 400.0, stream: int = 41):
    """lk_imu_meas samples on (t0, t1]: gravity-dominated accelerometer, small gyro, white noise."""
    from . import abi
    g = rng(stream)
    n = int(np.floor((t1 - t0) * rate_hz))
    m = np.zeros(n, abi.IMU_DTYPE)
    m["stamp"] = t0 + (np.arange(n) + 1) / rate_hz
    m["acc"] = np.array([0.15, -0.1, 9.79]) + 0.05 * g.standard_normal((n, 3))
    m["gyr"] = np.array([0.01, -0.02, 0.12]) + 0.005 * g.standard_normal((n, 3))
    return m


def kinimu_stream(t0: float, t1: float, rate_hz: float = 400.0, stream: int = 43):
    """lk_kinimu_meas samples (sensor_types.hpp:19-26): 2-4 feet in contact, trotting pattern."""
    from . import abi
    g = rng(stream)
    im = imu_stream(t0, t1, rate_hz, stream)
    n = len(im)
    m = np.zeros(n, abi.KINIMU_DTYPE)
    m["stamp"] = im["stamp"]; m["acc"] = im["acc"]; m["gyr"] = im["gyr"]
    hip = np.array([[0.19, -0.13, -0.3], [0.19, 0.13, -0.3], [-0.19, -0.13, -0.3], [-0.19, 0.13, -0.3]])
    m["foot_pos"] = hip[None] + 0.02 * g.standard_normal((n, 4, 3))
    m["foot_vel"] = 0.05 * g.standard_normal((n, 4, 3))
    phase = (np.arange(n) // 20) % 3
    pat = np.array([[1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 1]], np.int32)
    m["contact"] = pat[phase]
    return m
