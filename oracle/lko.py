"""oracle/lko.py — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/liblko.so, the CPU restatement of the reference hot path. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this; the
product (leg-kilo_b200/) never does. The reference ships no golden vectors; the restatement is pinned against the
reference's OWN sources compiled here over stand-in third-party headers (oracle/lkref.py, oracle/ref/,
tests/test_oracle_vs_reference.py) — see oracle/README.md for what that does and does not cover.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "leg-kilo_b200", "python"))
from legkilo_b200 import abi  # noqa: E402  (POD struct mirrors only — no product code paths)

_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liblko.so")
    srcs = [os.path.join(_HERE, f) for f in ("lko_core.cpp", "lko_capi.cpp", "lko_core.hpp", "lko_linalg.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "legkilo_b200.h"))
    # Rebuild only when asked to or when the library is missing: a gpurun snapshot copy resets
    # mtimes, so an mtime comparison would recompile on every fresh GPU box.
    if force or os.environ.get("LKO_REBUILD") == "1" or not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liblko.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.lko_create.restype = C.c_void_p
        L.lko_create.argtypes = [C.c_void_p] * 4
        L.lko_destroy.argtypes = [C.c_void_p]
        L.lko_set_filter.argtypes = [C.c_void_p] * 5
        L.lko_get_filter.argtypes = [C.c_void_p] * 5
        L.lko_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.lko_predict.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.lko_build_voxel_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        L.lko_predict_update_point.restype = C.c_int
        L.lko_predict_update_point.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_uint32] + [C.c_void_p] * 7
        L.lko_process_scan.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                       C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lko_update_by_points.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.lko_obs_imu.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.lko_obs_kinimu.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.lko_calc_body_cov.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.lko_init_plane.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_float] + [C.c_void_p] * 7
        L.lko_boxminus.argtypes = [C.c_void_p] * 3
        L.lko_boxplus.argtypes = [C.c_void_p] * 2
        L.lko_exp3.argtypes = [C.c_double] * 3 + [C.c_void_p]
        L.lko_log.argtypes = [C.c_void_p] * 2
        L.lko_map_export.restype = C.c_int
        L.lko_map_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.lko_map_import.restype = C.c_int
        L.lko_map_import.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.lko_map_num_roots.restype = C.c_uint64
        L.lko_map_num_roots.argtypes = [C.c_void_p]
        L.lko_decode_pointcloud2.restype = C.c_uint32
        L.lko_decode_pointcloud2.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_float, C.c_int, C.c_double] + [C.c_void_p] * 4
        L.lko_preprocess_scan.restype = C.c_int
        L.lko_preprocess_scan.argtypes = [C.c_void_p, C.c_uint32, C.c_float] + [C.c_void_p] * 5
        L.lko_batch_run.restype = C.c_double
        L.lko_batch_run.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_void_p] * 3
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


GAIN_LITERAL, GAIN_INFORMATION = 0, 1


class Oracle:
    """One reference-style KILO core: ESKF + VoxelMapManager + extrinsics."""

    def __init__(self, cfg: dict):
        self.cfg = cfg
        self._ec = abi.eskf_cfg(cfg)
        self._mc = abi.map_cfg(cfg)
        R, t = abi.extrinsics(cfg)
        self._R, self._t = R, t
        self.h = lib().lko_create(C.byref(self._ec), C.byref(self._mc), _p(R), _p(t))
        self.set_options()

    def __del__(self):
        if getattr(self, "h", None):
            lib().lko_destroy(self.h)
            self.h = None

    def set_options(self, gain_mode=GAIN_LITERAL, iters=1, update_map=True, imu_mode_only=True, gravity=9.81,
                    acc_norm=1.0):
        lib().lko_set_options(self.h, gain_mode, iters, int(update_map), int(imu_mode_only), gravity, acc_norm)

    def set_filter(self, x=None, P=None, Q=None, clk=None):
        x = None if x is None else np.ascontiguousarray(x, abi.STATE_DTYPE)
        P = None if P is None else np.ascontiguousarray(P, np.float64)
        Q = None if Q is None else np.ascontiguousarray(Q, np.float64)
        clk = None if clk is None else np.ascontiguousarray(clk, abi.CLOCK_DTYPE)
        lib().lko_set_filter(self.h, _p(x), _p(P), _p(Q), _p(clk))

    def get_filter(self):
        x = np.zeros(1, abi.STATE_DTYPE)
        P = np.zeros(900)
        Q = np.zeros(900)
        clk = np.zeros(1, abi.CLOCK_DTYPE)
        lib().lko_get_filter(self.h, _p(x), _p(P), _p(Q), _p(clk))
        return x, P, Q, clk

    def predict(self, dt, prop_state, prop_cov):
        lib().lko_predict(self.h, dt, int(prop_state), int(prop_cov))

    def build_voxel_map(self, xyz_world, xyz_body, R=None, rot_cov=None, pos_cov=None):
        xyz_world = np.ascontiguousarray(xyz_world, np.float32)
        xyz_body = np.ascontiguousarray(xyz_body, np.float32)
        R = np.eye(3) if R is None else np.ascontiguousarray(R, np.float64)
        rot_cov = 1e-6 * np.eye(3) if rot_cov is None else np.ascontiguousarray(rot_cov, np.float64)
        pos_cov = 1e-6 * np.eye(3) if pos_cov is None else np.ascontiguousarray(pos_cov, np.float64)
        lib().lko_build_voxel_map(self.h, _p(xyz_world), _p(xyz_body), len(xyz_world), _p(R), _p(rot_cov),
                                  _p(pos_cov))

    def predict_update_point(self, t, pts, debug=False):
        pts = np.ascontiguousarray(pts, np.float32)
        n = len(pts)
        world = np.zeros((n, 4), np.float32)
        neff = np.zeros(1, np.uint32)
        if debug:
            ok = np.zeros(n, np.uint8); h = np.zeros((n, 6)); z = np.zeros(n); R = np.zeros(n)
            key = np.zeros((n, 3), np.int32)
        else:
            ok = h = z = R = key = None
        upd = lib().lko_predict_update_point(self.h, t, _p(pts), n, _p(world), _p(neff), _p(ok), _p(h), _p(z),
                                             _p(R), _p(key))
        out = dict(updated=bool(upd), world=world, n_eff=int(neff[0]))
        if debug:
            out.update(ok=ok, h=h, z=z, R=R, key=key)
        return out

    def process_scan(self, begin_time, pts, imu=None, kin=None):
        pts = np.ascontiguousarray(pts, np.float32)
        n = len(pts)
        world = np.zeros((n, 4), np.float32)
        neff = np.zeros(1, np.uint32)
        ncons = np.zeros(1, np.uint32)
        imu = None if imu is None else np.ascontiguousarray(imu, abi.IMU_DTYPE)
        kin = None if kin is None else np.ascontiguousarray(kin, abi.KINIMU_DTYPE)
        lib().lko_process_scan(self.h, begin_time, _p(pts), n, _p(imu), 0 if imu is None else len(imu), _p(kin),
                               0 if kin is None else len(kin), _p(world), _p(neff), _p(ncons))
        return dict(world=world, n_eff=int(neff[0]), n_consumed=int(ncons[0]))

    def update_by_points(self, h, z, R, gain_mode=GAIN_LITERAL):
        h = np.ascontiguousarray(h, np.float64); z = np.ascontiguousarray(z, np.float64)
        R = np.ascontiguousarray(R, np.float64)
        lib().lko_update_by_points(self.h, len(z), _p(h), _p(z), _p(R), gain_mode)

    def obs_imu(self, imu):
        imu = np.ascontiguousarray(imu, abi.IMU_DTYPE)
        lib().lko_obs_imu(self.h, _p(imu), len(imu))

    def obs_kinimu(self, kin):
        kin = np.ascontiguousarray(kin, abi.KINIMU_DTYPE)
        lib().lko_obs_kinimu(self.h, _p(kin), len(kin))

    def map_export(self) -> np.ndarray:
        sz = C.c_size_t(0)
        lib().lko_map_export(self.h, None, 0, C.byref(sz))
        buf = np.zeros(sz.value, np.uint8)
        rc = lib().lko_map_export(self.h, _p(buf), buf.size, C.byref(sz))
        assert rc == 0
        return buf

    def map_import(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, np.uint8)
        rc = lib().lko_map_import(self.h, _p(blob), blob.size)
        assert rc == 0

    def num_roots(self) -> int:
        return int(lib().lko_map_num_roots(self.h))

    def batch_run(self, x, P, clk, pts, scan_offsets, bucket_times, iters=1, gain_mode=GAIN_INFORMATION, nthreads=1):
        """Independent single-bucket scans against the static map (cpu_baseline). Returns
        (seconds, x_out, P_out, n_eff)."""
        x = np.ascontiguousarray(x, abi.STATE_DTYPE); P = np.ascontiguousarray(P, np.float64)
        clk = np.ascontiguousarray(clk, abi.CLOCK_DTYPE); pts = np.ascontiguousarray(pts, np.float32)
        scan_offsets = np.ascontiguousarray(scan_offsets, np.uint32)
        bucket_times = np.ascontiguousarray(bucket_times, np.float64)
        batch = len(x)
        xo = np.zeros(batch, abi.STATE_DTYPE); Po = np.zeros((batch, 900)); ne = np.zeros(batch, np.uint32)
        sec = lib().lko_batch_run(self.h, batch, _p(x), _p(P), _p(clk), _p(pts), _p(scan_offsets), _p(bucket_times),
                                  iters, gain_mode, nthreads, _p(xo), _p(Po), _p(ne))
        return sec, xo, Po, ne


def calc_body_cov(pb, range_inc, degree_inc):
    pb = np.array(pb, np.float64)
    cov = np.zeros(9)
    lib().lko_calc_body_cov(_p(pb), range_inc, degree_inc, _p(cov))
    return cov.reshape(3, 3), pb


def init_plane(pw, var, planer_threshold=0.01):
    pw = np.ascontiguousarray(pw, np.float64); var = np.ascontiguousarray(var, np.float64)
    center = np.zeros(3); normal = np.zeros(3); pv = np.zeros(36)
    d = np.zeros(1, np.float32); radius = np.zeros(1, np.float32); isp = np.zeros(1, np.int32)
    eig = np.zeros(3, np.float32)
    lib().lko_init_plane(len(pw), _p(pw), _p(var), planer_threshold, _p(center), _p(normal), _p(pv), _p(d),
                         _p(radius), _p(isp), _p(eig))
    return dict(center=center, normal=normal, plane_var=pv.reshape(6, 6), d=float(d[0]), radius=float(radius[0]),
                is_plane=bool(isp[0]), eig=eig)


def boxminus(a, b):
    a = np.ascontiguousarray(a, abi.STATE_DTYPE); b = np.ascontiguousarray(b, abi.STATE_DTYPE)
    out = np.zeros(30)
    lib().lko_boxminus(_p(a), _p(b), _p(out))
    return out


def boxplus(a, delta):
    a = np.array(a, abi.STATE_DTYPE, copy=True)
    delta = np.ascontiguousarray(delta, np.float64)
    lib().lko_boxplus(_p(a), _p(delta))
    return a


def exp3(v):
    out = np.zeros(9)
    lib().lko_exp3(float(v[0]), float(v[1]), float(v[2]), _p(out))
    return out.reshape(3, 3)


def log_so3(R):
    R = np.ascontiguousarray(R, np.float64)
    out = np.zeros(3)
    lib().lko_log(_p(R), _p(out))
    return out


def decode_pointcloud2(data, layout, blind, filter_num, time_scale):
    """lidar_processing.cc:25-108 on a raw PointCloud2 byte buffer; layout = abi.LkPc2Layout."""
    data = np.ascontiguousarray(data, np.uint8)
    n = data.size // layout.point_step
    pts = np.zeros((n, 4), np.float32); inten = np.zeros(n, np.float32)
    ft = np.zeros(1); lt = np.zeros(1)
    m = lib().lko_decode_pointcloud2(_p(data), n, C.byref(layout), blind, filter_num, time_scale, _p(pts), _p(inten), _p(ft), _p(lt))
    return pts[:m].copy(), inten[:m].copy(), float(ft[0]), float(lt[0])


def preprocess_scan(pts, leaf):
    """pcl::VoxelGrid centroid filter + stable curvature sort + equal-curvature runs (KILO.cc:356-378)."""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
    n = len(pts)
    out = np.zeros((n, 4), np.float32); offs = np.zeros(n + 1, np.uint32); curv = np.zeros(max(n, 1), np.float32)
    no = np.zeros(1, np.uint32); nb = np.zeros(1, np.uint32)
    rc = lib().lko_preprocess_scan(_p(pts), n, leaf, _p(out), _p(no), _p(offs), _p(curv), _p(nb))
    assert rc == 0
    return out[:no[0]].copy(), offs[:nb[0] + 1].copy(), curv[:nb[0]].copy()
