"""oracle/lkref.py — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/_ref/liblkref.so: the reference's OWN eskf.cc / voxel_map.cc / KILO.cc, compiled unmodified
from /root/reference against stand-in third-party headers (oracle/ref/Makefile, oracle/ref/shim/). It exists to pin the
restatement in oracle/lko_core.cpp: tests/test_oracle_vs_reference.py drives both with the same buffers, and
tests/golden/make_ref_golden.py freezes its outputs as fixtures for the boxes that have no /root/reference.

Same method names as lko.Oracle where the reference has the operation; the reference has no gain / iteration / map
options, so there is no set_options beyond the stream's mode and constants.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "leg-kilo_b200", "python"))
from legkilo_b200 import abi  # noqa: E402  (POD struct mirrors only)

_SO = os.path.join(_HERE, "_ref", "liblkref.so")
_LIB = None


def available() -> bool:
    """True when the library is built, or can be built here (needs /root/reference)."""
    return os.path.exists(_SO) or os.path.isdir("/root/reference/legkilo/src")


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref")])
        L = _LIB = C.CDLL(_SO)
        vp = C.c_void_p
        L.lkref_create.restype = vp
        L.lkref_create.argtypes = [vp, vp, vp, vp, C.c_int, C.c_double]
        L.lkref_destroy.argtypes = [vp]
        L.lkref_set_filter.argtypes = [vp] * 5
        L.lkref_get_filter.argtypes = [vp] * 5
        L.lkref_set_runtime.argtypes = [vp, C.c_double, C.c_int]
        L.lkref_acc_norm.restype = C.c_double
        L.lkref_acc_norm.argtypes = [vp]
        L.lkref_init_process_cov.argtypes = [vp]
        L.lkref_predict.argtypes = [vp, C.c_double, C.c_int, C.c_int]
        L.lkref_build_voxel_map.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, vp]
        L.lkref_predict_update_point.restype = C.c_int
        L.lkref_predict_update_point.argtypes = [vp, C.c_double, vp, C.c_uint32, vp, vp]
        L.lkref_obs_imu.argtypes = [vp, vp, C.c_uint32]
        L.lkref_obs_kinimu.argtypes = [vp, vp, C.c_uint32]
        L.lkref_process.restype = C.c_int
        L.lkref_process.argtypes = [vp, C.c_double, C.c_double, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp]
        L.lkref_map_slide.restype = C.c_int
        L.lkref_map_slide.argtypes = [vp, vp]
        L.lkref_calc_body_cov.argtypes = [vp, C.c_float, C.c_float, vp]
        L.lkref_init_plane.argtypes = [C.c_uint32, vp, vp, C.c_float] + [vp] * 7
        L.lkref_boxminus.argtypes = [vp] * 3
        L.lkref_map_num_roots.restype = C.c_uint64
        L.lkref_map_num_roots.argtypes = [vp]
        L.lkref_map_export.restype = C.c_int
        L.lkref_map_export.argtypes = [vp, vp, C.c_size_t, vp]
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Reference:
    """One legkilo::KILO (KILO.h:20-66) built from the same config dict as lko.Oracle."""

    def __init__(self, cfg: dict, imu_mode_only=True, gravity=9.81, acc_norm=1.0, initialised=True):
        self.cfg = cfg
        ec, mc = abi.eskf_cfg(cfg), abi.map_cfg(cfg)
        R, t = abi.extrinsics(cfg)
        self.h = lib().lkref_create(C.byref(ec), C.byref(mc), _p(R), _p(t), int(imu_mode_only), gravity)
        lib().lkref_set_runtime(self.h, acc_norm, int(initialised))

    def __del__(self):
        if getattr(self, "h", None):
            lib().lkref_destroy(self.h)
            self.h = None

    def set_filter(self, x=None, P=None, Q=None, clk=None):
        x = None if x is None else np.ascontiguousarray(x, abi.STATE_DTYPE)
        P = None if P is None else np.ascontiguousarray(P, np.float64)
        Q = None if Q is None else np.ascontiguousarray(Q, np.float64)
        clk = None if clk is None else np.ascontiguousarray(clk, abi.CLOCK_DTYPE)
        lib().lkref_set_filter(self.h, _p(x), _p(P), _p(Q), _p(clk))

    def get_filter(self):
        x = np.zeros(1, abi.STATE_DTYPE)
        P = np.zeros(900)
        Q = np.zeros(900)
        clk = np.zeros(1, abi.CLOCK_DTYPE)
        lib().lkref_get_filter(self.h, _p(x), _p(P), _p(Q), _p(clk))
        return x, P, Q, clk

    def init_process_cov(self):
        lib().lkref_init_process_cov(self.h)

    def acc_norm(self) -> float:
        return float(lib().lkref_acc_norm(self.h))

    def predict(self, dt, prop_state, prop_cov):
        lib().lkref_predict(self.h, dt, int(prop_state), int(prop_cov))

    def build_voxel_map(self, xyz_world, xyz_body, R=None, rot_cov=None, pos_cov=None):
        xyz_world = np.ascontiguousarray(xyz_world, np.float32)
        xyz_body = np.ascontiguousarray(xyz_body, np.float32)
        R = np.eye(3) if R is None else np.ascontiguousarray(R, np.float64)
        rot_cov = 1e-6 * np.eye(3) if rot_cov is None else np.ascontiguousarray(rot_cov, np.float64)
        pos_cov = 1e-6 * np.eye(3) if pos_cov is None else np.ascontiguousarray(pos_cov, np.float64)
        lib().lkref_build_voxel_map(self.h, _p(xyz_world), _p(xyz_body), len(xyz_world), _p(R), _p(rot_cov), _p(pos_cov))

    def predict_update_point(self, t, pts):
        pts = np.ascontiguousarray(pts, np.float32)
        n = len(pts)
        world = np.zeros((n, 4), np.float32)
        neff = np.zeros(1, np.uint32)
        upd = lib().lkref_predict_update_point(self.h, t, _p(pts), n, _p(world), _p(neff))
        return dict(updated=bool(upd), world=world, n_eff=int(neff[0]))

    def obs_imu(self, imu):
        imu = np.ascontiguousarray(imu, abi.IMU_DTYPE)
        lib().lkref_obs_imu(self.h, _p(imu), len(imu))

    def obs_kinimu(self, kin):
        kin = np.ascontiguousarray(kin, abi.KINIMU_DTYPE)
        lib().lkref_obs_kinimu(self.h, _p(kin), len(kin))

    def process(self, begin_time, end_time, pts, imu=None, kin=None):
        """KILO::process on an already downsampled cloud; returns the cloud in the order process() sorted it into."""
        pts = np.ascontiguousarray(pts, np.float32)
        n = len(pts)
        body = np.zeros((n, 4), np.float32)
        world = np.zeros((n, 4), np.float32)
        neff = np.zeros(1, np.uint32)
        imu = None if imu is None else np.ascontiguousarray(imu, abi.IMU_DTYPE)
        kin = None if kin is None else np.ascontiguousarray(kin, abi.KINIMU_DTYPE)
        ok = lib().lkref_process(self.h, begin_time, end_time, _p(pts), n, _p(imu), 0 if imu is None else len(imu), _p(kin),
                                 0 if kin is None else len(kin), _p(body), _p(world), _p(neff))
        return dict(ok=bool(ok), body=body, world=world, n_eff=int(neff[0]))

    def map_slide(self, position_last) -> bool:
        p = np.ascontiguousarray(position_last, np.float64)
        return bool(lib().lkref_map_slide(self.h, _p(p)))

    def map_export(self) -> np.ndarray:
        sz = C.c_size_t(0)
        lib().lkref_map_export(self.h, None, 0, C.byref(sz))
        buf = np.zeros(sz.value, np.uint8)
        assert lib().lkref_map_export(self.h, _p(buf), buf.size, C.byref(sz)) == 0
        return buf

    def num_roots(self) -> int:
        return int(lib().lkref_map_num_roots(self.h))


def calc_body_cov(pb, range_inc, degree_inc):
    pb = np.array(pb, np.float64)
    cov = np.zeros(9)
    lib().lkref_calc_body_cov(_p(pb), range_inc, degree_inc, _p(cov))
    return cov.reshape(3, 3), pb


def init_plane(pw, var, planer_threshold=0.01):
    pw = np.ascontiguousarray(pw, np.float64)
    var = np.ascontiguousarray(var, np.float64)
    center = np.zeros(3); normal = np.zeros(3); pv = np.zeros(36)
    d = np.zeros(1, np.float32); radius = np.zeros(1, np.float32); isp = np.zeros(1, np.int32); eig = np.zeros(3, np.float32)
    lib().lkref_init_plane(len(pw), _p(pw), _p(var), planer_threshold, _p(center), _p(normal), _p(pv), _p(d), _p(radius),
                           _p(isp), _p(eig))
    return dict(center=center, normal=normal, plane_var=pv.reshape(6, 6), d=float(d[0]), radius=float(radius[0]),
                is_plane=bool(isp[0]), eig=eig)


def boxminus(a, b):
    a = np.ascontiguousarray(a, abi.STATE_DTYPE); b = np.ascontiguousarray(b, abi.STATE_DTYPE)
    out = np.zeros(30)
    lib().lkref_boxminus(_p(a), _p(b), _p(out))
    return out
