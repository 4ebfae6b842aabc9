"""SURVEY §8f rank 1 on the device: ESKF::updateByPoints from explicit rows, predictUpdateImu /
predictUpdateKinImu, and the full KILO::process bucket loop with the inertial / kinematic queue interleaved."""
import numpy as np
import pytest

import lko
import scenes
import test_gpu_parity as tp
from legkilo_b200 import Engine, abi, synth

pytestmark = pytest.mark.gpu
TOL = 1e-5
CFG = abi.CONFIGS["leg_fusion"]


def _filter(seed):
    g = synth.rng(seed)
    A = g.standard_normal((30, 30)) * 1e-3
    P = A @ A.T + 1e-6 * np.eye(30)
    x = tp._moving_state()
    x["rot"][0] = lko.exp3(g.normal(size=3) * 0.1).ravel()
    return x, P


@pytest.mark.parametrize("n", [1, 2, 40, 3000])
def test_update_by_points_rows(n):
    g = synth.rng(300 + n)
    x0, P0 = _filter(n)
    h = g.standard_normal((n, 6)); z = g.standard_normal(n) * 1e-2; r = g.uniform(1e-3, 1e-2, n)
    o = lko.Oracle(CFG); o.set_filter(x0, P0.ravel(), None, None)
    o.update_by_points(h, z, r, gain_mode=lko.GAIN_LITERAL if n <= 40 else lko.GAIN_INFORMATION)
    xo, Po, _, _ = o.get_filter()
    xg, Pg = Engine(CFG).update_by_points(x0, P0, h, z, r)
    assert scenes.rel_state_err(xg, xo, x0) < TOL and scenes.rel_cov_err(Pg, Po) < TOL


def test_predict_kernel():
    x0, P0 = _filter(5)
    Q = abi.process_cov_Q(CFG)
    for ps, pc in ((True, True), (True, False), (False, True)):
        o = lko.Oracle(CFG); o.set_filter(x0, P0.ravel(), Q, None); o.predict(0.0123, ps, pc)
        xo, Po, _, _ = o.get_filter()
        xg, Pg = Engine(CFG).predict(x0, P0.reshape(1, 900), Q, [0.0123], ps, pc)
        assert scenes.rel_state_err(xg, xo, x0) < 1e-9 or not ps
        assert np.abs(lko.boxminus(xg, xo)).max() < 1e-12
        assert scenes.rel_cov_err(Pg[0], Po) < 1e-10


@pytest.mark.parametrize("kind", ["imu", "kin"])
def test_inertial_and_kinematic_observations(kind):
    x0, P0 = _filter(9)
    Q = abi.process_cov_Q(CFG)
    clk = np.zeros(1, abi.CLOCK_DTYPE); clk["last_predict_time"] = 3.0; clk["last_update_time"] = 2.995
    meas = synth.imu_stream(3.0, 3.05) if kind == "imu" else synth.kinimu_stream(3.0, 3.05)
    o = lko.Oracle(CFG); o.set_filter(x0, P0.ravel(), Q, clk); o.set_options(imu_mode_only=(kind == "imu"), gravity=9.81, acc_norm=9.79)
    (o.obs_imu if kind == "imu" else o.obs_kinimu)(meas)
    xo, Po, _, co = o.get_filter()
    eng = Engine(CFG)
    xg, Pg, cg = (eng.obs_imu if kind == "imu" else eng.obs_kinimu)(x0, P0, Q, clk, meas, gravity=9.81, acc_norm=9.79)
    assert scenes.rel_state_err(xg, xo, x0) < TOL and scenes.rel_cov_err(Pg, Po) < TOL
    assert cg.tobytes() == co.tobytes()


@pytest.mark.parametrize("kind,update_map", [("imu", False), ("kin", False), ("imu", True), ("kin", True), ("imu", "in-kernel"), ("kin", "in-kernel")])
def test_process_scan_with_interleaved_queue(kind, update_map):
    """KILO::process's second lambda: ~50 buckets, every sample with stamp < bucket time applied first
    (KILO.cc:379-390). update_map=False runs the fused persistent kernel, True the per-bucket kernels, "in-kernel" the
    persistent kernel with UpdateVoxelMap inside (queue drain, predict, update and insert of all buckets in ONE launch)."""
    in_kernel = update_map == "in-kernel"
    update_map = bool(update_map)
    cfg, blob, scans = scenes.box_scene(batch=1, streaming=True, stream0=1100)
    pts, offs, times = synth.bucketize(scans[0], begin_time=20.0)
    x0 = tp._moving_state(); P0 = abi.init_cov(1)
    clk = np.zeros(1, abi.CLOCK_DTYPE); clk["last_predict_time"] = 19.995; clk["last_update_time"] = 19.995
    meas = synth.imu_stream(19.996, 20.12) if kind == "imu" else synth.kinimu_stream(19.996, 20.12)
    Q = abi.process_cov_Q(cfg)
    o = lko.Oracle(cfg); o.map_import(blob); o.set_filter(x0, P0, Q, clk)
    o.set_options(gain_mode=lko.GAIN_INFORMATION, update_map=update_map, imu_mode_only=(kind == "imu"), gravity=9.81, acc_norm=9.79)
    ro = o.process_scan(20.0, pts, imu=meas if kind == "imu" else None, kin=meas if kind == "kin" else None)
    xo, Po, _, co = o.get_filter()
    eng = Engine(cfg); eng.map_upload(blob)
    eng.set_param("fused_insert", 1 if in_kernel else 0)
    out = eng.process_scan(x0, P0, Q, clk, pts, offs, times, imu=meas if kind == "imu" else None,
                           kin=meas if kind == "kin" else None, gravity=9.81, acc_norm=9.79, update_map=update_map)
    assert out["n_consumed"] == ro["n_consumed"] > 30 and out["n_eff"] == ro["n_eff"] > 0
    assert scenes.rel_state_err(out["x"], xo, x0) < TOL and scenes.rel_cov_err(out["P"], Po) < TOL
    assert out["clk"].tobytes() == co.tobytes()
    np.testing.assert_allclose(out["world"][:, :3], ro["world"][:, :3], rtol=0, atol=1e-5)
