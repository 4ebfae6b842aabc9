"""the Eigen-typed facade (leg-kilo_b200/host/legkilo_facade.hpp) is header-only and meant to be
compiled inside the reference's catkin workspace. This image has no Eigen, so the header is type-checked and its
templates instantiated against a ~50-line stand-in (tests/stubs/Eigen/Dense) with mock State / Config types that have the
reference's member names (eskf.h:15-32, :49-65, voxel_map.h:41-57) — a syntax / interface check, not a numerical one.
Also covers the two host-side helpers of the same row: lk_tum_line (no GPU needed) and lk_map_slide (GPU)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
#include <vector>
#include "legkilo_facade.hpp"
#if !__has_include(<Eigen/Dense>)
#error "stub Eigen not found"
#endif
using namespace legkilo::b200;
struct State {  // legkilo::State (eskf.h:15-32)
    Mat3D rot_; Vec3D pos_, vel_, ba_, bw_, grav_, imu_a_, imu_w_, bv_, contact_;
};
struct EskfConfig { double v[14]; };  // ESKF::Config (eskf.h:49-65): 14 doubles
struct VoxelMapConfig {  // VoxelMapConfig (voxel_map.h:41-57)
    double max_voxel_size_, planner_threshold_, beam_err_, dept_err_, sigma_num_;
    int max_layer_, max_points_num_;
    std::vector<int> layer_init_num_;
};
int main() {
    EskfConfig ec{}; VoxelMapConfig mc{}; mc.layer_init_num_ = {5, 5, 5, 5, 5};
    Mat3D Re; Vec3D te;
    Core core(ec, mc, Re, te, 0);
    State s; StateCov P; double tp = 0, tu = 0;
    std::vector<float> xyzt, world; std::vector<uint32_t> bo{0}; std::vector<double> bt;
    std::vector<lk_imu_meas> imu; std::vector<lk_kinimu_meas> kin;
    size_t n = core.processScan(s, P, tp, tu, xyzt, bo, bt, imu, kin, 9.81, 9.79, world);
    core.BuildVoxelMap(nullptr, nullptr, 0, Re, Re, Re);
    uint64_t removed = 0;
    bool slid = core.mapSliding(te, &removed);
    std::string line = core.tumLine(0.0, s);
    lk_state x = toAbi(s); fromAbi(x, s);
    return (int)n + (int)slid + (int)line.size();
}
'''


def test_facade_header_type_checks_against_stub_eigen():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "facade_driver.cpp")
        with open(src, "w") as f:
            f.write(DRIVER)
        cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "stubs"),
               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "leg-kilo_b200", "host"), src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def _quat_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_tum_line_format_and_quaternion_branches():
    from legkilo_b200 import synth, tum_line
    # one rotation per branch of Eigen's Quaterniond(Matrix3d): trace > 0, and the largest diagonal entry being 0 / 1 / 2
    cases = [synth.exp_so3([0.1, -0.2, 0.3]), synth.exp_so3([3.0, 0.05, -0.02]), synth.exp_so3([0.03, 3.05, 0.01]),
             synth.exp_so3([-0.02, 0.04, 3.1])]
    seen = set()
    for R in cases:
        tr = np.trace(R)
        seen.add("t" if tr > 0 else int(np.argmax(np.diag(R))))
        line = tum_line(1234.5678901234, R, [1.5, -2.25, 0.125])
        assert line.endswith("\n")
        f = line.split()
        assert len(f) == 8 and all(len(v.split(".")[1]) == 9 for v in f)
        assert f[0] == "1234.567890123" and f[1] == "1.500000000" and f[2] == "-2.250000000"
        q = np.array([float(v) for v in f[4:]])
        assert abs(np.linalg.norm(q) - 1) < 1e-8
        np.testing.assert_allclose(_quat_to_rot(q), R, atol=5e-9)
        if tr > 0:
            assert q[3] > 0
    assert seen == {"t", 0, 1, 2}


@pytest.mark.gpu
def test_map_slide_drops_roots_outside_the_window():
    import scenes
    from legkilo_b200 import Engine, abi
    cfg = dict(abi.CONFIGS["leg_fusion"], half_map_size=10, sliding_thresh=8.0)  # window of +-10 voxels = +-5 m
    _, blob, scans = scenes.box_scene(batch=1)
    hd, roots, nodes, aux, pts = abi.parse_map_blob(blob)
    eng = Engine(cfg)
    eng.map_upload(blob)
    n0 = eng.map_stats()["roots"]
    assert eng.map_slide([3.0, 0.0, 0.0]) == (False, 0)  # closer than sliding_thresh to the last slide position (the origin)
    slid, removed = eng.map_slide([9.0, 1.0, 0.2])
    k = np.floor(np.array([9.0, 1.0, 0.2]) / 0.5).astype(int)
    keep = np.all((roots["key"] <= k + 10) & (roots["key"] >= k - 10), axis=1)
    assert slid and removed == int((~keep).sum()) > 0
    assert eng.map_stats()["roots"] == n0 - removed == int(keep.sum())
    _, r2, _, _, _ = abi.parse_map_blob(eng.map_download())
    assert {tuple(x) for x in r2["key"].tolist()} == {tuple(x) for x in roots["key"][keep].tolist()}
    assert eng.map_slide([9.5, 1.0, 0.2]) == (False, 0)  # measured from the position of the last slide now
    # the surviving map still serves the hot path
    x0 = abi.default_states(1); x0["pos"][0] = (0.0, 0.0, 0.0)
    out = eng.scan_update(x0, abi.init_cov(1), abi.process_cov_Q(cfg), np.zeros(1, abi.CLOCK_DTYPE), scans[0], [0, len(scans[0])], [0.0])
    assert 0 < int(out["n_eff"][0]) < len(scans[0])
