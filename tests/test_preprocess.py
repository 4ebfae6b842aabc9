"""PointCloud2 wire decode (lidar_processing.cc:25-108) and voxel-grid down-sampling +
curvature sort + bucketing (KILO.cc:356-378). CPU: oracle vs a numpy statement of the same rules; GPU:
device vs oracle, bit-exact (float / integer work)."""
import numpy as np
import pytest

import lko
from legkilo_b200 import abi, synth


def _raw_cloud(lidar_type, n=6000, seed=1):
    g = synth.rng(seed)
    dt = abi.PC2_DTYPES[lidar_type]
    a = np.zeros(n, dt)
    r = g.uniform(0.3, 40.0, n); az = np.linspace(0, 2 * np.pi, n, endpoint=False); el = g.uniform(-0.3, 0.3, n)
    a["x"] = r * np.cos(el) * np.cos(az); a["y"] = r * np.cos(el) * np.sin(az); a["z"] = r * np.sin(el)
    a["intensity"] = g.uniform(0, 255, n)
    t = np.linspace(0, 0.1, n)
    if lidar_type == 1:
        a["time"] = t * 1e6  # microseconds, time_scale 1e-6 (nclt.yaml)
        scale = 1e-6
    elif lidar_type == 2:
        a["t"] = (t * 1e9).astype(np.uint32)  # nanoseconds, time_scale 1e-9 (diter.yaml)
        scale = 1e-9
    else:
        a["timestamp"] = 1.7e9 + t  # absolute seconds (hilti.yaml)
        scale = 1.0
    return a, scale


def _numpy_decode(a, lidar_type, blind, filter_num, scale):
    x, y, z = a["x"], a["y"], a["z"]
    r2 = (x * x + y * y) + z * z  # float32 arithmetic, left to right
    keep = (np.arange(len(a)) % filter_num == 0) & ~(np.float32(blind) * np.float32(blind) > r2)
    if lidar_type == 3:
        first = scale * a["timestamp"][0]
        cur = scale * a["timestamp"]
        curv = (np.round((cur - first) * np.float64(np.float32(500.0))) / np.float64(np.float32(500.0))).astype(np.float32)
    else:
        tt = a["time"].astype(np.float64) if lidar_type == 1 else a["t"].astype(np.float64)
        first = np.float32(scale * tt[0]); cur = (scale * tt).astype(np.float32)
        curv = np.round((cur - first) * np.float32(500.0)) / np.float32(500.0)
    out = np.stack([x, y, z, curv.astype(np.float32)], 1)[keep]
    return out.astype(np.float32), a["intensity"][keep]


@pytest.mark.parametrize("lidar_type", [1, 2, 3])
def test_oracle_decode_vs_numpy(lidar_type):
    a, scale = _raw_cloud(lidar_type)
    pts, inten, ft, lt = lko.decode_pointcloud2(a.view(np.uint8), abi.pc2_layout(lidar_type), 1.5, 3, scale)
    ref, iref = _numpy_decode(a, lidar_type, 1.5, 3, scale)
    assert pts.shape == ref.shape and 0 < len(pts) < len(a) // 3 + 1
    # np.round is half-to-even, std::round half-away-from-zero: identical unless exactly on .5
    np.testing.assert_array_equal(pts[:, :3], ref[:, :3])
    assert np.abs(pts[:, 3] - ref[:, 3]).max() <= 0.002 + 1e-7 and (pts[:, 3] != ref[:, 3]).mean() < 0.01
    np.testing.assert_array_equal(inten, iref)
    q = pts[:, 3] * 500.0
    assert np.abs(q - np.round(q)).max() < 1e-3  # multiples of 2 ms (lidar_processing.cc:48)


def test_oracle_voxel_grid_properties():
    cfg = abi.CONFIGS["leg_fusion"]; R, t = abi.extrinsics(cfg)
    s = synth.BoxScene(ground_half_extent=20.0).scan(rotvec=[0, 0, 0], trans=[0, 0, 0], ext_R=R, ext_t=t, blind=1.5, streaming=True,
                                                     **synth.VLP16)
    out, offs, curv = lko.preprocess_scan(s, 0.3)
    assert 0 < len(out) < len(s)
    # one output per occupied leaf (leaf index = floor(p / leaf) relative to the cloud's minimum)
    inv = np.float32(1.0) / np.float32(0.3)
    ijk = np.floor(s[:, :3] * inv).astype(np.int64)
    assert len(out) == len(np.unique(ijk, axis=0))
    # every centroid stays inside its leaf (up to float rounding) and the mean of means is the mean
    np.testing.assert_allclose(out[:, :3].mean(0), np.array([s[(ijk == u).all(1), :3].mean(0) for u in np.unique(ijk, axis=0)]).mean(0), atol=1e-4)
    assert np.all(np.diff(out[:, 3]) >= 0) and offs[0] == 0 and offs[-1] == len(out)
    for b in range(len(curv)):
        assert np.all(out[offs[b]:offs[b + 1], 3] == curv[b])
    assert np.all(np.diff(curv) > 0)
    # a cloud that already has one point per leaf is only re-ordered
    out2, _, _ = lko.preprocess_scan(out, 0.3)
    assert len(out2) == len(out)


@pytest.mark.gpu
@pytest.mark.parametrize("lidar_type", [1, 2, 3])
def test_gpu_decode_bit_exact(lidar_type):
    from legkilo_b200 import Engine
    a, scale = _raw_cloud(lidar_type, n=50000, seed=7)
    ref = lko.decode_pointcloud2(a.view(np.uint8), abi.pc2_layout(lidar_type), 1.5, 3, scale)
    got = Engine(abi.CONFIGS["leg_fusion"]).decode_pointcloud2(a.view(np.uint8), abi.pc2_layout(lidar_type), 1.5, 3, scale)
    np.testing.assert_array_equal(got[0], ref[0]); np.testing.assert_array_equal(got[1], ref[1])
    assert got[2] == ref[2] and got[3] == ref[3]


@pytest.mark.gpu
@pytest.mark.parametrize("leaf,lidar", [(0.3, "VLP16"), (0.5, "OS64")])
def test_gpu_voxel_grid_sort_bucket_bit_exact(leaf, lidar):
    from legkilo_b200 import Engine
    cfg = abi.CONFIGS["leg_fusion"]; R, t = abi.extrinsics(cfg)
    s = synth.BoxScene(ground_half_extent=20.0).scan(rotvec=[1e-3, 2e-3, 0], trans=[0.1, 0, 0], ext_R=R, ext_t=t, blind=1.5,
                                                     streaming=True, **getattr(synth, lidar))
    s[::997, 0] = np.nan  # non-finite points are skipped by the filter
    ref = lko.preprocess_scan(s, leaf)
    got = Engine(cfg).preprocess_scan(s, leaf)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)
    assert len(got[0]) > 1000 and len(got[2]) > 40
