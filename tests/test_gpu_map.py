"""Device-side map construction (BuildVoxelMap, init_plane, init/cut_octo_tree) against the oracle."""
import numpy as np
import pytest

import lko
import mapcmp
import scenes
from legkilo_b200 import Engine, abi, synth

pytestmark = pytest.mark.gpu


def _both(cfg, pw, pb, R=None, rot_cov=None, pos_cov=None):
    o = lko.Oracle(cfg)
    o.build_voxel_map(pw, pb, R, rot_cov, pos_cov)
    eng = Engine(cfg)
    eng.map_build(pw, pb, R, rot_cov, pos_cov)
    return o.map_export(), eng.map_download(), eng


def test_build_planar_map_matches_oracle():
    cfg = abi.CONFIGS["leg_fusion"]
    R, t = abi.extrinsics(cfg)
    pw, pb = synth.planar_map_points(half_extent=10.0, ext_R=R, ext_t=t)
    a, b, eng = _both(cfg, pw, pb)
    st = mapcmp.compare_blobs(a, b)
    assert st["planes"] == 1600 and st["points"] == len(pw)
    s = eng.map_stats()
    assert s["roots"] == 1600 and s["planes"] == 1600 and s["points"] == len(pw)


def test_build_box_room_matches_oracle():
    cfg = abi.CONFIGS["diter"]
    R, t = abi.extrinsics(cfg)
    pw, pb = synth.BoxScene(ground_half_extent=18.0).map_points(ext_R=R, ext_t=t)
    a, b, _ = _both(cfg, pw, pb)
    st = mapcmp.compare_blobs(a, b)
    assert st["planes"] > 5000


def test_build_cluttered_scene_subdivides():
    """Dense non-planar clutter: roots fail the plane test, get cut into octants (2 layers),
    big leaves freeze (> max_points_num) — exercises cut_octo_tree and the freeze rules."""
    cfg = abi.CONFIGS["leg_fusion"]
    g = synth.rng(77)
    n = 60000
    pw = np.concatenate([
        g.uniform(-4, 4, (n // 2, 3)),                                   # volumetric clutter
        np.c_[g.uniform(-4, 4, (n // 4, 2)), 0.13 + 0.002 * g.standard_normal(n // 4)],  # a thin slab inside it
        g.uniform(4, 6, (n // 4, 3)) * np.array([1, 1, 0.05])]).astype(np.float32)
    pb = pw.copy()
    pb[:, 2] -= 0.2
    rot = synth.exp_so3([0.01, -0.02, 0.03])
    rc = np.diag([1e-6, 2e-6, 3e-6]); pc = np.diag([4e-6, 5e-6, 6e-6])
    a, b, _ = _both(cfg, pw, pb, rot, rc, pc)
    st = mapcmp.compare_blobs(a, b)
    assert st["interior"] > 100 and st["planes"] > 100


def test_map_upload_download_roundtrip():
    cfg, blob, _ = scenes.planar_scene(half_extent=6.0)
    eng = Engine(cfg)
    eng.map_upload(blob)
    st = mapcmp.compare_blobs(blob, eng.map_download(), rtol=1e-15)
    assert st["planes"] == 576
