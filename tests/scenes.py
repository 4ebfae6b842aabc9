"""Shared fixtures-as-functions for the parity tests: scenes from legkilo_b200.synth fed, unchanged,
to the CPU oracle and to the CUDA library."""
import numpy as np

import lko
from legkilo_b200 import abi, synth


def planar_scene(cfg_name="leg_fusion", n=2048, half_extent=20.0, seed_stream=2, rotvec=(2e-3, -1e-3, 3e-3),
                 trans=(0.02, -0.01, 0.03)):
    cfg = abi.CONFIGS[cfg_name]
    R, t = abi.extrinsics(cfg)
    pw, pb = synth.planar_map_points(half_extent=half_extent, ext_R=R, ext_t=t)
    o = lko.Oracle(cfg)
    o.build_voxel_map(pw, pb)
    blob = o.map_export()
    pts = synth.planar_scan(n=n, ext_R=R, ext_t=t, stream=seed_stream, rotvec=rotvec, trans=trans)
    return cfg, blob, pts


def box_scene(cfg_name="leg_fusion", lidar=None, ground_half_extent=20.0, batch=1, rot_sigma=2e-3, trans_sigma=0.02,
              streaming=False, stream0=100):
    """Box room, map built by the ORACLE's BuildVoxelMap over a small ground patch."""
    cfg = abi.CONFIGS[cfg_name]
    R, t = abi.extrinsics(cfg)
    sc = synth.BoxScene(ground_half_extent=ground_half_extent)
    pw, pb = sc.map_points(ext_R=R, ext_t=t)
    o = lko.Oracle(cfg)
    o.build_voxel_map(pw, pb)
    blob = o.map_export()
    lidar = lidar or synth.VLP16
    rv, tv = synth.random_poses(batch, rot_sigma, trans_sigma, stream=stream0)
    scans = [sc.scan(rotvec=rv[i], trans=tv[i], ext_R=R, ext_t=t, blind=cfg["blind"], stream=stream0 + 1 + i,
                     streaming=streaming, **lidar) for i in range(batch)]
    return cfg, blob, scans


def rel_state_err(x_a, x_b, x_prior):
    """||x_a [-] x_b||_inf / max(||x_b [-] x_prior||_inf, eps)  (SURVEY §8d pose error)."""
    num = np.abs(lko.boxminus(x_a, x_b)).max()
    den = max(np.abs(lko.boxminus(x_b, x_prior)).max(), 1e-12)
    return num / den


def rel_cov_err(P_a, P_b):
    P_a = np.asarray(P_a).reshape(30, 30)
    P_b = np.asarray(P_b).reshape(30, 30)
    return np.abs(P_a - P_b).max() / np.abs(P_b).max()
