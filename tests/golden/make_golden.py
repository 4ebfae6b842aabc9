"""Regenerates tests/golden/*.npz from the CPU oracle (run here, in the build container; the GPU box
only reads the committed fixtures). The reference itself ships no golden vectors and cannot be
compiled or imported (C++/Eigen/PCL/ROS), so these are oracle-made: see oracle/README.md."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in ("leg-kilo_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import lko  # noqa: E402
import scenes  # noqa: E402
from legkilo_b200 import abi, synth  # noqa: E402


def config1():
    cfg, blob, pts = scenes.planar_scene()
    x0 = abi.default_states(1); P0 = abi.init_cov(1)
    o = lko.Oracle(cfg); o.map_import(blob)
    o.set_filter(x0, P0, abi.process_cov_Q(cfg), np.zeros(1, abi.CLOCK_DTYPE))
    o.set_options(gain_mode=lko.GAIN_LITERAL, update_map=False)
    r = o.predict_update_point(0.0, pts)
    x, P, _, _ = o.get_filter()
    np.savez_compressed(os.path.join(HERE, "config1_planar.npz"), pts=pts, x=x.view(np.float64), P=P, n_eff=r["n_eff"],
                        world=r["world"])


def streaming():
    """One streaming box-room scan (50 buckets, map updated every bucket), information-form gain."""
    cfg, blob, scans = scenes.box_scene(batch=1, streaming=True, stream0=700, ground_half_extent=12.0)
    pts, offs, times = synth.bucketize(scans[0], begin_time=10.0)
    x0 = abi.default_states(1); x0["vel"][0] = (0.4, -0.2, 0.05); x0["imu_w"][0] = (0.02, -0.03, 0.15)
    P0 = abi.init_cov(1)
    clk = np.zeros(1, abi.CLOCK_DTYPE); clk["last_predict_time"] = 9.99; clk["last_update_time"] = 9.985
    o = lko.Oracle(cfg); o.map_import(blob)
    o.set_filter(x0, P0, abi.process_cov_Q(cfg), clk)
    o.set_options(gain_mode=lko.GAIN_INFORMATION, update_map=True)
    r = o.process_scan(10.0, pts)
    x, P, _, c = o.get_filter()
    _, roots, nodes, aux, mp = abi.parse_map_blob(o.map_export())
    np.savez_compressed(os.path.join(HERE, "streaming_box.npz"), x0=x0.view(np.float64), x=x.view(np.float64), P=P,
                        n_eff=r["n_eff"], clk=c.view(np.float64), n_roots=len(roots), n_nodes=len(nodes), n_points=len(mp),
                        n_planes=int((nodes["flags"] & 1).sum()), pts_sha=np.frombuffer(pts.tobytes()[:4096], np.uint8))


if __name__ == "__main__":
    config1()
    streaming()
    print("golden fixtures written")
