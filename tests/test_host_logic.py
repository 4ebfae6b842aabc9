"""Host-side logic that needs no GPU: canonical bucket ordering (a1), synthetic scene invariants,
multi-GPU sharding (incl. a world_size-2 gloo run)."""
import os
import subprocess
import sys

import numpy as np

from legkilo_b200 import abi, shard, synth


def test_bucketize_is_stable_and_matches_reference_rule():
    g = synth.rng(3)
    pts = np.zeros((1000, 4), np.float32)
    pts[:, :3] = g.normal(size=(1000, 3))
    pts[:, 3] = (np.round(g.uniform(0, 0.1, 1000) / 0.002) * 0.002).astype(np.float32)  # lidar_processing.cc:48
    pts[:, 0] = np.arange(1000)  # tag with the input position
    s, offs, times = synth.bucketize(pts, begin_time=7.0)
    assert offs[0] == 0 and offs[-1] == 1000 and len(times) == len(offs) - 1 <= 51
    for b in range(len(times)):
        blk = s[offs[b]:offs[b + 1]]
        assert np.all(blk[:, 3] == blk[0, 3])                # maximal equal-curvature run (KILO.cc:377-378)
        assert np.all(np.diff(blk[:, 0]) > 0)                # stable: input order kept inside a bucket
        assert times[b] == 7.0 + float(blk[0, 3])            # begin_time + curvature (KILO.cc:376)
    assert np.all(np.diff(s[:, 3]) >= 0)


def test_box_scene_scan_shapes():
    cfg = abi.CONFIGS["leg_fusion"]; R, t = abi.extrinsics(cfg)
    sc = synth.BoxScene(ground_half_extent=20.0)
    s = sc.scan(rotvec=[1e-3, 0, 0], trans=[0.01, 0, 0], ext_R=R, ext_t=t, blind=cfg["blind"], streaming=True, **synth.VLP16)
    assert 28000 < len(s) <= 28800 and s.dtype == np.float32
    assert np.all(np.linalg.norm(s[:, :3], axis=1) >= 1.4)   # blind zone
    q = s[:, 3] / 0.002
    assert np.allclose(q, np.round(q), atol=1e-3) and s[:, 3].max() <= 0.1001
    pw, pb = sc.map_points(ext_R=R, ext_t=t)
    np.testing.assert_allclose(pw, pb + t.astype(np.float32), atol=1e-5)  # seen from the (single) room centre


def test_shard_ranges_cover_exactly_once():
    for n in (0, 1, 7, 128, 1024, 1025):
        for w in (1, 2, 4, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_rebases_offsets():
    offs = np.array([0, 5, 9, 20, 26], np.uint32)
    pts = np.arange(26 * 4, dtype=np.float32).reshape(26, 4)
    x = abi.default_states(4); P = abi.init_cov(4); clk = np.zeros(4, abi.CLOCK_DTYPE)
    s = shard.shard_batch(1, 2, x, P, clk, pts, offs, np.arange(4.0))
    assert s["lo"] == 2 and s["hi"] == 4
    np.testing.assert_array_equal(s["scan_offsets"], [0, 11, 17])
    np.testing.assert_array_equal(s["pts"], pts[9:26])
    np.testing.assert_array_equal(s["bucket_times"], [2.0, 3.0])


_GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from legkilo_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 37
lo, hi = shard.shard_range(n, rank, world)
owned = torch.zeros(n, dtype=torch.int64); owned[lo:hi] = 1
dist.all_reduce(owned)                       # every scan owned exactly once across ranks
t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)   # per-rank device time; job time = max over ranks
dist.all_reduce(t, op=dist.ReduceOp.MAX)
work = torch.tensor([float(hi - lo)], dtype=torch.float64); dist.all_reduce(work)
if rank == 0:
    assert bool((owned == 1).all()) and float(work) == n and abs(float(t) - 0.010 * world) < 1e-12
    print("GLOO_OK", world)
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding(tmp_path):
    """The N > 1 plumbing of bench.py (shard -> run -> max-over-ranks time, summed work) on CPU with gloo."""
    w = tmp_path / "worker.py"
    w.write_text(_GLOO_WORKER)
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "leg-kilo_b200", "python")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29541", str(w), pkg], capture_output=True, text=True, timeout=240, env=env)
    assert "GLOO_OK 2" in out.stdout, out.stdout + out.stderr


def test_bench_clock_sampler_windows_and_traffic_lookup():
    """bench.py's host-side bookkeeping: clock samples are taken from the timed region, or — when that is shorter than the sampling
    period — from the last samples before its end; throttle reasons are collected; the throughput family's traffic is the sum of its two
    committed ncu captures."""
    import time

    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import bench

    class FakeProc:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    now = time.time()
    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.rows = [(now - 1.0, "210, 1965, Not Active, Not Active, Not Active, Not Active"),      # idle, long before
              (now - 0.5, "1965, 1965, Not Active, Not Active, Not Active, Not Active"),    # warm-up
              (now - 0.01, "1950, 1965, Not Active, Not Active, Not Active, Active")]       # inside the region
    s.t_begin = now - 0.02
    out = s.stop()
    assert out["window"] == "timed region" and out["samples"] == 1 and out["sm_mhz"] == 1950.0 and out["reasons"] == ["sw_power_cap"]
    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.rows = [(now - 1.0, "210, 1965, Not Active, Not Active, Not Active, Not Active"),
              (now - 0.5, "1965, 1965, Not Active, Not Active, Not Active, Not Active")]
    s.t_begin = time.time()  # a region with no sample of its own
    out = s.stop()
    assert out["window"].startswith("last samples") and out["samples"] == 2 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    total, note = bench.ncu_traffic_family()
    a, _ = bench.ncu_traffic("k_residual_stream2")
    b, _ = bench.ncu_traffic("k_residual_fallback")
    assert total == a + b and "k_residual_fallback" in note
    assert bench.ncu_traffic("k_scan_fused")[0] > 1e6
