"""Where the end-to-end time of a one-scan lk_scan_update goes: the same call with parts of it removed.
Usage (GPU box): python tools/e2e_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bench import abi  # noqa: E402
from legkilo_b200 import Engine, pinned_empty, _p, lib  # noqa: E402

w = bench.WORKLOADS["leg_fusion_b1"]
wl = bench.build_workload(w, 0, 64)
cfg = wl["cfg"]
eng = Engine(cfg)
eng.map_build(wl["map_world"], wl["map_body"])
Q = abi.process_cov_Q(cfg)
x0 = wl["x0"]
offs = wl["offs"]
maxp = int(np.diff(offs).max())
h_pts = pinned_empty((maxp, 4), np.float32)
h_world = pinned_empty((maxp, 4), np.float32)
pg_pts = np.zeros((maxp, 4), np.float32)
pg_world = np.zeros((maxp, 4), np.float32)


def loop(name, npts=None, world=True, pinned=True, reps=200, iters=3, params=None):
    for k, v in (params or {}).items():
        eng.set_param(k, v)
    t = 0.0
    for i in range(reps + 5):
        g = i % 64
        o0, o1 = int(offs[g]), int(offs[g + 1])
        n = o1 - o0 if npts is None else min(npts, o1 - o0)
        src = h_pts if pinned else pg_pts
        dst = (h_world if pinned else pg_world) if world else None
        src[:n] = wl["pts"][o0:o0 + n]
        so = np.array([0, n], np.uint32); sbp = np.array([0, 1], np.uint32); bt = np.zeros(1)
        xi, Pi, ci = x0[g:g + 1].copy(), abi.init_cov(1), np.zeros(1, abi.CLOCK_DTYPE)
        ne = np.zeros(1, np.uint32)
        cargs = (eng.h, 1, _p(xi), _p(Pi), _p(Q), _p(ci), _p(src), _p(so), _p(sbp), _p(so), _p(bt), iters, 0, _p(dst), _p(ne))
        fn = lib().lk_scan_update
        t0 = time.perf_counter(); rc = fn(*cargs); t1 = time.perf_counter()
        assert rc == 0
        if i >= 5:
            t += t1 - t0
    print("%-46s %7.1f us / call" % (name, t / reps * 1e6), flush=True)
    for k in (params or {}):
        eng.set_param(k, 1)


loop("full (direct mode, 28.8k pts, world out)")
loop("no world cloud out", world=False)
loop("256 points (one block), world out", npts=256)
loop("256 points, 1 iteration", npts=256, iters=1)
loop("28.8k pts, 1 iteration", iters=1)
loop("staged (direct_io = 0), pinned", params=dict(direct_io=0))
loop("pageable buffers", pinned=False)
# resident inputs: launch + kernel + sync only
nsc = 64
eng.stage(x0, abi.init_cov(nsc), Q, np.zeros(nsc, abi.CLOCK_DTYPE), wl["pts"], offs, np.zeros(nsc))
for it in (3, 1):
    t = 0.0
    for i in range(205):
        t0 = time.perf_counter(); eng.run_range(i % nsc, 1, iters=it); eng.sync(); t1 = time.perf_counter()
        if i >= 5:
            t += t1 - t0
    print("%-46s %7.1f us / call" % ("resident: run_range + sync, iters=%d" % it, t / 200 * 1e6), flush=True)
