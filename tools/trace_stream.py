"""Per-bucket %globaltimer trace of the persistent per-scan kernel in streaming mode (queue + map insert inside)."""
import sys, os, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("leg-kilo_b200/python","tests"): sys.path.insert(0, os.path.join(ROOT,p))
sys.path.insert(0, ROOT)
from legkilo_b200 import Engine, abi, lib, _p, synth
cfg = abi.CONFIGS["nclt"]; R, t = abi.extrinsics(cfg)
scene = synth.BoxScene(ground_half_extent=40.0); pw, pb = scene.map_points(ext_R=R, ext_t=t)
eng = Engine(cfg); eng.map_build(pw, pb)
for kv in sys.argv[1:]:
    k, v = kv.split("="); eng.set_param(k, float(v))
g = synth.rng(77); n = 8
rv = np.cumsum(2e-3 * g.standard_normal((n, 3)), 0); tv = np.cumsum(0.01 * g.standard_normal((n, 3)), 0) * np.array([1, 1, 0.1])
x = abi.default_states(1); P = abi.init_cov(1); clk = np.zeros(1, abi.CLOCK_DTYPE); Q = abi.process_cov_Q(cfg)
for i in range(n):
    sc = scene.scan(rotvec=rv[i], trans=tv[i], ext_R=R, ext_t=t, blind=cfg["blind"], stream=5000 + i, streaming=True, **synth.VLP16)
    t0 = 0.1 * i
    pts, offs, times = synth.bucketize(sc, begin_time=t0)
    meas = synth.imu_stream(t0 - 0.1 if i else -0.005, t0 + 0.1, 400.0, stream=9000 + i)
    meas = meas[meas["stamp"] > float(clk["last_update_time"][0]) - 1.0]
    if i == n - 2: eng.set_param("trace", 1)
    o = eng.process_scan(x, P, Q, clk, pts, offs, times, imu=meas, gravity=9.81, acc_norm=9.79, iters=1, update_map=True)
    x, P, clk = o["x"], o["P"].reshape(1, 900), o["clk"]
    if i >= n - 2:
        tr = np.zeros((1 << 16) * 8, np.uint64); lib().lk_debug_read(eng.h, 2, _p(tr), tr.nbytes)
        nb = 148
        st = tr[nb * 32: nb * 32 + 64 * 8].reshape(64, 8).astype(np.int64)
        nbk = min(len(times), 64)
        st = st[:nbk]
        d = np.diff(st, axis=1) / 1e3
        names = ["queue drain", "predict", "pass", "all-reduce", "solve+cov", "P1+barrier", "P2+barrier"]
        print("scan %d: %d buckets, %d samples; per bucket us (median / mean / max):" % (i, len(times), len(meas)))
        for j, nm in enumerate(names): print("   %-12s %7.2f %7.2f %7.2f" % (nm, np.median(d[:, j]), d[:, j].mean(), d[:, j].max()))
        tot = (st[:, 7] - st[:, 0]) / 1e3
        gap = (st[1:, 0] - st[:-1, 7]) / 1e3
        print("   bucket total %.2f (sum %.1f), between buckets %.2f, kernel span %.1f us" % (np.median(tot), tot.sum(), np.median(gap), (st[-1, 7] - st[0, 0]) / 1e3))
