"""Per-phase %globaltimer trace of the fused per-scan kernel (latency diagnosis)."""
import sys, os, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("leg-kilo_b200/python","tests"): sys.path.insert(0, os.path.join(ROOT,p))
sys.path.insert(0, ROOT)
import bench
from legkilo_b200 import Engine, abi, lib, _p
name = sys.argv[1] if len(sys.argv) > 1 else "small"
w = bench.WORKLOADS[name]; ring = 16 if name == "small" else 128
wl = bench.build_workload(w, 0, ring); cfg = wl["cfg"]
eng = Engine(cfg); eng.map_build(wl["map_world"], wl["map_body"])
CL = True
for kv in sys.argv[2:]:
    k, v = kv.split("="); eng.set_param(k, float(v))
    if k == "cluster" and float(v) == 0: CL = False
eng.stage(wl["x0"], abi.init_cov(ring), abi.process_cov_Q(cfg), np.zeros(ring, abi.CLOCK_DTYPE), wl["pts"], wl["offs"], np.zeros(ring))
for i in range(ring * 2): eng.run_range(i % ring, 1, iters=3)
eng.sync(); eng.set_param("trace", 1)
us = lambda v: v / 1e3
for scan in (3, 4, 5):
    eng.run_range(scan, 1, iters=3); eng.sync()
    tr = np.zeros((1 << 16) * 8, np.uint64); lib().lk_debug_read(eng.h, 2, _p(tr), tr.nbytes)
    nb = int((wl["offs"][scan + 1] - wl["offs"][scan] + 255) // 256)
    nbg = nb
    b = tr[:nb * 32].reshape(nb, 32).astype(np.int64); t0 = b[:, 0].min()
    print("scan %d (%d blocks): start spread %.2f, load filter + init %.2f" % (scan, nb, us(b[:, 0].max() - t0), us(np.median(b[:, 1] - b[:, 0]))))
    prev = b[:, 1]
    for it in range(3):
        pts = b[:, 2 + 4 * it] - prev; ar = b[:, 3 + 4 * it] - b[:, 2 + 4 * it]; sol = b[:, 4 + 4 * it] - b[:, 3 + 4 * it]
        last_in = b[:, 2 + 4 * it].max()
        print("  it%d: points med %.2f max %.2f (last block in at %.2f) | all-reduce wait med %.2f, done %.2f after the last block (at %.2f) | solve %.2f" % (
            it, us(np.median(pts)), us(pts.max()), us(last_in - t0), us(np.median(ar)), us(np.median(b[:, 3 + 4 * it]) - last_in), us(np.median(b[:, 3 + 4 * it]) - t0), us(np.median(sol))))
        prev = b[:, 4 + 4 * it]
    print("  reproject (+cov on block 0) med %.2f, block 0 %.2f; end at med %.2f max %.2f us" % (us(np.median(b[:, 30] - prev)), us(b[0, 30] - prev[0]), us(np.median(b[:, 31]) - t0), us(b[:, 31].max() - t0)))
