"""Per-phase %globaltimer trace of the residual kernel for one batch=1 step (latency diagnosis)."""
import sys, os, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("leg-kilo_b200/python","tests"): sys.path.insert(0, os.path.join(ROOT,p))
sys.path.insert(0, ROOT)
import bench
from legkilo_b200 import Engine, abi, lib, _p
w = bench.WORKLOADS["small"]; wl = bench.build_workload(w, 0, 16); cfg = wl["cfg"]
eng = Engine(cfg); eng.map_build(wl["map_world"], wl["map_body"])
eng.stage(wl["x0"], abi.init_cov(16), abi.process_cov_Q(cfg), np.zeros(16, abi.CLOCK_DTYPE), wl["pts"], wl["offs"], np.zeros(16))
for i in range(32): eng.run_range(i % 16, 1, iters=3)
eng.sync(); eng.set_param("trace", 1)
for scan in (3, 4):
    # the trace buffer holds the LAST residual launch of the run (iteration 3, warm L2)
    eng.run_range(scan, 1, iters=3); eng.sync()
    tr = np.zeros((1 << 16) * 8, np.uint64); lib().lk_debug_read(eng.h, 2, _p(tr), tr.nbytes)
    nb = 113; blk = tr[:nb * 8].reshape(nb, 8).astype(np.int64); tail = tr[nb * 8: nb * 8 + 8].astype(np.int64)
    t0 = blk[:, 0].min()
    print("blocks: start spread %.2f us; sc-load %.2f; points %.2f; reduce+ticket %.2f (medians, us)" % (
        (blk[:, 0].max() - t0) / 1e3, np.median(blk[:, 1] - blk[:, 0]) / 1e3, np.median(blk[:, 2] - blk[:, 1]) / 1e3, np.median(blk[:, 3] - blk[:, 2]) / 1e3))
    print("last block done at %.2f us; tail: start %.2f partials %.2f ->solve %.2f solve %.2f ->stateupd %.2f end %.2f" % (
        (blk[:, 3].max() - t0) / 1e3, (tail[0] - t0) / 1e3, (tail[1] - tail[0]) / 1e3, (tail[2] - tail[1]) / 1e3, (tail[3] - tail[2]) / 1e3, (tail[4] - tail[3]) / 1e3, (tail[7] - t0) / 1e3))
