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
eng.stage(wl["x0"], abi.init_cov(ring), abi.process_cov_Q(cfg), np.zeros(ring, abi.CLOCK_DTYPE), wl["pts"], wl["offs"], np.zeros(ring))
for i in range(ring * 2): eng.run_range(i % ring, 1, iters=3)
eng.sync(); eng.set_param("trace", 1)
for scan in (3, 4, 5):
    eng.run_range(scan, 1, iters=3); eng.sync()
    tr = np.zeros((1 << 16) * 8, np.uint64); lib().lk_debug_read(eng.h, 2, _p(tr), tr.nbytes)
    nb = 113; b = tr[:nb * 32].reshape(nb, 32).astype(np.int64); t0 = b[:, 0].min()
    us = lambda v: v / 1e3
    print("scan %d: start spread %.2f, load filter %.2f" % (scan, us(b[:, 0].max() - t0), us(np.median(b[:, 1] - b[:, 0]))))
    prev = b[:, 1]
    for it in range(3):
        pts = b[:, 2 + 4 * it] - prev; bar = b[:, 3 + 4 * it] - b[:, 2 + 4 * it]; red = b[:, 4 + 4 * it] - b[:, 3 + 4 * it]; sol = b[:, 5 + 4 * it] - b[:, 4 + 4 * it]
        print("  it%d: points med %.2f max %.2f | barrier wait med %.2f (released at %.2f) | partial-sum %.2f | solve %.2f" % (
            it, us(np.median(pts)), us(pts.max()), us(np.median(bar)), us(b[:, 3 + 4 * it].max() - t0), us(np.median(red)), us(np.median(sol))))
        prev = b[:, 5 + 4 * it]
    print("  reproject %.2f, end at %.2f us" % (us(np.median(b[:, 30] - prev)), us(b[:, 31].max() - t0)))
    pt = tr[nb * 64: nb * 64 + nb * 64].reshape(nb, 8, 8).astype(np.int64)   # [block][warp][slot], iteration 1
    d = np.diff(pt, axis=2) / 1e3
    names = ["load+xform+probe", "tma issue+wait", "read smem", "eval", "sync", "fallback", "accum"]
    print("  pass phases (it1, median over warps / max): " + ", ".join("%s %.2f/%.2f" % (n, np.median(d[:, :, i]), d[:, :, i].max()) for i, n in enumerate(names)))
    print("  fallback entries/block: n/a; pass total med %.2f max %.2f" % (np.median(pt[:, :, 7] - pt[:, :, 0]) / 1e3, (pt[:, :, 7] - pt[:, :, 0]).max() / 1e3))
    sv = tr[2 * nb * 64: 3 * nb * 64].reshape(nb, 64).astype(np.int64)
    for it, o in ((1, 0), (2, 8)):
        dd = np.diff(sv[:, o:o + 8], axis=1) / 1965.0
        print("  solve it%d (us, median over blocks): " % it + ", ".join("%s %.2f" % (n, np.median(dd[:, k])) for k, n in enumerate(["A", "M cols", "gauss-jordan", "y/W/delta", "exp+state", "syncthreads", "P update"])))
