"""Throughput-family probe: old stream kernel vs the warp-specialised one; a watchdog thread dumps the
kernel's page-locked debug records if a run does not come back."""
import sys, os, time, threading, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("leg-kilo_b200/python","tests"): sys.path.insert(0, os.path.join(ROOT,p))
sys.path.insert(0, ROOT)
import bench
from legkilo_b200 import Engine, abi, lib, _p
ring = int(sys.argv[1]) if len(sys.argv) > 1 else 16
batches = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [ring]
w = bench.WORKLOADS["small"]
wl = bench.build_workload(w, 0, ring); cfg = wl["cfg"]
eng = Engine(cfg); eng.map_build(wl["map_world"], wl["map_body"])
if os.environ.get("WS_DEBUG", "1") == "1": eng.set_param("ws_debug", 1)
offs = wl["offs"]
state = dict(t=time.time(), busy=False)
def dog():
    while True:
        time.sleep(1.0)
        if state["busy"] and time.time() - state["t"] > 15:
            d = np.zeros(256 * 16 * 8, np.uint64); lib().lk_debug_read(eng.h, 4, _p(d), d.nbytes)
            d = d.reshape(256, 16, 8)
            names = {1: "prod wait empty", 2: "cons wait full", 10: "cons at barA", 11: "cons at barB", 12: "cons past barB", 20: "prod done", 0: "-"}
            shown = 0
            for b in range(256):
                if np.isin(d[b, :, 0], [1, 2, 10, 11]).any() and shown < 5:
                    shown += 1
                    print("block", b, flush=True)
                    for wi, r in enumerate(d[b]):
                        print("   w%d %s a=%d b=%d c=%d" % (wi, names.get(int(r[0]), str(int(r[0]))), r[1], r[2], r[3]), flush=True)
            print("stuck blocks:", int(np.isin(d[:, :, 0], [1, 2, 10, 11]).any(axis=1).sum()), flush=True)
            os._exit(3)
threading.Thread(target=dog, daemon=True).start()
for nb in batches:
    o1 = int(offs[nb]); res = {}
    WS = int(os.environ.get('WS_MODE', '1'))
    for ws in (0, WS):
        eng.set_param("ws", ws)
        eng.stage(wl["x0"][:nb], abi.init_cov(nb), abi.process_cov_Q(cfg), np.zeros(nb, abi.CLOCK_DTYPE), wl["pts"][:o1], offs[:nb + 1], np.zeros(nb))
        print("batch", nb, "ws", ws, "running", flush=True)
        state["t"] = time.time(); state["busy"] = True
        eng.run(iters=3); t = eng.last_timing()
        state["busy"] = False
        res[ws] = eng.fetch(want_world=False); print("   total_ms %.3f" % t["total_ms"], "n_eff", res[ws]["n_eff"][:4], flush=True)
    d = np.zeros(256 * 16 * 8, np.uint64)
    if os.environ.get("WS_DEBUG", "1") == "1": lib().lk_debug_read(eng.h, 4, _p(d), d.nbytes)
    d = d.reshape(256, 16, 8).astype(np.float64)
    pr = d[:148, :8]; co = d[:148, 8:]
    ghz = 1.965e3
    print("   producers (us, median over blocks x warps): total %.1f wait-empty %.1f resolve %.1f issue %.1f B %.1f items %.0f" % tuple(np.median(pr[:, :, k]) / (ghz if k != 7 else 1) for k in (2, 3, 4, 5, 6, 7)))
    print("   consumers (us): total %.1f wait-full %.1f eval %.1f barA %.1f fallback %.1f barB+reduce %.1f" % tuple(np.median(co[:, :, k]) / ghz for k in (2, 3, 4, 5, 6, 7)))
    dx = np.abs(res[0]["P"] - res[WS]["P"]).max() / np.abs(res[0]["P"]).max()
    print("   rel diff P", dx, "n_eff equal", np.array_equal(res[0]["n_eff"], res[WS]["n_eff"]), flush=True)
