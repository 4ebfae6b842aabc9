import sys, os, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("leg-kilo_b200/python","oracle","tests"): sys.path.insert(0, os.path.join(ROOT,p))
import scenes
from legkilo_b200 import Engine, abi, synth
which = sys.argv[1]
cfg, blob, scans = scenes.box_scene(batch=1, streaming=True, stream0=500)
pts, offs, times = synth.bucketize(scans[0], begin_time=100.0)
x0 = abi.default_states(1); x0["vel"][0]=(0.4,-0.2,0.05); P0 = abi.init_cov(1)
clk0 = np.zeros(1, abi.CLOCK_DTYPE); clk0["last_predict_time"]=99.99; clk0["last_update_time"]=99.985
eng = Engine(cfg); eng.map_upload(blob); Q = abi.process_cov_Q(cfg)
if len(sys.argv) > 2: eng.set_param("lane_cache", int(sys.argv[2]))
if which == "A":
    eng.set_param("fused", 0)
    out = eng.scan_update(x0,P0,Q,clk0,pts,[0,len(pts)],times,scan_bucket_ptr=[0,len(times)],bucket_offsets=offs,iters=1)
elif which == "B":
    n = int(offs[1])
    out = eng.scan_update(x0,P0,Q,clk0,pts[:n],[0,n],times[:1],iters=1)
elif which == "C":
    n = int(offs[2])
    out = eng.scan_update(x0,P0,Q,clk0,pts[:n],[0,n],times[:2],scan_bucket_ptr=[0,2],bucket_offsets=offs[:3],iters=1)
elif which == "E":
    out = eng.scan_update(x0,P0,Q,clk0,pts,[0,len(pts)],times,scan_bucket_ptr=[0,len(times)],bucket_offsets=offs,iters=1)
elif which == "F":
    x0["imu_w"][0] = (0.02, -0.03, 0.15); x0["imu_a"][0] = (0.3, 0.1, 9.7); x0["ba"][0] = (0.01, -0.02, 0.03)
    n = int(offs[2])
    out = eng.scan_update(x0,P0,Q,clk0,pts[:n],[0,n],times[:2],scan_bucket_ptr=[0,2],bucket_offsets=offs[:3],iters=1)
elif which.startswith("G"):
    nb = int(which[1:]); n = int(offs[nb])
    out = eng.scan_update(x0,P0,Q,clk0,pts[:n],[0,n],times[:nb],scan_bucket_ptr=[0,nb],bucket_offsets=offs[:nb+1],iters=1)
elif which == "D":
    clk0["last_predict_time"]=times[0]; clk0["last_update_time"]=times[0]
    n = int(offs[1])
    out = eng.scan_update(x0,P0,Q,clk0,pts[:n],[0,n],times[:1],iters=1)
print(which, "ok", out["n_eff"], out["x"]["pos"], "bucket sizes", np.diff(offs)[:12])
