#!/bin/bash
# the two lines the driver runs, on the final tree
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 10 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/ref.err; tail -c 200 gpurun_out/ref.err
timeout -k 10 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_leg_fusion_b1_driver.json 2> gpurun_out/drv.err; tail -c 200 gpurun_out/drv.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_leg_fusion_b1_driver.json")); t=d["roofline"]["throughput_mode"]
print(d["value"], d["roofline"]["frac"], t["frac"], t.get("traffic"), d["e2e"]["us_per_step"], d["e2e"]["stream_p50_ms"], d["clocks"])
r=json.load(open("gpurun_out/r2_bench_reference.json")); print("reference", r["value"], r["ms_per_step"])
PY
