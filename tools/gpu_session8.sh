#!/bin/bash
# validate the new bench.py: default run as the driver launches it, reference arm, a 2-rank torchrun on one GPU is not possible -> N=1 only
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
( time timeout -k 10 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench_err.log ) 2> gpurun_out/s8_time.txt
tail -3 gpurun_out/s8_bench_err.log; cat gpurun_out/s8_time.txt | tail -4
python - <<'PY'
import json
d=json.load(open("gpurun_out/s8_bench.json"))
print("value %.3e us/step %.2f frac %.4f" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"]))
t=d["roofline"]["throughput_mode"]; print("throughput_mode", {k:t[k] for k in ("value","frac","ms_per_step","points_per_scan","map","pose_vs_cpu")})
print("e2e", {k:d["e2e"][k] for k in ("value","us_per_step","stream_p50_ms")}, d["e2e"]["stream"])
print("cpu", d["cpu_baseline"], d["pose_vs_cpu"], d["clocks"], d["config"]["host_cpus"])
PY
( time timeout -k 10 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/s8_ref.json 2> gpurun_out/s8_ref_err.log ) 2> gpurun_out/s8_time_ref.txt
tail -3 gpurun_out/s8_ref_err.log; tail -4 gpurun_out/s8_time_ref.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/s8_ref.json"))
print("ref value %.3e ms/step %.1f" % (d["value"], d["ms_per_step"]), d["config"]["step_seconds"], d["cpu_baseline"])
PY
