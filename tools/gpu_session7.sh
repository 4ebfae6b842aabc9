#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_concurrency.py tests/test_gpu_parity.py tests/test_gpu_inputs.py -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
tail -5 gpurun_out/s7_pytest.log
timeout -k 10 300 python tools/trace_fused.py leg_fusion_b1 > gpurun_out/s7_trace.log 2>&1; tail -5 gpurun_out/s7_trace.log
timeout -k 10 300 python tools/trace_fused.py leg_fusion_b1 poll_ns=100 > gpurun_out/s7_trace_p100.log 2>&1; tail -5 gpurun_out/s7_trace_p100.log
for v in "poll_ns=0" "poll_ns=50" "poll_ns=100" "poll_ns=200" "poll_ns=400" "slim_p=0"; do
  timeout -k 10 300 python bench.py --steps 400 --warmup 40 --no-cpu-baseline --no-throughput --no-stream --param $v > gpurun_out/s7_bench_$(echo $v | tr -d ' =-').json 2> gpurun_out/s7_bench_err.log
  python - <<PY
import json
d=json.load(open("gpurun_out/s7_bench_$(echo $v | tr -d ' =-').json"))
print("$v", "us/step %.2f" % (d["ms_per_step"]*1e3), "frac %.4f" % d["roofline"]["frac"], "e2e us %.1f" % d["e2e"]["us_per_step"], d["e2e"].get("host_phases_us"))
PY
done
