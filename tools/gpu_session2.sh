#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_concurrency.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
tail -4 gpurun_out/s2_pytest.log
timeout -k 10 300 python tools/trace_fused.py leg_fusion_b1 > gpurun_out/s2_trace.log 2>&1; tail -24 gpurun_out/s2_trace.log
for v in "pdl=1" "pdl=0"; do
  timeout -k 10 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-batched --param $v > gpurun_out/s2_bench_$(echo $v | tr -d ' =-').json 2> gpurun_out/s2_bench_err.log
  python - <<PY
import json
d=json.load(open("gpurun_out/s2_bench_$(echo $v | tr -d ' =-').json"))
print("$v", "us/step", d["ms_per_step"]*1e3, "frac", d["roofline"]["frac"], "e2e us", d["e2e"]["us_per_step"], d["e2e"].get("host_phases_us"))
PY
done
