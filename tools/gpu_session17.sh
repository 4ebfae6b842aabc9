#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_gpu_map.py tests/test_gpu_filter.py -m gpu -x -q > gpurun_out/s17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s17_pytest.log
tail -4 gpurun_out/s17_pytest.log
timeout -k 5 120 python tools/trace_stream.py > gpurun_out/s17_trace.log 2>&1; tail -10 gpurun_out/s17_trace.log
for v in 1 0; do
timeout -k 5 120 python bench.py --workload nclt_stream --steps 40 --warmup 5 --param fused_insert=$v > gpurun_out/s17_nclt_$v.json 2> gpurun_out/s17_err.log
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s17_nclt_$v.json"))
    print("nclt_stream fused_insert=$v p50 %.3f ms mean %.3f p95 %.3f" % (d["value"], d["ms_per_step"], d["p95_ms"]), d["config"]["n_eff_mean"])
except Exception as e: print("failed", e)
PY
done
