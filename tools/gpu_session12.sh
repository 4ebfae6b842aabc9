#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_filter.py tests/test_gpu_parity.py tests/test_facade_compiles.py tests/test_gpu_edges.py -m gpu -x -q > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s12_pytest.log
tail -12 gpurun_out/s12_pytest.log
for wl in nclt_stream leg_fusion_stream; do
timeout -k 10 300 python bench.py --workload $wl --steps 40 --warmup 5 > gpurun_out/s12_$wl.json 2> gpurun_out/s12_err.log
python - <<PY
import json
d=json.load(open("gpurun_out/s12_$wl.json"))
print("$wl p50 %.3f ms mean %.3f p95 %.3f" % (d["value"], d["ms_per_step"], d["p95_ms"]), d["config"]["n_eff_mean"])
PY
done
