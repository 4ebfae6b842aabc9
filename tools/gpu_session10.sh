#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_inputs.py -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s10_pytest.log
tail -4 gpurun_out/s10_pytest.log
for wlk in diter_b128 synth100k_b1024; do
timeout -k 10 600 python bench.py --workload $wlk --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s10_$wlk.json 2> gpurun_out/s10_err.log
python - <<PY
import json
d=json.load(open("gpurun_out/s10_$wlk.json"))
print("$wlk", "value %.4e ms/step %.3f frac %.4f launch_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))
PY
done
