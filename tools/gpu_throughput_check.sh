#!/bin/bash
# throughput family: parity + the two throughput lines + the driver's default line
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
( timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_inputs.py tests/test_properties.py tests/test_reference_golden.py tests/test_gpu_edges.py -m gpu -x -q ) 2>&1 | tail -4
run() { name=$1; shift; timeout -k 10 600 python bench.py "$@" > gpurun_out/s14_$name.json 2> gpurun_out/s14_$name.err || echo "$name FAILED"; tail -c 300 gpurun_out/s14_$name.err; }
run diter_b128 --workload diter_b128 --steps 6 --warmup 3
run synth100k_b1024 --workload synth100k_b1024 --steps 6 --warmup 3
run default --gpus 1 --steps 20 --warmup 5
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/s14_*.json")):
    d = json.load(open(f)); r = d.get("roofline") or {}
    print(f.split("s14_")[1][:-5], "value %.4e ms/step %.4f frac %s" % (d["value"], d["ms_per_step"], r.get("frac")), d.get("clocks"), (r.get("throughput_mode") or {}).get("frac"))
PY
