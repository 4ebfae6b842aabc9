#!/bin/bash
# merged queue-drain + predict launch of the streaming path: parity + latency
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
( timeout -k 10 600 python -m pytest tests/test_gpu_filter.py tests/test_gpu_map.py tests/test_reference_golden.py -m gpu -x -q ) 2>&1 | tail -4
for w in nclt_stream leg_fusion_stream; do
  timeout -k 10 300 python bench.py --workload $w --steps 100 --warmup 5 > gpurun_out/s16_$w.json 2> gpurun_out/s16_$w.err || echo "$w FAILED"
  python -c "import json; d=json.load(open('gpurun_out/s16_$w.json')); print('$w', d['value'], d['p95_ms'], d['config'].get('launches_per_scan'))"
done
