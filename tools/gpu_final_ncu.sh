#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-throughput --no-stream"
timeout -k 10 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-throughput --no-stream > gpurun_out/r2_e2e_check.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_e2e_check.json")); print("e2e us/step", d["e2e"]["us_per_step"], d["e2e"]["host_phases_us"], "us/step", d["ms_per_step"]*1e3)
PY
timeout -k 10 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r2_launches_leg_fusion_b1.csv python bench.py --steps 40 --warmup 10 $B > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout -k 10 700 ncu --set full --clock-control none --import-source on -k regex:k_scan_fused -s 30 -c 2 -f -o gpurun_out/r2_fused python bench.py --steps 40 --warmup 10 $B > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout -k 10 700 ncu --set full --clock-control none --import-source on -k regex:k_residual_stream2 -s 6 -c 1 -f -o gpurun_out/r2_stream2 python bench.py --workload synth100k_b1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
timeout -k 10 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file gpurun_out/r2_launches_nclt_stream.csv python bench.py --workload nclt_stream --steps 8 --warmup 3 > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/ncu4.log
timeout -k 10 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 12 --csv --log-file gpurun_out/r2_launches_nclt_stream_inkernel.csv python bench.py --workload nclt_stream --steps 8 --warmup 3 --param fused_insert=1 > gpurun_out/ncu5.log 2>&1; tail -2 gpurun_out/ncu5.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches*.csv
