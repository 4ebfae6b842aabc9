#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 10 500 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_residual|k_scan_tail|k_predict' -s 9 -c 12 --csv --log-file gpurun_out/r2_launches_synth100k.csv python bench.py --workload synth100k_b1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu6.log 2>&1
grep -v "^==" gpurun_out/r2_launches_synth100k.csv | awk -F'","' 'NR>1{print substr($5,1,60), $NF}'
