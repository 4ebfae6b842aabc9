#!/bin/bash
# round-2 closing session: full GPU test suite + smoke, the committed bench lines, the in-kernel launch list
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
bash tools/gpu_final_tests.sh
bash tools/gpu_final_bench.sh
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_scan_fused -s 4 -c 8 --csv --log-file gpurun_out/r2_launches_nclt_stream_inkernel.csv python bench.py --workload nclt_stream --steps 8 --warmup 3 --param fused_insert=1 > gpurun_out/ncu5.log 2>&1; tail -1 gpurun_out/ncu5.log | cut -c1-200
