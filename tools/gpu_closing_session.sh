#!/bin/bash
# closing session of round 2: tests, bench lines, ncu of the throughput kernels, launch list
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
bash tools/gpu_final_tests.sh
bash tools/gpu_final_bench.sh
B="--workload synth100k_b1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_stream2 -s 6 -c 1 -f -o gpurun_out/r2_stream2 python bench.py $B > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log | cut -c1-120
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_fallback -s 6 -c 1 -f -o gpurun_out/r2_fallback python bench.py $B > gpurun_out/ncu7.log 2>&1; tail -1 gpurun_out/ncu7.log | cut -c1-120
bash tools/gpu_launches_throughput.sh
