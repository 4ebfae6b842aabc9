#!/bin/bash
# re-capture the throughput kernel after the shared-memory accumulators (ncu --set full on the configs[3] shard)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 10 700 ncu --set full --clock-control none --import-source on -k regex:k_residual_stream2 -s 6 -c 1 -f -o gpurun_out/r2_stream2 python bench.py --workload synth100k_b1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log | cut -c1-200
ls -la gpurun_out/r2_stream2.ncu-rep
