#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 10 300 python tools/trace_fused.py leg_fusion_b1 > gpurun_out/s4_trace.log 2>&1; tail -45 gpurun_out/s4_trace.log
