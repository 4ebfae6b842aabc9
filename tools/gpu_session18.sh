#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
bash tools/gpu_session14.sh
bash tools/gpu_session17.sh
