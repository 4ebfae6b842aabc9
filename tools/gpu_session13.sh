#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout -k 5 90 python -m pytest "tests/test_gpu_map.py::test_update_map_streaming" -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s13_pytest.log
tail -25 gpurun_out/s13_pytest.log
