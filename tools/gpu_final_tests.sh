#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
( time timeout -k 10 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -8 gpurun_out/f_pytest.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
