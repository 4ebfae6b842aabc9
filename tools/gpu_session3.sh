#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,pstate --format=csv
for cfg in "20 5" "200 20" "4096 512" "20 5" "20000 4096"; do
  set -- $cfg
  timeout -k 10 300 python bench.py --steps $1 --warmup $2 --no-cpu-baseline --no-batched --no-e2e > gpurun_out/s3_b.json 2>gpurun_out/s3_err.log
  python - <<PY
import json
d=json.load(open("gpurun_out/s3_b.json"))
print("steps $1 warmup $2", "us/step %.2f" % (d["ms_per_step"]*1e3), d["clocks"])
PY
done
# clocks while a long latency-mode loop runs
nvidia-smi --query-gpu=clocks.sm,power.draw,pstate --format=csv,noheader -lms 20 > gpurun_out/s3_clk.txt &
SMI=$!
timeout -k 10 300 python bench.py --steps 60000 --warmup 100 --no-cpu-baseline --no-batched --no-e2e > gpurun_out/s3_long.json
kill $SMI
python - <<PY
import json
d=json.load(open("gpurun_out/s3_long.json"))
print("long", "us/step %.2f" % (d["ms_per_step"]*1e3), d["clocks"])
PY
sort gpurun_out/s3_clk.txt | uniq -c | sort -rn | head -12
