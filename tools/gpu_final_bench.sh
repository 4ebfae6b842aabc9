#!/bin/bash
# the bench lines committed under profiles/ (round 2): never under a profiler
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
run() { name=$1; shift; timeout -k 10 600 python bench.py "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err || echo "$name FAILED"; tail -c 300 gpurun_out/r2_bench_$name.err; }
run leg_fusion_b1_driver --gpus 1 --steps 20 --warmup 5
run leg_fusion_b1 --steps 4096 --warmup 512 --no-throughput --no-stream
run diter_b128 --workload diter_b128 --steps 6 --warmup 3
run synth100k_b1024 --workload synth100k_b1024 --steps 6 --warmup 3
run nclt_stream --workload nclt_stream --steps 100 --warmup 5
run leg_fusion_stream --workload leg_fusion_stream --steps 100 --warmup 5
run nclt_stream_inkernel --workload nclt_stream --steps 100 --warmup 5 --param fused_insert=1
run reference --impl reference --gpus 1 --steps 20 --warmup 5
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline") or {}
    print("%-34s value %.4e %s | ms/step %.4f | frac %s | e2e %s" % (f.split("r2_bench_")[1][:-5], d["value"], d["unit"], d["ms_per_step"], r.get("frac"), (d.get("e2e") or {}).get("us_per_step")))
PY
