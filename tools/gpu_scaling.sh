#!/bin/bash
# N-GPU scaling check of bench.py exactly as the driver launches it
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
N=${1:-8}
( time timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s11_n$N.out 2> gpurun_out/s11_n${N}_err.log ) 2> gpurun_out/s11_time.txt
tail -3 gpurun_out/s11_n${N}_err.log | cut -c1-300; tail -4 gpurun_out/s11_time.txt
python - <<PY
import json
line = [l for l in open("gpurun_out/s11_n$N.out") if l.startswith("{")][-1]
d = json.loads(line)
json.dump(d, open("gpurun_out/s11_n$N.json", "w"), indent=1)
print("N=$N value %.3e us/step %.2f frac %.4f" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"]))
t=d["roofline"]["throughput_mode"]; print("throughput_mode", {k:t[k] for k in ("value","frac","ms_per_step","n_gpus")})
print("e2e", d["e2e"]["value"], d["e2e"]["us_per_step"], d["e2e"].get("stream_p50_ms"), d["config"]["host_cpus"], d["clocks"])
PY
