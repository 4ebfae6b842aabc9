#!/bin/bash
# N-GPU torchrun diagnosis: per-rank logs, with and without the configs[3] shard
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
N=${1:-2}
free -g | head -2; nvidia-smi -L | head -8; ulimit -a | grep -i "mem\|lock" 
for tag in nothr full; do
  extra=""; [ $tag = nothr ] && extra="--no-throughput --no-stream"
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 --tee 3 --log-dir gpurun_out/tlogs_$tag bench.py --gpus $N --steps 20 --warmup 5 $extra > gpurun_out/s13_$tag.out 2> gpurun_out/s13_$tag.err
  echo "$tag rc=$?"; tail -c 600 gpurun_out/s13_$tag.out; echo; tail -5 gpurun_out/s13_$tag.err | cut -c1-300
  dmesg 2>/dev/null | tail -3
done
