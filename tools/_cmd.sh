timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 400 python bench.py --workload diter_b128 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_diter_b128.json 2>gpurun_out/err_diter.txt; python -c "
import json
d=json.load(open('gpurun_out/bench_diter_b128.json')); print('diter_b128 %.3e frac %.3f launch_us %.1f' % (d['value'], d['roofline']['frac'], d['roofline']['avg_launch_us']), d['pose_vs_cpu'], d['clocks'])"
tail -2 gpurun_out/err_diter.txt
