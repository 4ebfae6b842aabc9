timeout 200 python tools/trace_fused.py leg_fusion_b1 2>&1 | sed -n 1,5p\;8,9p
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_filter.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python bench.py --steps 2048 --warmup 256 --no-cpu-baseline --no-batched 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4e us/step %.2f kernel_us %.2f e2e_us %.2f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['avg_launch_us'], d['e2e']['us_per_step']))"
