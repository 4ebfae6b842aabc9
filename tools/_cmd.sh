timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python bench.py --steps 1024 --warmup 128 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3e us/step %.2f' % (d['value'], d['ms_per_step']*1e3), d['e2e']['us_per_step'], d['batched'], d['pose_vs_cpu'])"
