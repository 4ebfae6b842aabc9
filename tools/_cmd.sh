timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json
d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['us_per_step'], d['batched']['roofline']['frac'], d['pose_vs_cpu'], d['clocks'])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-e2e --no-batched > gpurun_out/b_under_ncu.log 2>&1; tail -3 gpurun_out/launches.csv | cut -c1-300
