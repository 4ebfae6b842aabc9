timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/launches_stream.csv python bench.py --workload nclt_stream > gpurun_out/b_stream_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_stream.csv')) if len(r)>10 and r[0].isdigit()]
t=collections.defaultdict(float); n=collections.Counter()
for r in rows:
    k=r[4].split('(')[0].replace('lk::<unnamed>::','')[:50]; t[k]+=float(r[-1])/1e3; n[k]+=1
tot=sum(t.values())
for k,v in sorted(t.items(), key=lambda kv:-kv[1])[:14]: print('%-50s n=%5d total %8.1f us avg %6.2f us share %.1f%%' % (k, n[k], v, v/n[k], 100*v/tot))
print('launches', len(rows), 'total us', tot)
PY
