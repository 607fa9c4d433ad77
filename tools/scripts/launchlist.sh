python bench.py --steps 2 --warmup 3 --no-cfg3 --no-cpu-baseline --no-graph > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cfg3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; k=h.index('Kernel Name'); v=h.index('Metric Value')
data=[(r[k], float(r[v].replace(',',''))) for r in rows[hdr+1:] if len(r)==len(h)]
# find one step: use the last occurrence window between two 'proposal_sort_runs' TRAIN launches
idx=[i for i,(n,_) in enumerate(data) if 'proposal_sort_runs' in n]
print("launches captured", len(data), "sort_runs at", idx[-6:])
# step = from idx[-4] to idx[-2] (two proposal calls per step)
s,e=idx[-4],idx[-2]
agg=collections.OrderedDict()
for n,t in data[s:e]:
    n=n.split('(')[0][-60:]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(a[1] for a in agg.values())
print("one step: %d launches, %.1f us of kernel time"%(e-s, tot/1e3))
for n,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%6.1f us %5.1f%% x%d %s"%(a[1]/1e3, 100*a[1]/tot, a[0], n))
PY
