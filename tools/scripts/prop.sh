python tools/prof_proposals.py 2 20 > gpurun_out/prop_plain.log 2>&1 && cat gpurun_out/prop_plain.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 28 -c 40 --csv --log-file gpurun_out/prop_launches.csv python tools/prof_proposals.py 2 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prop_launches.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; k=h.index('Kernel Name'); v=h.index('Metric Value')
for r in rows[hdr+1:]:
    if len(r)==len(h): print(r[k][:40], r[v])
PY
