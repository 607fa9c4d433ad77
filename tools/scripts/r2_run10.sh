set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/prof_proposals.py 2 30 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_prop_launches.csv python tools/prof_proposals.py 2 1 > gpurun_out/r2_prop_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_prop_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
out=[(r[ki][:44], float(r[vi])/1000.0) for r in rows[1:]]
# last TRAIN call = launches [-?]; print the final 12 (TEST last call) and the 6 before (TRAIN last)
for k,v in out[-36:]: print("%-46s %8.1f us"%(k,v))
PY
