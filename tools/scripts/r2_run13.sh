set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_proposals.py tests/test_class_nms.py tests/test_dropin.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_reference_cuda.py -x -q -m gpu -k "cfg5 or nms" 2>&1 | tail -3
timeout 300 python tools/prof_proposals.py 2 30 2>&1 | tail -2
timeout 300 python tools/prof_proposals.py 64 10 2>&1 | tail -2
timeout 300 python - <<'PY'
import sys
sys.path[:0]=['transfer-learning-library-for-object-detection_b200','.']
import torch, bench
d=bench.secondary_metrics(torch.device('cuda:0'))
for e in d['nms_sweep_test_6000_300']: print(e)
print({k:v['us_per_call'] for k,v in d['proposal_layer'].items()}, d['nms_12000'], d['nms_6000'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_prop_launches2.csv python tools/prof_proposals.py 2 1 > gpurun_out/r2_prop_ncu2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_prop_launches2.csv')) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
out=[(r[ki][:44], float(r[vi])/1000.0) for r in rows[1:]]
cur=[]; seqs=[]
for k,v in out:
    if 'sort_runs' in k:
        if cur: seqs.append(cur)
        cur=[]
    cur.append((k,v))
seqs.append(cur)
for s in seqs[2:4]+seqs[-1:]:
    print(len(s), [ (k.split('::')[1][:14] if '::' in k else k[:10], round(v,1)) for k,v in s])
PY
