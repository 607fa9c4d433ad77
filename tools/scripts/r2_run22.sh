mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --no-reference --no-cfg3 --no-cfg4 --no-cpu-baseline 2>gpurun_out/b22.err > gpurun_out/b22.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b22.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d.get('graph_check'))
print(json.dumps(d.get('secondary'),indent=0)[:1800])
PY
