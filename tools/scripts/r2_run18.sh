set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_targets_da.py tests/test_gpu_roi.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --no-reference --no-cfg3 --no-cfg4 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench6.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","graph_check"): print(k, d.get(k))
print("e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
PY
timeout 300 python tools/step_breakdown.py 2>&1 | tail -25
