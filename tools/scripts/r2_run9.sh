set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py tests/test_gpu_targets_da.py -x -q -m gpu > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t9.log
tail -12 gpurun_out/r2_t9.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-reference --no-cfg3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench4.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","graph_check"): print(k, d.get(k))
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
for k,v in list(d["kernels"].items()):
    if "gt_max" in k or "anchor" in k: print("%-34s %.1f us x %d  share %.3f"%(k, v["ms_per_launch"]*1e3, v["launches"], v["share"]))
PY
python tools/step_breakdown.py 2>&1 | tail -12
