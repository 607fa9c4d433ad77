set -x
mkdir -p gpurun_out
timeout 900 python bench.py --no-reference --no-cfg3 --no-cfg4 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench5.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","graph_check"): print(k, d.get(k))
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pcie"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
PY
