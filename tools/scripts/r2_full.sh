set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_full_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_full_tests.log
tail -6 gpurun_out/r2_full_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_full.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","graph_check","gpu_launches","cuda_graph","clocks"): print(k, d.get(k))
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("roofline", {k:d["roofline"][k] for k in ("kernel","frac","achieved","l2_resident")})
c3=d["roi_align"]["cfg3"]; print("cfg3", {k:c3[k] for k in ("frac_fwd_bwd","frac_fwd_bwd_graph","graph_plan_fwd_bwd_us","graph_plan_avg_fwd_bwd_us","frac_avg_fwd_bwd_fused_bytes_graph")}, c3["fwd"]["us"], c3["bwd"]["us"], c3["plan"]["us"])
print("cfg4", d["cfg4"]["value"], d["cfg4"]["ms_per_step"])
print("speedups", d.get("speedup_vs_reference_kernels"))
print("cpu", d.get("cpu_baseline"), d.get("cpu_as_shipped"))
PY
