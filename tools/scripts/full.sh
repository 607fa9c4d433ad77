timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json; d=json.load(open("gpurun_out/bench.json"))
for k in ("value","ms_per_step","e2e","gpu_launches","cuda_graph","proposals_per_s","roofline","roofline_roi_align_bwd"): print(k, d.get(k))
print({k:(round(v["ms_per_launch"]*1e3,1), v["launches"]) for k,v in d["kernels"].items()})
print(d["roi_align_cfg3"]); print(d.get("cpu_baseline"))
PY
