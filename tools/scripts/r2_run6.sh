set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench1.json').read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","graph_check","gpu_launches","cuda_graph"): print(k, d.get(k))
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
    print("roofline", d["roofline"])
    for k,v in list(d["kernels"].items())[:14]: print("%-34s %.1f us x %d  share %.3f"%(k, v["ms_per_launch"]*1e3, v["launches"], v["share"]))
    c3=d["roi_align"]["cfg3"]; print("cfg3 frac_fwd_bwd", c3["frac_fwd_bwd"], "fwd", c3["fwd"]["us"], "bwd", c3["bwd"]["us"], "plan", c3["plan"]["us"], "avg_fwd", c3["avg_fwd"]["us"], "avg_bwd", c3["avg_bwd"]["us"])
    print("cfg4", d["cfg4"])
    print("speedups", d.get("speedup_vs_reference_kernels"))
    print("cpu_baseline", d.get("cpu_baseline"))
    print("secondary proposal", d["secondary"]["proposal_layer"])
    print("da cfg4", d["secondary"]["da_image_losses_cfg4"])
except Exception as e:
    print("parse failed", e)
PY
