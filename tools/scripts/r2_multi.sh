set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-reference > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$?"
tail -4 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","per_rank_ms_per_step","graph_check","n_gpus"): print(k, d.get(k))
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["per_rank_ms_per_step"], d["e2e"]["pcie"])
print("allreduce", d["with_grad_allreduce"])
print("cfg4", {k:d["cfg4"][k] for k in ("value","ms_per_step","per_rank_ms_per_step","with_grad_allreduce")})
print(d["scaling_limiter"])
PY
nvidia-smi topo -m | head -12
