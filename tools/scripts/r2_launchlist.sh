set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-reference --no-cfg3 --no-cfg4 --no-graph-check > gpurun_out/r2_b_plain.json 2> gpurun_out/r2_b_plain.err; echo "bench rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-reference --no-cfg3 --no-cfg4 --no-graph-check > gpurun_out/r2_ncu_bench.log 2>&1; tail -1 gpurun_out/r2_ncu_bench.log | cut -c1-160
