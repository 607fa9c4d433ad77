# round-1 evidence: (1) launch list of the bench command, (2) --set full of the RoIAlign kernels at cfg3,
# (3) --set full of the dominant kernel inside the bench workload (cfg2)
set -e
python bench.py --steps 2 --warmup 3 --no-cfg3 --no-cpu-baseline --no-graph > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cfg3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"roi_align_(fwd_planes|bwd_rows)" -s 12 -c 4 -f -o gpurun_out/r01_bench_bwd python bench.py --steps 2 --warmup 3 --no-cfg3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench2.log 2>&1
python tools/prof_roi_align.py roi 2 > gpurun_out/prof_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"roi_align_(fwd_planes|bwd_rows)" -s 2 -c 2 -f -o gpurun_out/r01_cfg3_roi python tools/prof_roi_align.py roi 2 > gpurun_out/ncu_roi.log 2>&1
echo done
