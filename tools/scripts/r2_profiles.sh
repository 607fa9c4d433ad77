# round-2 profile evidence: ncu --set full of the judged kernels (last repetition of tools/prof_all.py),
# the launch list of the default bench command, nothing timed here is a bench value
set -x
mkdir -p gpurun_out
python tools/prof_all.py 2 > gpurun_out/r2_prof_all_plain.log 2>&1; tail -1 gpurun_out/r2_prof_all_plain.log
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:"roi_align_fwd8|roi_align_bwd_rows|roi_align_plan|avgpool2x2|roi_crop_pool|roi_pool_|proposal_|nms_mask|nms_scan|da_image" \
  -f -o gpurun_out/r2_all python tools/prof_all.py 1 > gpurun_out/r2_ncu_all.log 2>&1; tail -2 gpurun_out/r2_ncu_all.log
python bench.py --steps 2 --warmup 3 --no-reference --no-cfg3 --no-cfg4 > gpurun_out/r2_b_plain.json 2> gpurun_out/r2_b_plain.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-reference --no-cfg3 --no-cfg4 > gpurun_out/r2_ncu_bench.log 2>&1; tail -1 gpurun_out/r2_ncu_bench.log | cut -c1-200
