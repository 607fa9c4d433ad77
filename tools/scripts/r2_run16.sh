set -x
mkdir -p gpurun_out
timeout 300 python tools/prof_roi_pool.py 2>&1 | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"roi_pool_" -c 12 -f -o gpurun_out/r2_pool python tools/prof_roi_pool.py > gpurun_out/r2_ncu_pool.log 2>&1; tail -2 gpurun_out/r2_ncu_pool.log
