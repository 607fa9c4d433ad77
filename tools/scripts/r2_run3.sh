set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi.py tests/test_gpu_reference_cuda.py -x -q > gpurun_out/r2_t_roi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_roi.log
tail -8 gpurun_out/r2_t_roi.log
timeout 300 python tools/prof_roi_pool.py 2>&1 | tail -3
TLOD_DISABLE_POOL_PLANES=1 timeout 300 python tools/prof_roi_pool.py 2>&1 | tail -3
timeout 300 python tools/prof_roi_align.py roi 30 cfg3 2>&1 | grep -E "plan|fwd\+bwd"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"roi_pool_" -s 4 -c 2 -f -o gpurun_out/r2_pool python tools/prof_roi_pool.py > gpurun_out/r2_ncu_pool.log 2>&1; tail -2 gpurun_out/r2_ncu_pool.log
