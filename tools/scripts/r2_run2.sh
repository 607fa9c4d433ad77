set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi.py -x -q > gpurun_out/r2_t_roi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_roi.log
tail -5 gpurun_out/r2_t_roi.log
timeout 300 python tools/prof_roi_align.py roi 30 cfg3 > gpurun_out/r2_time_cfg3.log 2>&1; tail -14 gpurun_out/r2_time_cfg3.log
TLOD_DISABLE_REGROW=1 timeout 300 python tools/prof_roi_align.py roi 30 cfg3 2>&1 | grep bwd
timeout 300 python tools/prof_roi_align.py roi 30 cfg2 > gpurun_out/r2_time_cfg2.log 2>&1; tail -14 gpurun_out/r2_time_cfg2.log
TLOD_DISABLE_REGROW=1 timeout 300 python tools/prof_roi_align.py roi 30 cfg2 2>&1 | grep bwd
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"roi_align_fwd8|roi_align_bwd_regrow|roi_align_plan" -s 6 -c 6 -f -o gpurun_out/r2_roi_b python tools/prof_roi_align.py roi 2 cfg3 > gpurun_out/r2_ncu_b.log 2>&1; tail -2 gpurun_out/r2_ncu_b.log
