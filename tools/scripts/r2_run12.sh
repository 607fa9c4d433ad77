set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi.py -x -q > gpurun_out/r2_t12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t12.log
tail -6 gpurun_out/r2_t12.log
timeout 300 python tools/prof_roi_align.py roi 30 cfg3 2>&1 | grep -E "plan|bwd|fwd\+bwd"
TLOD_SWEEP_CPL1=1 timeout 300 python tools/prof_roi_align.py roi 30 cfg3 2>&1 | grep -E "bwd \[rows|fwd\+bwd"
TLOD_DISABLE_SWEEP=1 timeout 300 python tools/prof_roi_align.py roi 30 cfg3 2>&1 | grep -E "plan|bwd \[rows|fwd\+bwd"
timeout 300 python tools/prof_roi_align.py roi 30 cfg2 2>&1 | grep -E "plan|bwd|fwd\+bwd"
TLOD_SWEEP_CPL1=1 timeout 300 python tools/prof_roi_align.py roi 30 cfg2 2>&1 | grep -E "bwd \[rows"
TLOD_DISABLE_SWEEP=1 timeout 300 python tools/prof_roi_align.py roi 30 cfg2 2>&1 | grep -E "bwd \[rows"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"roi_align_bwd_sweep|roi_align_plan" -s 3 -c 3 -f -o gpurun_out/r2_sweep python tools/prof_roi_align.py roi 2 cfg3 > gpurun_out/r2_ncu_sweep.log 2>&1; tail -2 gpurun_out/r2_ncu_sweep.log
