set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_reference_cuda.py tests/test_gpu_targets_da.py tests/test_gpu_roi.py tests/test_roi_crop.py -x -q -m gpu > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log
tail -12 gpurun_out/r2_t4.log
timeout 900 python tools/reference_bench.py --cpu-steps 0 > gpurun_out/r2_refbench.json 2> gpurun_out/r2_refbench.err; cat gpurun_out/r2_refbench.json; tail -3 gpurun_out/r2_refbench.err
timeout 300 python tools/prof_roi_align.py roi 30 cfg3 2>&1 | grep -E "plan|fwd\+bwd"
