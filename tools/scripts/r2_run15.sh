set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi.py tests/test_gpu_reference_cuda.py tests/test_dropin.py -x -q -m gpu 2>&1 | tail -8
timeout 300 python tools/prof_roi_pool.py 2>&1 | tail -8
