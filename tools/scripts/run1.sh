timeout 300 python -m pytest tests/test_gpu_roi.py tests/test_gpu_reference_cuda.py -x -q -m gpu 2>&1 | tail -4
timeout 120 python tools/prof_roi_align.py roi 20 2>&1 | tail -1
timeout 120 python tools/prof_roi_cfg2.py 2>&1 | tail -3
