mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_proposals.py tests/test_class_nms.py tests/test_gpu_reference_cuda.py tests/test_dropin.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/prof_proposals.py 2 20 2>&1 | tail -12
