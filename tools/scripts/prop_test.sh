timeout 600 python -m pytest tests/test_gpu_proposals.py tests/test_class_nms.py tests/test_gpu_reference_cuda.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/prof_proposals.py 2 20
