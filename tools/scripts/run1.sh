python -m pytest tests/test_gpu_roi.py tests/test_gpu_reference_cuda.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/t_roi.log
cat gpurun_out/t_roi.log
python tools/prof_roi_align.py roi 20 2>&1 | tail -3 | tee gpurun_out/prof_roi_time.log
