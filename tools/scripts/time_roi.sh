python tools/prof_roi_align.py roi 20 2>&1 | tail -3 | tee gpurun_out/prof_roi_time.log
