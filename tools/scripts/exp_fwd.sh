timeout 300 python -m pytest tests/test_gpu_roi.py -x -q -m gpu 2>&1 | tail -3
for d in 0 2; do echo "TLOD_FWD_DEBUG=$d (0 = TMA bulk stores, 2 = 256-bit LSU stores)"; TLOD_FWD_DEBUG=$d timeout 120 python tools/prof_roi_align.py roi 20 2>&1 | tail -1; TLOD_FWD_DEBUG=$d timeout 120 python tools/prof_roi_cfg2.py 2>&1 | tail -3; done
