python tools/prof_roi_align.py roi 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:roi_align_bwd -s 1 -c 1 -f -o gpurun_out/prof_bwd_rows python tools/prof_roi_align.py roi 2 > gpurun_out/ncu_roi.log 2>&1
tail -3 gpurun_out/prof_plain.log
