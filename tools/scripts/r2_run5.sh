set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_roi_crop.py -x -q -m gpu > gpurun_out/r2_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t5.log
tail -12 gpurun_out/r2_t5.log
timeout 300 python tools/prof_roi_crop.py 5 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"roi_crop_pool_fwd|roi_crop_pool_bwd" -s 2 -c 2 -f -o gpurun_out/r2_crop python tools/prof_roi_crop.py 1 > gpurun_out/r2_ncu_crop.log 2>&1; tail -2 gpurun_out/r2_ncu_crop.log
