python -m pytest tests/test_gpu_roi.py -x -q -m gpu 2>&1 | tail -3
python tools/prof_roi_align.py roi 20 2>&1 | tail -1
M=gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
for d in 0 1; do
TLOD_FWD_DEBUG=$d ncu --metrics $M --clock-control none -k regex:roi_align_fwd -s 1 -c 1 --csv python tools/prof_roi_align.py roi 2 2>&1 | grep -E "roi_align_fwd" | awk -F'","' -v d=$d '{print "dbg="d, $(NF-2), $NF}'
done
