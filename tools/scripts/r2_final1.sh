set -x
mkdir -p gpurun_out
bash tools/scripts/r2_full.sh > gpurun_out/r2_full.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref_arm.json 2> gpurun_out/r2_ref_arm.err; echo "ref rc=$?"
bash tools/scripts/r2_launchlist.sh > gpurun_out/r2_launchlist.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nms_scan|nms_mask|proposal_" -c 16 -f -o gpurun_out/r2_prop python tools/prof_proposals.py 2 2 > gpurun_out/r2_ncu_prop.log 2>&1; tail -1 gpurun_out/r2_ncu_prop.log
