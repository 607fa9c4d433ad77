mkdir -p gpurun_out
for w in 7 10 14 16 20; do for c in 1 2 3; do TLOD_POOL_WARPS=$w TLOD_POOL_CHUNKS=$c timeout 300 python tools/prof_roi_pool.py 2>&1 | grep "planes /" | sed "s/^/warps=$w chunks=$c /"; done; done
