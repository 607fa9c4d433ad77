set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_proposals.py tests/test_class_nms.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_reference_cuda.py -x -q -m gpu -k "cfg5 or nms" 2>&1 | tail -3
timeout 300 python tools/prof_proposals.py 2 30 2>&1 | tail -2
timeout 300 python tools/prof_proposals.py 64 10 2>&1 | tail -2
timeout 300 python - <<'PY'
import sys
sys.path[:0]=['transfer-learning-library-for-object-detection_b200','.']
import torch, bench
d=bench.secondary_metrics(torch.device('cuda:0'))
for e in d['nms_sweep_test_6000_300']: print(e)
print(d['proposal_layer'])
PY
