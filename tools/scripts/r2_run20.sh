mkdir -p gpurun_out
bash tools/scripts/r2_run21.sh 2>&1 | grep -v warning | tail -6
timeout 600 python -m pytest tests/test_gpu_targets_da.py tests/test_dropin.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py --no-reference --no-cfg3 --no-cfg4 --no-cpu-baseline 2>gpurun_out/b20.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d.get('graph_check'), d.get('gpu_launches'))"
tail -2 gpurun_out/b20.err
python tools/step_timeline.py 2>&1 | tail -20
