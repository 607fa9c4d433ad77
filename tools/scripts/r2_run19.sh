mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --no-reference --no-cfg3 --no-cfg4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prio   value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d.get('graph_check'))"
TLOD_BENCH_FLAT_PRIORITY=1 timeout 900 python bench.py --no-reference --no-cfg3 --no-cfg4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flat   value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d.get('graph_check'))"
done
