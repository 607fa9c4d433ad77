for i in 1 2 3; do timeout 900 python bench.py --no-reference --no-cfg3 --no-cfg4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"; done
python tools/step_timeline.py 2>&1 | tail -19
python tools/step_breakdown.py 2>&1 | grep graph
