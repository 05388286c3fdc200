#!/bin/bash
mkdir -p gpurun_out
ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "ritz m = 3" | tail -3
timeout 600 python bench.py --workload solve --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['profile']
print(round(d['value'],4), [round(x,3) for x in d['seconds_each']], 'iterate', round(p['iterate'],4), 'op', round(p['op_solve'],4))"
python -m pytest tests/test_modal_solve_gpu.py -m gpu -q 2>&1 | tail -1
