#!/bin/bash
# Round 2: concurrent solves with the yielding spin barrier: the batch workload, the solve, the concurrency parity test.
mkdir -p gpurun_out
nproc
timeout 600 python bench.py --workload batch --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch', round(d['value'],2), 'meshes/s', round(d['seconds_per_batch'],3), 's')"
ME_BATCH_WORKERS=1 timeout 600 python bench.py --workload batch --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch, one at a time', round(d['value'],2), 'meshes/s')"
timeout 600 python bench.py --workload solve --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['profile']
print('solve', round(d['value'],4), [round(x,3) for x in d['seconds_each']], 'analyse', round(p['analyse'],4), 'iterate', round(p['iterate'],4))"
python -m pytest tests/test_modal_solve_gpu.py -m gpu -q 2>&1 | tail -1
