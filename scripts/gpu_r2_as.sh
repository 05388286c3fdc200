#!/bin/bash
# Round 2, visit AS: orthogonal basis of the Rayleigh-Ritz tridiagonalisation accumulated on the device: solve tests, A/B.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_modal_solve_gpu.py tests/test_pipeline_gpu.py tests/test_reference_suite_gpu.py -m gpu -q -x) > gpurun_out/pytest_solve.log 2>&1; head -3 gpurun_out/pytest_solve.log
for env in A=device ME_HOST_BASIS=1 A=device ME_HOST_BASIS=1; do
env $env timeout 600 python bench.py --workload solve --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['profile']
print('$env', round(d['value'],4), [round(x,3) for x in d['seconds_each']], 'iterate', round(p['iterate'],4), 'op', round(p['op_solve'],4), 'restarts', p['restarts'], 'apps', p['op_applications'], [(c['workload'][:12], round(c['value'],3)) for c in d.get('other_configs',[])])"
done
ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "ritz m = 3" | tail -2
