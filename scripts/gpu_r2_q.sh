#!/bin/bash
# Round 2, visit Q (2 GPUs): the two-device tests, then the default bench line at N = 2 (resonator sharded + NCCL, solve replicas, batch dealt over 2 ranks).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
(time timeout 900 python -m pytest tests/test_c5_parity_gpu.py tests/test_resonator_tensor_gpu.py -m gpu -q -x) > gpurun_out/pytest_n2.log 2>&1; tail -5 gpurun_out/pytest_n2.log
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('resonator', d['value'], d['ms_per_step'], d['run']['step_breakdown_ms_rank0'], 'e2e', d['e2e']['ms_per_step'])
print('parity', d.get('parity',{}).get('slice'), d.get('parity',{}).get('full_config'))
print('solve', d['solve']['value'], 'batch', d['batch']['value'], d['batch']['seconds_per_batch'], d['batch']['load_balance'])
PY
