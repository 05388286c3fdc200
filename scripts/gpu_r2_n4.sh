#!/bin/bash
# Round 2, 4-GPU visit: the default bench line at N = 4.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3) > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err; echo "exit $?"; tail -3 gpurun_out/n4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n4_bench.json').read().strip().splitlines()[-1])
print('resonator', d['value'], d['ms_per_step'], d['run']['step_breakdown_ms_rank0'], 'e2e', d['e2e']['ms_per_step'])
print('parity', d.get('parity',{}).get('slice',{}).get('gpu_vs_exact'), d.get('parity',{}).get('full_config',{}).get('gpu_vs_reference'))
print('solve', d['solve']['value'], 'batch', d['batch']['value'], d['batch']['seconds_per_batch'], d['batch']['load_balance'])
PY
