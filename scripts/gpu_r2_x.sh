#!/bin/bash
# Round 2, 8-GPU visit: the default bench line at N = 8.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3) > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "exit $?"; tail -3 gpurun_out/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print('resonator', d['value'], d['ms_per_step'], d['run']['step_breakdown_ms_rank0'], 'e2e', d['e2e']['ms_per_step'])
print('parity', d.get('parity',{}).get('slice'), d.get('parity',{}).get('full_config'))
print('solve', d['solve']['value'], [ (c['workload'][:12], round(c['value'],3)) for c in d['solve'].get('other_configs',[])], 'batch', d['batch']['value'], d['batch']['seconds_per_batch'], d['batch']['load_balance'])
PY
