#!/bin/bash
# Round 2, visit U: the other analysis configs (configs[0] P2, configs[1] torus in P1 and P2, a 196k-tet block in P2).
mkdir -p gpurun_out
timeout 900 python scripts/bench_solve.py c1 c2 c2 c2p2 c2p2 kuhn:32:2:100 kuhn:32:2:100 > gpurun_out/bench_c1_c2.jsonl 2> gpurun_out/bench_c1_c2.err; tail -3 gpurun_out/bench_c1_c2.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_c1_c2.jsonl'):
    d=json.loads(l); print(d['case'], d['tets'], 'P%d'%d['order'], d['modes'], 'status', d['status'], 'sec %.3f'%d['seconds'], 'dofs', d['dofs'], 'analyse %.3f factorize %.3f iterate %.3f op %.3f'%(d['analyse'],d['factorize'],d['iterate'],d['op_solve']), 'nnzL', d['factor_nonzeros'], 'ops', d['op_applications'])
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
