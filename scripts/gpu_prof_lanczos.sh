#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_solve.py c3 c3 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('seconds','assemble','factorize','analyse','iterate','op_solve','op_applications','restarts')})"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Gemv|TallGemm|Axpby|SpmvMass|Permute' -c 6000 --csv --log-file gpurun_out/launches_lanczos.csv python scripts/bench_solve.py kuhn:40:1:200 > gpurun_out/lanczos_ncu.log 2>&1
tail -1 gpurun_out/lanczos_ncu.log | cut -c1-300
