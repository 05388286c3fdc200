#!/bin/bash
mkdir -p gpurun_out
python scripts/stage_bench.py 55 1 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('solve_ms','solve_GBs')})"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:SweepKernel -s 2 -c 2 -o gpurun_out/sweep_full python scripts/stage_bench.py 55 1 2 > gpurun_out/sweep_ncu.log 2>&1
tail -2 gpurun_out/sweep_ncu.log | cut -c1-200
