#!/bin/bash
mkdir -p gpurun_out
python scripts/stage_bench.py 55 1 3 > gpurun_out/stage_c3.json 2> gpurun_out/stage_c3.err; cat gpurun_out/stage_c3.json; tail -3 gpurun_out/stage_c3.err
python scripts/stage_bench.py 24 2 3 > gpurun_out/stage_p2.json 2> gpurun_out/stage_p2.err; cat gpurun_out/stage_p2.json; tail -3 gpurun_out/stage_p2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stage_c3.csv python scripts/stage_bench.py 55 1 1 > gpurun_out/stage_ncu.log 2>&1
tail -2 gpurun_out/stage_ncu.log
