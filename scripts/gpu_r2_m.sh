#!/bin/bash
# Round 2, visit M: full capture (with source) of the new state walk kernel, one whole-span launch.
mkdir -p gpurun_out
ME_WALK_SUBWINDOW_TILES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ResonatorKernel --launch-skip 3 --launch-count 1 -o gpurun_out/walk2_full -f python bench.py --workload resonator --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_walk2.log 2>&1; tail -3 gpurun_out/ncu_walk2.log
