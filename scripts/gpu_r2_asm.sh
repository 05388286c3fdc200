#!/bin/bash
# Round 2: full capture of AssembleKernel on the 1M-tet mesh (profiles/r02c_assemble_before_order_full_raw.csv was taken with it before the count order).
mkdir -p gpurun_out
cat > /tmp/asm.py <<'PY'
import sys
sys.path.insert(0,'.')
from mesheditor_b200 import FemSystem, workloads as wl
from mesheditor_b200.modal import MATERIALS
points,tets=wl.kuhn_block(55,55,55,(0.3,0.3,0.3))
for _ in range(2):
    fem=FemSystem(points,tets,MATERIALS["Steel"],1); print(fem.info["assemble_kernel_ms"]); fem.close()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:AssembleKernel --launch-skip 1 --launch-count 1 -o gpurun_out/assemble_full -f python /tmp/asm.py > gpurun_out/ncu_asm.log 2>&1; tail -2 gpurun_out/ncu_asm.log
ncu -i gpurun_out/assemble_full.ncu-rep --page raw --csv > gpurun_out/assemble_full_raw.csv 2>/dev/null; wc -c gpurun_out/assemble_full_raw.csv
