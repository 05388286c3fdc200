#!/bin/bash
# Round 2, visit AB: task timelines (ME_SWEEP_TRACE, timer reads pinned between the barriers) of one panel application on the 1M-tet
# factor, macro blocks off and 8 panels.
mkdir -p gpurun_out
for g in 1 8; do
  ME_SWEEP_TRACE=gpurun_out/sweep_trace_c3_macro$g.bin timeout 600 python scripts/gpu_sweep_ab.py 55 $g 2>&1 | tail -1
  python scripts/sweep_trace.py gpurun_out/sweep_trace_c3_macro$g.bin > gpurun_out/sweep_trace_c3_macro$g.txt; head -8 gpurun_out/sweep_trace_c3_macro$g.txt
  gzip -f gpurun_out/sweep_trace_c3_macro$g.bin
done
