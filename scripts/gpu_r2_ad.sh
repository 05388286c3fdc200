#!/bin/bash
# Round 2, visit AD: ring version with the producer's ticket pipeline: A/B and timelines, macro blocks off and 8.
mkdir -p gpurun_out
timeout 600 python scripts/gpu_sweep_ab.py 55 1 8 1 8 2>&1 | tee gpurun_out/sweep_ab.txt
for g in 1 8; do
ME_SWEEP_TRACE=gpurun_out/sweep_trace_ring_macro$g.bin timeout 600 python scripts/gpu_sweep_ab.py 55 $g 2>&1 | tail -1
python scripts/sweep_trace.py gpurun_out/sweep_trace_ring_macro$g.bin > gpurun_out/sweep_trace_ring_macro$g.txt
gzip -f gpurun_out/sweep_trace_ring_macro$g.bin
done
