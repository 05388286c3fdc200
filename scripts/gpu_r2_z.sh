#!/bin/bash
# Round 2, visit Z: task timelines of one panel application (ME_SWEEP_TRACE) at 1M tets, macro blocks off and 4 panels.
mkdir -p gpurun_out
for g in 1 4; do
  ME_MACRO_PANELS=$g ME_SWEEP_TRACE=gpurun_out/sweep_trace_macro$g.bin timeout 600 python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_trace$g.json 2> gpurun_out/bench_trace$g.err; tail -2 gpurun_out/bench_trace$g.err
  python scripts/sweep_trace.py gpurun_out/sweep_trace_macro$g.bin > gpurun_out/sweep_trace_macro$g.txt; head -3 gpurun_out/sweep_trace_macro$g.txt
  gzip -f gpurun_out/sweep_trace_macro$g.bin
done
