#!/bin/bash
# compute-sanitizer over the panel-sweep parity tests (slab ring, macro blocks in one and both sweeps) and small modal solves:
# memcheck, then racecheck (shared-memory hazards of the ring / queue / operand) and synccheck.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_cholesky_gpu.py -m gpu -q -x -k "wide_panel or panel_solve" > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool exit $?"; grep -c "Race reported\|hazard\|Invalid\|Barrier error" gpurun_out/sanitize_$tool.log; tail -3 gpurun_out/sanitize_$tool.log
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_modal_solve_gpu.py -m gpu -q -x -k "golden_models and bar" > gpurun_out/sanitize_solve.log 2>&1; echo "memcheck solve exit $?"; tail -2 gpurun_out/sanitize_solve.log
