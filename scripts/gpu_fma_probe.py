import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesheditor_b200 import measure_fp32_fma_rate
for mode in range(8):
    print(mode, ["FFMA", "FFMA2", "FFMA2:FFMA 1:1", "FFMA2:FFMA 1:2", "FADD", "FFMA2 3 fresh pairs", "FFMA2 2 fresh pairs", "FFMA 3 fresh regs"][mode], "%.3e lane-ops/s" % measure_fp32_fma_rate(0, mode, 20))
