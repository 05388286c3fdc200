"""Timing probe of the tcgen05 mix kernel on random stage images (no parity check here; tests/test_tensor_mix_gpu.py has it)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesheditor_b200 import lib
from mesheditor_b200._lib import check

groups, tiles, n = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 37, 128
rng = np.random.default_rng(1)
powers = rng.standard_normal(groups * 256 * 2 * 256 * 16, dtype=np.float32)
states = rng.standard_normal(tiles * groups * n * 4096, dtype=np.float32)
frames = tiles * n * 256
gpr = int(sys.argv[3]) if len(sys.argv) > 3 else 1
out = np.zeros((groups // gpr, frames), np.float32)
ms = C.c_float(0)
check(lib().me_debug_tensor_mix(0, powers.ctypes.data, states.ctypes.data, groups, gpr, tiles, n, frames, 5, out.ctypes.data, C.byref(ms)))
mode_samples = groups * 2048 * frames
flops = 3 * 2 * 256 * n * 4096 * groups * tiles
print(f"groups {groups} tiles {tiles}: {ms.value:.3f} ms, {flops / ms.value / 1e9:.1f} TFLOP/s tf32 issued, {mode_samples / ms.value / 1e9:.2f} T mode-samples/s, "
      f"operand bytes {(powers.nbytes * tiles + states.nbytes) / 1e9:.2f} GB -> {(powers.nbytes * tiles + states.nbytes) / ms.value / 1e6:.0f} GB/s into the SMs")
