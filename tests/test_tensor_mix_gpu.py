"""The tcgen05 mix kernel (mesheditor_b200/csrc/tensor_mix.cu) against a float64 matrix product.

out[row][tile*N*256 + n*256 + r] = sum_k P_group[r, k] * W_tile,group[n, k] over the 4096 reduction elements of each of the
row's groups (r = frame inside the 256-frame time block, n = time block inside the tile),
operands split into FP16 hi + lo (22 significant bits; hi*hi + hi*lo + lo*hi in FP32): powers pre-split in the shared-memory
stage layout, states plain row-major FP32 (scaled per 512-element range and split inside the kernel). Tolerance: 4e-6 of the
row scale at the maximum over 65k outputs (the 3xTF32 / BF16 split this replaced needed 1.2e-5; a plain TF32 product would miss by 1e-3).
The states span twelve orders of magnitude between ranges, which the per-range scale has to absorb.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STAGES = 256  # per group
KC = 16
BLOCK = 256  # frames per time block


def pack(mat):
    """[256, 4096] -> [stages][4096 floats]: per stage the FP16 hi (8 KB) and lo (8 KB) images of tensor_mix.cuh."""
    rows = mat.shape[0]
    hi = mat.astype(np.float16)
    lo = (mat.astype(np.float32) - hi.astype(np.float32)).astype(np.float16)
    out = np.zeros((STAGES, 16384), np.uint8)
    r = np.arange(rows)[:, None]
    k = np.arange(KC)[None, :]
    at16 = (k // 8) * rows * 16 + (r // 8) * 128 + (r % 8) * 16 + (k % 8) * 2
    for s in range(STAGES):
        block = slice(s * KC, (s + 1) * KC)
        for base, values in ((0, hi[:, block]), (8192, lo[:, block])):
            b = np.ascontiguousarray(values).view(np.uint8).reshape(rows, KC, 2)
            for byte in range(2):
                out[s, base + at16 + byte] = b[:, :, byte]
    return out.view(np.float32)


@pytest.mark.parametrize("n_blocks,groups,per_row,tiles,ragged", [(128, 2, 1, 2, 0), (128, 1, 1, 2, 777), (128, 4, 2, 1, 5000)])
def test_tensor_mix_matches_float64_product(n_blocks, groups, per_row, tiles, ragged):
    from mesheditor_b200 import lib
    from mesheditor_b200._lib import check

    rng = np.random.default_rng(7)
    K = STAGES * KC
    P = rng.standard_normal((groups, BLOCK, K)).astype(np.float32) * np.exp(rng.uniform(-6, 0, (groups, 1, K))).astype(np.float32)
    W = rng.standard_normal((tiles, groups, n_blocks, K)).astype(np.float32)
    # every (tile, group, block, 512-element range) at its own magnitude, 1e-6 .. 1e6: what the state scales are for
    W *= np.repeat(10.0 ** rng.uniform(-6, 6, (tiles, groups, n_blocks, K // 512)), 512, axis=-1).astype(np.float32)
    powers = np.stack([pack(P[g]) for g in range(groups)])
    states = np.ascontiguousarray(W)  # [tiles][groups][blocks][4096], row-major FP32: the kernel splits them itself
    frames = tiles * n_blocks * BLOCK - ragged
    rows = groups // per_row
    out = np.full((rows, frames), np.nan, np.float32)
    ms = C.c_float(0)
    check(lib().me_debug_tensor_mix(0, powers.ctypes.data, states.ctypes.data, groups, per_row, tiles, n_blocks, frames, 1, out.ctypes.data, C.byref(ms)))
    for r in range(rows):
        want, scale = 0.0, 0.0
        for g in range(r * per_row, (r + 1) * per_row):
            want = want + np.concatenate([(W[t, g].astype(np.float64) @ P[g].astype(np.float64).T).reshape(-1) for t in range(tiles)])[:frames]
            # per output: the root sum of squares of its terms (what a rounding error of the operands is relative to)
            scale = scale + np.concatenate([np.sqrt((W[t, g].astype(np.float64) ** 2) @ (P[g].astype(np.float64).T ** 2)).reshape(-1) for t in range(tiles)])[:frames] ** 2
        scale = np.sqrt(scale)
        assert np.isfinite(out[r]).all()
        err = np.abs(out[r] - want) / scale
        assert err.max() <= 4e-6, f"row {r}: max error {err.max():.3e} of the terms' root sum of squares"
