"""The tcgen05 mix kernel (mesheditor_b200/csrc/tensor_mix.cu) against a float64 matrix product.

out[row][tile*N*256 + n*256 + r] = sum_k P_group[r, k] * W_tile,group[n, k] over the 4096 reduction elements of each of the
row's groups (r = frame inside the 256-frame time block, n = time block inside the tile),
operands split into a TF32 head and an FP32 tail (3xTF32): powers pre-split in the shared-memory stage layout, states plain
row-major FP32 (split inside the kernel). Tolerance: 1.2e-5 of the row scale at the 4.5-sigma maximum over 65k outputs (the BF16 cross products leave ~2e-6 rms; a plain TF32
product would miss it by 1e-3).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STAGES = 256  # per group
KC = 16
BLOCK = 256  # frames per time block


def split_tf32(x):
    bits = x.astype(np.float32).view(np.uint32)
    head = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    return head, (x.astype(np.float32) - head).astype(np.float32)


def bf16_bits(x):
    """float32 -> bfloat16 bits, round to nearest even (__floats2bfloat162_rn)."""
    bits = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return (((bits + 0x7FFF + ((bits >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def pack(mat):
    """[256, 4096] -> [stages][8192 floats]: per stage the TF32 head (16 KB), BF16 value (8 KB) and BF16 tail (8 KB) images
    of tensor_mix.cuh."""
    rows = mat.shape[0]
    head, tail = split_tf32(mat)
    out = np.zeros((STAGES, 32768), np.uint8)
    r = np.arange(rows)[:, None]
    k = np.arange(KC)[None, :]
    at32 = (k // 4) * rows * 16 + (r // 8) * 128 + (r % 8) * 16 + (k % 4) * 4
    at16 = (k // 8) * rows * 16 + (r // 8) * 128 + (r % 8) * 16 + (k % 8) * 2
    for s in range(STAGES):
        block = slice(s * KC, (s + 1) * KC)
        h = np.ascontiguousarray(head[:, block]).view(np.uint8).reshape(rows, KC, 4)
        for byte in range(4):
            out[s, at32 + byte] = h[:, :, byte]
        for base, values in ((16384, mat[:, block]), (24576, tail[:, block])):
            b = bf16_bits(values).view(np.uint8).reshape(rows, KC, 2)
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
            scale += (P[g].astype(np.float64) ** 2).sum(axis=1).max()
        scale = np.sqrt(scale)
        err = np.abs(out[r] - want).max()
        assert np.isfinite(out[r]).all()
        assert err <= 1.2e-5 * scale, f"row {r}: max error {err:.3e} vs scale {scale:.3e}"
