// Tensor-core form of the free-running resonator bank (tensor_mix.cu).
//
// Between excitations every mode is z[n] = c z[n-1], so the samples of one 256-frame time block are a linear map of the
// state at the block's start:  y[b*256 + j] = sum_modes Im(c^(j+1) w_b) = sum_modes Re(c^(j+1)) Im w_b + Im(c^(j+1)) Re w_b.
// Over many blocks that is a GEMM  Y[256 x blocks] = P[256 x 2*modes] * W[2*modes x blocks]  whose reduction runs over
// the modes of every object in a chunk group: the mode sum AND the object mix of RenderModal (ModalAudio.cpp:125-128,
// 553-555) happen in the tensor-core accumulator. P (powers of the coefficients) is fixed by the tuning; W (block-start
// states) costs 4 FMAs and 16 bytes of HBM per mode per 256 samples: the block length trades the size of P (L2-resident)
// against the HBM traffic of W, which is what bounds both kernels.  FP32 accuracy comes from the 3xTF32 split: both operands are stored
// as a TF32 head and an FP32 tail (P once per tuning, W inside the mix kernel), and head*head + head*tail + tail*head
// accumulate in FP32 (the dropped tail*tail term is 2^-22 relative). The two cross products only need ~9 bits of each
// factor, so they run as kind::f16 MMAs on BF16 copies (half the operand bytes and twice the rate of TF32).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace me {

constexpr uint32_t kTmBlock = 256;     // frames per time block: two M = 128 tcgen05.mma halves
constexpr uint32_t kTmKChunk = 16;     // reduction elements per pipeline stage = 8 modes = one chunk
constexpr uint32_t kTmGroupChunks = 256; // chunk slots per reduction group (== kBlockThreads: one resonator CTA)
constexpr uint32_t kTmStagesPerGroup = kTmGroupChunks * 8 * 2 / kTmKChunk; // 256

// Power stages in HBM are exactly what a pipeline stage holds in shared memory (canonical K-major, no-swizzle UMMA
// layouts), so one plain bulk copy fills them. A stage is the [256 x 16] block of one chunk in three images:
//   TF32 head   (16 KB): element (row r, k) at byte (k/4)*4096 + (r/8)*128 + (r%8)*16 + (k%4)*4
//   BF16 value  ( 8 KB): element (row r, k) at byte (k/8)*4096 + (r/8)*128 + (r%8)*16 + (k%8)*2
//   BF16 tail   ( 8 KB): same layout, value - head
// Reduction index 2*m holds Re(c^(j+1)) in P and Im w in W; 2*m+1 holds Im(c^(j+1)) and Re w, m = mode inside the chunk.
constexpr uint32_t kTmPowerHeadBytes = kTmBlock * kTmKChunk * 4, kTmPowerBf16Bytes = kTmBlock * kTmKChunk * 2;
// States are written by the walk kernel as plain FP32 row-major matrices, one row of the group's 4096 reduction elements
// per time block, so a warp of chunk-threads stores 2 KB contiguous per step:
//   States[tile][group][time block][4096]
// A stage (16 reduction elements of all blocks) reaches shared memory by one 3-D TMA tile copy with the 64-byte swizzle the
// UMMA descriptor expects; the FP32 rows serve as the head operand as they are (kind::tf32 ignores the low 13 mantissa bits)
// and two splitter warps of the mix kernel write BF16 copies of x and of the tail x - truncated(x) next to them. Splitting in the kernel instead of
// in the walk halves the HBM traffic of the states (8 B per mode per block).
__host__ __device__ constexpr size_t TmPowerStageFloats() { return size_t(2) * kTmBlock * kTmKChunk; }
constexpr uint32_t kTmGroupK = kTmGroupChunks * 8 * 2; // 4096 reduction elements per group
__host__ __device__ constexpr size_t TmStateTileFloats(uint32_t blocks_per_tile) { return size_t(blocks_per_tile) * kTmGroupK; }

struct TensorMixPlan {
    uint32_t Groups;          // chunk groups (reduction ranges of 256 chunk slots)
    uint32_t StagesPerRow;    // consecutive stages (of the group-major sequence, 256 per group) reduced by one CTA into one
                              // partial row: a multiple of 4 that divides Groups * 256
    uint32_t Tiles;           // time tiles in the window
    uint32_t BlocksPerTile;   // 128 time blocks (the N extent)
    uint32_t Frames;          // valid frames of the window (the last tile may be ragged)
    const float *Powers;      // [Groups][256 stages] power stages
    const float *States;      // [Tiles][Groups][BlocksPerTile][4096]
    float *Partial;           // [Groups * 256 / StagesPerRow][Frames] partial mixes
};

void LaunchTensorMixKernel(const TensorMixPlan &, cudaStream_t);

} // namespace me
