// Tensor-core form of the free-running resonator bank (tensor_mix.cu).
//
// Between excitations every mode is z[n] = c z[n-1], so the samples of one 256-frame time block are a linear map of the
// state at the block's start:  y[b*256 + j] = sum_modes Im(c^(j+1) w_b) = sum_modes Re(c^(j+1)) Im w_b + Im(c^(j+1)) Re w_b.
// Over many blocks that is a GEMM  Y[256 x blocks] = P[256 x 2*modes] * W[2*modes x blocks]  whose reduction runs over
// the modes of every object in a chunk group: the mode sum AND the object mix of RenderModal (ModalAudio.cpp:125-128,
// 553-555) happen in the tensor-core accumulator. P (powers of the coefficients) is fixed by the tuning; W (block-start
// states) costs 4 FMAs and 16 bytes of HBM per mode per 256 samples: the block length trades the size of P (L2-resident)
// against the HBM traffic of W, which is what bounds both kernels.
//
// FP32 accuracy comes from a two-term FP16 split of both operands, x = hi + lo with hi = fp16(x) and lo = fp16(x - hi): 22
// significant bits, and hi*hi + hi*lo + lo*hi accumulate in FP32 (fp16 x fp16 products are exact there; the dropped lo*lo term
// is 2^-22 relative). Three kind::f16 MMAs per 16 reduction elements - half the operand bytes and three quarters of the
// tensor-pipe time of the 3xTF32 split this replaced. FP16 has five
// exponent bits, so the states are SCALED: the walk kernel records, per time block and per warp of chunk-threads (512
// reduction elements = 32 stages), the power of two that brings the largest state of that range to [2^13, 2^14)
// (TmStateScale) and stores the states multiplied by it, split; the mix kernel divides it out again when it folds an accumulator
// chain (8 stages, never across a range) into its FP32 registers. Entries far below a range's maximum lose relative precision
// against themselves, never against the sum they enter. The powers are at most 1 in magnitude and are split as they are.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace me {

constexpr uint32_t kTmBlock = 256;     // frames per time block: two M = 128 tcgen05.mma halves
constexpr uint32_t kTmKChunk = 16;     // reduction elements per pipeline stage = 8 modes = one chunk
constexpr uint32_t kTmGroupChunks = 256; // chunk slots per reduction group (== kBlockThreads: one resonator CTA)
constexpr uint32_t kTmStagesPerGroup = kTmGroupChunks * 8 * 2 / kTmKChunk; // 256

// Power stages in HBM are exactly what a pipeline stage holds in shared memory (canonical K-major, no-swizzle UMMA
// layouts), so one plain bulk copy fills them. A stage is the [256 x 16] block of one chunk in two FP16 images:
//   hi (8 KB): element (row r, k) at byte (k/8)*4096 + (r/8)*128 + (r%8)*16 + (k%8)*2
//   lo (8 KB): same layout, fp16(value - hi)
// Reduction indices 4p .. 4p+3 of a chunk hold (Re c_x^(j+1), Re c_y^(j+1), Im c_x^(j+1), Im c_y^(j+1)) in P and
// (Im w_x, Im w_y, Re w_x, Re w_y) in W for the chunk's mode pair p = (x, y): the walk's packed register pairs.
constexpr uint32_t kTmPowerImageBytes = kTmBlock * kTmKChunk * 2;
// States are written by the walk kernel as row-major matrices, one row of the group's 4096 reduction elements per time
// block (16 KB; a warp of chunk-threads stores 2 KB contiguous per step):
//   States[tile][group][time block][4096]
// ... of FP16 pairs, the scaled states' hi = fp16(s x) and lo = fp16(s x - hi): a chunk's 16 words (64 bytes) of a row are
// [16 x hi][16 x lo]. A stage (16 reduction elements of all blocks, hi and lo) reaches shared memory by one 3-D TMA tile copy
// with the 64-byte swizzle; the two operands are the two 32-byte halves of its rows.
__host__ __device__ constexpr size_t TmPowerStageFloats() { return size_t(2) * kTmPowerImageBytes / 4; }
// Scales[tile][group][range of 32 stages (8)][time block]: see above. One float per (time block, walk warp).
constexpr uint32_t kTmScaleRanges = 8, kTmStagesPerRange = 32;
__host__ __device__ constexpr size_t TmScaleTileFloats(uint32_t blocks_per_tile) { return size_t(kTmScaleRanges) * blocks_per_tile; }
// The power of two that brings `largest` (the largest magnitude of a range, >= 0) into [2^13, 2^14); 1 for an empty range.
__host__ __device__ inline float TmStateScale(float largest) {
    union {
        float f;
        uint32_t u;
    } v;
    v.f = largest;
    uint32_t biased = (v.u >> 23) & 0xFFu;
    if (largest == 0.f || biased == 0xFFu) return 1.f; // (a NaN or infinity poisons its own rows only)
    biased = biased < 14u ? 14u : biased;
    v.u = (267u - biased) << 23;
    return v.f;
}
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
    const float *States;      // [Tiles][Groups][BlocksPerTile][256 chunks][FP16 hi x 16, FP16 lo x 16]
    const float *Scales;      // [Tiles][Groups][8][BlocksPerTile]
    float *Partial;           // [Groups * 256 / StagesPerRow][Frames] partial mixes
};

void LaunchTensorMixKernel(const TensorMixPlan &, cudaStream_t);
// The scale table of a state array, as the walk kernel writes it (for callers that bring their own states: the unit test).
void LaunchStateScaleKernel(const float *states, uint32_t tiles_times_groups, uint32_t blocks_per_tile, float *scales, cudaStream_t);
// ... and its FP16 hi / lo rows from FP32 rows [tile x group][block][4096] (same size in bytes).
void LaunchStateSplitKernel(const float *states, uint32_t tiles_times_groups, uint32_t blocks_per_tile, const float *scales, float *planes, cudaStream_t);

} // namespace me
