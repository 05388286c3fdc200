// sm_100a tcgen05 kernel of the resonator bank's tensor-core form; see tensor_mix.cuh for the algebra.
//
// One CTA renders one time tile (BlocksPerTile x 256 frames) of a few chunk groups: a 128 x 256 accumulator (time blocks x
// frames of a block) in tensor memory, double-buffered, fed stage by stage (16 reduction elements = 8 modes) through a
// ring of shared-memory buffers.
//   warp 0 (one lane): producer. Per stage a bulk copy (power stage: the FP16 hi / lo images, already a shared-memory image)
//                      and a 3-D TMA tile copy of the states (64 bytes of every time block's row: a chunk's 16 FP16 hi
//                      values, then its 16 lo values; 64-byte swizzle): nothing in this kernel touches an operand between
//                      the copy and the tensor core.
//   warp 1 (one lane): per stage three kind::f16 MMAs (hi*lo, lo*hi, hi*hi; K = 16), M = 128 time blocks x N = 256 frames,
//                      then tcgen05.commit to the stage's "empty" mbarrier; owns the TMEM allocation.
//   warps 4-11:        epilogue, one warp per (TMEM lane quarter, column half). tcgen05.ld of the accumulator (lane = time
//                      block, column = frame inside the block) folded into FP32 registers with the range's scale divided
//                      out, then coalesced stores of the partial mix row.
// (Until the walk kernel wrote the split states itself, two splitter warps converted FP32 rows here: ~160 dependent
// instructions per stage on two warps, which - not shared-memory bandwidth - was what the kernel waited for. And a first
// layout with separate hi and lo planes, fetched by a 5-D copy in 16-byte rows, was bound by the TMA unit's request rate:
// 512 requests per stage instead of 128.)
#include "tensor_mix.cuh"

#ifndef ME_TM_FOLD_STAGES
#define ME_TM_FOLD_STAGES 8
#endif

#include "common.h"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cudaTypedefs.h>

namespace me {
namespace {

constexpr uint32_t kThreads = 384;
constexpr uint32_t kSpinLimit = 1u << 24;

__device__ __forceinline__ uint32_t SmemAddr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void BarrierInit(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void BarrierExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void BarrierWait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = SmemAddr(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > kSpinLimit) __trap(); // a lost arrival must surface as an error, never as a hung device
    }
}
__device__ __forceinline__ void BulkCopy(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(SmemAddr(dst)), "l"(src), "r"(bytes), "r"(SmemAddr(bar))
                 : "memory");
}

__device__ __forceinline__ void TensorCopy3(void *dst, const CUtensorMap *map, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(SmemAddr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0),
                 "r"(c1), "r"(c2), "r"(SmemAddr(bar))
                 : "memory");
}

// Shared-memory matrix descriptor, K-major, 64-byte swizzle: rows of 64 bytes, 8-row groups 512 bytes apart. A K = 16 FP16
// operand is 32 bytes of every row: the hi half at the row's start, the lo half 32 bytes in (inside the swizzle atom a
// step along K is a plain byte offset of the start address).
__device__ __forceinline__ uint64_t SwizzledDescriptor(uint32_t smem_addr) {
    return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(4) << 61);
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): 8-row core matrices of 128
// contiguous bytes; `lbo` = byte step between the two 16-byte K halves of one MMA, `sbo` = byte step between 8-row groups.
__device__ __forceinline__ uint64_t MatrixDescriptor(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) | (uint64_t(1) << 46);
}
// Instruction descriptor of kind::f16 (cute::UMMA::InstrDescriptor): FP32 accumulate, FP16 A and B (format 0), both K-major;
// K = 16 per instruction.
__host__ __device__ constexpr uint32_t InstructionDescriptorF16(uint32_t m, uint32_t n) { return (1u << 4) | (0u << 7) | (0u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24); }
// Executed by a whole (converged) warp, issued by the one lane elect.sync picks: the operands are then warp-uniform values the
// compiler keeps in uniform registers, instead of per-lane values it has to funnel there with a loop around every instruction.
__device__ __forceinline__ void MmaF16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void CommitElected(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
}
// Four scaled states -> their FP16 hi and lo parts (two packed pairs each).
__device__ __forceinline__ void SplitF16x4(float4 v, float scale, uint2 &hi, uint2 &lo) {
    const float a = v.x * scale, b = v.y * scale, c = v.z * scale, d = v.w * scale;
    const __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(a - f0.x, b - f0.y), l1 = __floats2half2_rn(c - f1.x, d - f1.y);
    hi = {reinterpret_cast<const uint32_t &>(h0), reinterpret_cast<const uint32_t &>(h1)};
    lo = {reinterpret_cast<const uint32_t &>(l0), reinterpret_cast<const uint32_t &>(l1)};
}
__device__ __forceinline__ void MmaCommit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemAddr(bar)) : "memory");
}
__device__ __forceinline__ void TmemLoad16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                   "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
[[maybe_unused]] __device__ __forceinline__ void TmemLoad32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, "
        "%28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Tensor-core accumulation rounds toward zero: n MMAs chained on one accumulator lose up to n * 2^-24 of its
// magnitude, always in the same direction (measured: 1e-4 after the 1536 MMAs of a group). So a chain is cut after
// kFoldStages stages (12 MMAs per accumulator, <= 1e-6) and folded into FP32 registers by the epilogue warps with
// round-to-nearest operations; two accumulators alternate so the fold of one overlaps the MMAs into the other. A chain
// is also the unit that shares one state scale (kTmStagesPerRange is a multiple of kFoldStages).
constexpr uint32_t kFoldStages = ME_TM_FOLD_STAGES;
constexpr uint32_t kFoldStagesHost = kFoldStages;
static_assert(kTmStagesPerRange % kFoldStages == 0, "an accumulator chain never crosses a scale range");

template<uint32_t N, uint32_t Stages>
__global__ void __launch_bounds__(kThreads, 1) TensorMixKernel(const TensorMixPlan plan, const __grid_constant__ CUtensorMap states_map) {
    static_assert(N == 128, "the epilogue keeps one accumulator row of N columns in registers");
    constexpr uint32_t kImageBytes = N * kTmKChunk * 2; // a stage's FP16 hi image of the states, and the lo image behind it
    constexpr uint32_t kStateBytes = 2 * kImageBytes, kPowerBytes = 2 * kTmPowerImageBytes;
    constexpr uint32_t kStageBytes = kStateBytes + kPowerBytes;
    static_assert(kTmKChunk == 16, "one K = 16 MMA per product and stage");
    extern __shared__ __align__(1024) uint8_t stage_storage_raw[];
    // The swizzle is a function of the shared-memory address: the ring starts on a 1024-byte boundary.
    uint8_t *stage_storage = stage_storage_raw + ((1024u - (SmemAddr(stage_storage_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full_bar[Stages], empty_bar[Stages], accum_full[2], accum_empty[2];
    __shared__ uint32_t tmem_base_slot;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Tiles of one group are adjacent in launch order, so the CTAs streaming the same power stages run together and
    // share them through L2.
    const uint32_t tile = blockIdx.x % plan.Tiles, row_index = blockIdx.x / plan.Tiles;
    const uint32_t n_stages = plan.StagesPerRow, first_stage = row_index * plan.StagesPerRow; // in the group-major stage sequence

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < Stages; ++s) BarrierInit(&full_bar[s], 1), BarrierInit(&empty_bar[s], 1);
        for (uint32_t b = 0; b < 2; ++b) BarrierInit(&accum_full[b], 1), BarrierInit(&accum_empty[b], 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(&tmem_base_slot)), "r"(4 * N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    // Producer and MMA issuer are one warp each, and a lone warp retires an instruction every ten cycles or so: at ~60
    // instructions per stage the issue loop itself, not the tensor pipe (384 cycles per stage) or the copies, paced the kernel
    // (58 % tensor-pipe activity in the ncu capture). So both loops are unrolled over the ring (a ring round is one accumulator
    // chain: Stages == kFoldStages), every address and descriptor of a slot is a constant offset from a base computed once, a
    // barrier that is already complete costs one try_wait, and the whole warp runs the loop with elect.sync around the
    // asynchronous instructions (operands stay in uniform registers).
    static_assert(Stages == kFoldStages, "one ring round = one accumulator chain");
    const uint32_t ring = SmemAddr(stage_storage), full0 = SmemAddr(&full_bar[0]), empty0 = SmemAddr(&empty_bar[0]);
    const auto wait = [](uint32_t bar, uint32_t parity) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        for (uint32_t spin = 0;; ++spin) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
            if (done) return;
            if (spin > kSpinLimit) __trap(); // a lost arrival must surface as an error, never as a hung device
        }
    };
    if (warp == 0) {
        {
            const uint8_t *powers = reinterpret_cast<const uint8_t *>(plan.Powers) + size_t(first_stage) * kPowerBytes;
            const uint64_t map = reinterpret_cast<uint64_t>(&states_map);
            asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
            for (uint32_t k = 0, round = 0; k < n_stages; k += Stages, ++round) {
                // a round stays inside one group (rows of stages are multiples of the ring)
                const uint32_t c0 = ((first_stage + k) % kTmStagesPerGroup) * kTmKChunk, c2 = tile * plan.Groups + (first_stage + k) / kTmStagesPerGroup;
#pragma unroll
                for (uint32_t s = 0; s < Stages; ++s) {
                    if (round) wait(empty0 + s * 8, (round - 1) & 1);
                    const uint32_t stage = ring + s * kStageBytes, bar = full0 + s * 8;
                    asm volatile(
                        "{\n\t.reg .pred q;\n\t"
                        "elect.sync _|q, 0xffffffff;\n\t"
                        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
                        "@q cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%2], [%3, {%4, %5, %6}], [%0];\n\t"
                        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%7], [%8], %9, [%0];\n\t}" ::"r"(bar),
                        "n"(kStageBytes), "r"(stage), "l"(map), "r"(c0 + s * kTmKChunk), "r"(0), "r"(c2), "r"(stage + kStateBytes), "l"(powers + size_t(k + s) * kPowerBytes), "n"(kPowerBytes)
                        : "memory");
                }
            }
        }
    } else if (warp == 1) {
        {
            // The states are the A operand (M = N time blocks = TMEM lanes), the powers the B operand (N = 256 frames =
            // accumulator columns): one MMA covers the whole time block, so neither operand is read twice per K step.
            static_assert(N == 128 && kTmBlock == 256, "M = 128 time blocks, N = 256 frames");
            constexpr uint32_t idesc = InstructionDescriptorF16(N, kTmBlock);
            // Power images: the 16-byte K pieces are 4096 bytes apart and the 8-row groups 128 bytes.
            constexpr uint32_t lbo_p = kTmBlock * 16, sbo = 128;
            // Slot 0's descriptors; slot s adds s * kStageBytes / 16 to the address field (the ring stays below 256 KB: no carry).
            // Stage layout: [state rows 8 KB: 64 bytes per time block = FP16 hi x 16, FP16 lo x 16, swizzled][FP16 power hi 8 KB][FP16 power lo 8 KB]
            const uint64_t w_hi0 = SwizzledDescriptor(ring), w_lo0 = SwizzledDescriptor(ring + 32);
            const uint64_t p_hi0 = MatrixDescriptor(ring + kStateBytes, lbo_p, sbo), p_lo0 = MatrixDescriptor(ring + kStateBytes + kTmPowerImageBytes, lbo_p, sbo);
            const uint32_t accum_full0 = SmemAddr(&accum_full[0]), accum_empty0 = SmemAddr(&accum_empty[0]);
            for (uint32_t k = 0, round = 0; k < n_stages; k += Stages, ++round) {
                const uint32_t buffer = round & 1, tmem_d = tmem_base + buffer * 2 * N;
                if (round >= 2) {
                    wait(accum_empty0 + buffer * 8, ((round >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
#pragma unroll
                for (uint32_t s = 0; s < Stages; ++s) {
                    wait(full0 + s * 8, round & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    constexpr uint64_t step = kStageBytes >> 4;
                    // The two cross products first (small terms), then hi x hi: one K = 16 MMA each.
                    MmaF16(tmem_d, w_hi0 + s * step, p_lo0 + s * step, idesc, s != 0);
                    MmaF16(tmem_d, w_lo0 + s * step, p_hi0 + s * step, idesc, 1);
                    MmaF16(tmem_d, w_hi0 + s * step, p_hi0 + s * step, idesc, 1);
                    CommitElected(empty0 + s * 8); // arrives when the MMAs above have read the stage
                }
                CommitElected(accum_full0 + buffer * 8);
            }
        }
    } else if (warp < 4) {
        // (warps 2 and 3 keep the epilogue warps on the TMEM lane quarters their indices select)
    } else {
        // TMEM lanes (time blocks) are reachable from the warp whose index mod 4 matches the lane quarter; warps 4-7 take
        // accumulator columns (frames of the block) 0-127, warps 8-11 columns 128-255.
        const uint32_t quarter = warp & 3, half = (warp - 4) >> 2;
        const uint32_t block = quarter * 32 + lane; // time block inside the tile
        float acc[N];
#pragma unroll
        for (uint32_t c = 0; c < N; ++c) acc[c] = 0.f;
#pragma unroll 1
        for (uint32_t fold = 0; fold < n_stages / kFoldStages; ++fold) {
            const uint32_t buffer = fold & 1;
            // the scale this chain's states were multiplied by (a power of two: its reciprocal is exact)
            const uint32_t global_stage = first_stage + fold * kFoldStages;
            const float unscale = __frcp_rn(__ldg(plan.Scales + (size_t(tile) * plan.Groups + global_stage / kTmStagesPerGroup) * TmScaleTileFloats(N) + (global_stage % kTmStagesPerGroup) / kTmStagesPerRange * N + block));
            BarrierWait(&accum_full[buffer], (fold >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (uint32_t c0 = 0; c0 < N; c0 += 16) {
                uint32_t v[16];
                TmemLoad16(tmem_base + ((quarter * 32) << 16) + (buffer * 2 + half) * N + c0, v);
#pragma unroll
                for (uint32_t c = 0; c < 16; ++c) acc[c0 + c] = fmaf(__uint_as_float(v[c]), unscale, acc[c0 + c]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(SmemAddr(&accum_empty[buffer])) : "memory");
        }
        // This thread holds 128 consecutive frames of its time block: 16-byte stores (the row is 16-byte aligned: every
        // offset below is a multiple of four frames), scalar ones across the ragged end of the window.
        float *out = plan.Partial + size_t(row_index) * plan.Frames;
        const uint32_t frame0 = tile * N * kTmBlock + block * kTmBlock + half * 128;
        const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll
        for (uint32_t c = 0; c < N; c += 4) {
            const uint32_t frame = frame0 + c;
            if (aligned && frame + 3 < plan.Frames) {
                *reinterpret_cast<float4 *>(out + frame) = {acc[c], acc[c + 1], acc[c + 2], acc[c + 3]};
            } else {
#pragma unroll
                for (uint32_t i = 0; i < 4; ++i)
                    if (frame + i < plan.Frames) out[frame + i] = acc[c + i];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(4 * N) : "memory");
}

// One warp per (tile x group, time block, range): the scale of 512 consecutive reduction elements of one state row.
__global__ void StateScaleKernel(const float *__restrict__ states, uint32_t blocks_per_tile, float *__restrict__ scales) {
    const uint32_t lane = threadIdx.x & 31, range = threadIdx.x >> 5, block = blockIdx.x, tg = blockIdx.y;
    const float4 *row = reinterpret_cast<const float4 *>(states + (size_t(tg) * blocks_per_tile + block) * kTmGroupK + range * kTmStagesPerRange * kTmKChunk);
    float largest = 0.f;
    for (uint32_t i = lane; i < kTmStagesPerRange * kTmKChunk / 4; i += 32) {
        const float4 v = row[i];
        largest = fmaxf(largest, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (uint32_t d = 16; d; d >>= 1) largest = fmaxf(largest, __shfl_xor_sync(0xffffffffu, largest, d));
    if (lane == 0) scales[(size_t(tg) * kTmScaleRanges + range) * blocks_per_tile + block] = TmStateScale(largest);
}

// FP32 state rows -> the FP16 hi / lo rows the walk kernel writes, with the scales of StateScaleKernel (unit test only).
__global__ void StateSplitKernel(const float *__restrict__ states, uint32_t blocks_per_tile, const float *__restrict__ scales, uint16_t *__restrict__ planes) {
    const uint32_t block = blockIdx.x, tg = blockIdx.y;
    const size_t row = size_t(tg) * blocks_per_tile + block;
    for (uint32_t k = threadIdx.x * 4; k < kTmGroupK; k += blockDim.x * 4) {
        const float scale = scales[(size_t(tg) * kTmScaleRanges + k / (kTmStagesPerRange * kTmKChunk)) * blocks_per_tile + block];
        uint2 hi, lo;
        SplitF16x4(*reinterpret_cast<const float4 *>(states + row * kTmGroupK + k), scale, hi, lo);
        uint16_t *chunk = planes + row * 2 * kTmGroupK + size_t(k / kTmKChunk) * 2 * kTmKChunk; // [hi x 16][lo x 16]
        *reinterpret_cast<uint2 *>(chunk + k % kTmKChunk) = hi;
        *reinterpret_cast<uint2 *>(chunk + kTmKChunk + k % kTmKChunk) = lo;
    }
}

PFN_cuTensorMapEncodeTiled EncodeTiled() {
    static PFN_cuTensorMapEncodeTiled fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &status) != cudaSuccess || status != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }();
    if (!fn) Fail(ME_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from this driver");
    return fn;
}

template<uint32_t N, uint32_t Stages>
void Launch(const TensorMixPlan &plan, cudaStream_t stream) {
    constexpr uint32_t bytes = Stages * (2 * N * kTmKChunk * 2 + 2 * kTmPowerImageBytes) + 1024; // stage ring + alignment slack
    // States[tile*group][block][4096 words] as a 3-D tensor of 32-bit words, innermost first; one box = 16 words (one chunk's FP16
    // hi x 16 and lo x 16) of every time block.
    CUtensorMap map;
    const cuuint64_t dims[3] = {kTmGroupK, N, cuuint64_t(plan.Tiles) * plan.Groups};
    const cuuint64_t strides[2] = {cuuint64_t(kTmGroupK) * 4, cuuint64_t(N) * kTmGroupK * 4};
    const cuuint32_t box[3] = {kTmKChunk, N, 1}, unit[3] = {1, 1, 1};
    const CUresult r = EncodeTiled()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(plan.States), dims, strides, box, unit, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) Fail(ME_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d)", int(r));
    ME_CUDA(cudaFuncSetAttribute(TensorMixKernel<N, Stages>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes))); // per device, cheap
    TensorMixKernel<N, Stages><<<plan.Groups * kTmStagesPerGroup / plan.StagesPerRow * plan.Tiles, kThreads, bytes, stream>>>(plan, map);
    ME_CUDA(cudaGetLastError());
}

} // namespace

void LaunchStateScaleKernel(const float *states, uint32_t tiles_times_groups, uint32_t blocks_per_tile, float *scales, cudaStream_t stream) {
    if (tiles_times_groups == 0 || blocks_per_tile == 0) return;
    StateScaleKernel<<<dim3(blocks_per_tile, tiles_times_groups), kTmScaleRanges * 32, 0, stream>>>(states, blocks_per_tile, scales);
    ME_CUDA(cudaGetLastError());
}

void LaunchStateSplitKernel(const float *states, uint32_t tiles_times_groups, uint32_t blocks_per_tile, const float *scales, float *planes, cudaStream_t stream) {
    if (tiles_times_groups == 0 || blocks_per_tile == 0) return;
    StateSplitKernel<<<dim3(blocks_per_tile, tiles_times_groups), 256, 0, stream>>>(states, blocks_per_tile, scales, reinterpret_cast<uint16_t *>(planes));
    ME_CUDA(cudaGetLastError());
}

void LaunchTensorMixKernel(const TensorMixPlan &plan, cudaStream_t stream) {
    if (plan.Groups == 0 || plan.Tiles == 0) return;
    if (plan.StagesPerRow == 0 || plan.StagesPerRow % kFoldStagesHost != 0 || (uint64_t(plan.Groups) * kTmStagesPerGroup) % plan.StagesPerRow != 0)
        Fail(ME_BAD_ARG, "tensor mix: stages per row must be a multiple of %u that divides the stage count", kFoldStagesHost);
    if (plan.BlocksPerTile == 128) Launch<128, 8>(plan, stream);
    else Fail(ME_BAD_ARG, "tensor mix: blocks per tile must be 128");
}

} // namespace me
