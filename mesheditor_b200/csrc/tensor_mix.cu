// sm_100a tcgen05 kernel of the resonator bank's tensor-core form; see tensor_mix.cuh for the algebra.
//
// One CTA renders one time tile (BlocksPerTile x 256 frames) of a few chunk groups: a 128 x 256 accumulator (time blocks x
// frames of a block) in tensor memory, double-buffered, fed stage by stage (16 reduction elements = 8 modes) through a
// ring of shared-memory buffers.
//   warp 0 (one lane): producer. Per stage a bulk copy (power stage, already a shared-memory image) and a 3-D TMA tile
//                      copy (16 reduction elements x all time blocks of the row-major FP32 states, 64-byte swizzle)
//                      straight into the stage's head half: kind::tf32 reads the top 19 bits, so FP32 rows are the head.
//   warps 2-3:         splitters. BF16 copies of x and of the tail x - truncated(x) for the two cross products, which run
//                      as kind::f16 MMAs (they need ~9 bits of each factor); they and the power copy complete the
//                      stage's "full" mbarrier.
//   warp 1 (one lane): per stage two kind::f16 MMAs (value*tail, tail*value on BF16 copies, K = 16) and two kind::tf32
//                      MMAs (head*head, K = 8), M = 128 time blocks x N = 256 frames, then tcgen05.commit to the stage's "empty" mbarrier; owns the TMEM allocation.
//   warps 4-11:        epilogue, one warp per (TMEM lane quarter, column half). tcgen05.ld of the accumulator (lane =
//                      frame inside the half block, column = block) folded into FP32 registers, then coalesced stores
//                      of the partial mix row.
#include "tensor_mix.cuh"

#include "common.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

namespace me {
namespace {

constexpr uint32_t kThreads = 384;
constexpr uint32_t kPowerHalfBytes = kTmBlock * kTmKChunk * 4; // 16 KB
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
__device__ __forceinline__ float Tf32Truncated(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); } // what kind::tf32 reads of x

// Shared-memory matrix descriptor, K-major, 64-byte swizzle: rows of 64 bytes (16 TF32), 8-row groups 512 bytes apart.
__device__ __forceinline__ uint64_t SwizzledDescriptor(uint32_t smem_addr) {
    return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(4) << 61);
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): 8-row core matrices of 128
// contiguous bytes; `lbo` = byte step between the two 16-byte K halves of one MMA, `sbo` = byte step between 8-row groups.
__device__ __forceinline__ uint64_t MatrixDescriptor(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) | (uint64_t(1) << 46);
}
// Instruction descriptor of kind::tf32 (cute::UMMA::InstrDescriptor): FP32 accumulate, TF32 A and B, both K-major.
__host__ __device__ constexpr uint32_t InstructionDescriptor(uint32_t m, uint32_t n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24); }

__device__ __forceinline__ void MmaTf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f16 with BF16 operands (format 1), FP32 accumulate, both K-major: K = 16 per instruction.
__host__ __device__ constexpr uint32_t InstructionDescriptorBf16(uint32_t m, uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24); }
__device__ __forceinline__ void MmaBf16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint2 PackBf16x4(float a, float b, float c, float d) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    return {reinterpret_cast<const uint32_t &>(lo), reinterpret_cast<const uint32_t &>(hi)};
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
// kFoldStages stages (16 MMAs per accumulator, <= 1e-6) and folded into FP32 registers by the epilogue warps with
// round-to-nearest adds; two accumulator pairs alternate so the fold of one overlaps the MMAs into the other.
constexpr uint32_t kFoldStages = 4;
constexpr uint32_t kFoldStagesHost = kFoldStages;

template<uint32_t N, uint32_t Stages>
__global__ void __launch_bounds__(kThreads, 1) TensorMixKernel(const TensorMixPlan plan, const __grid_constant__ CUtensorMap states_map) {
    static_assert(N == 128, "the epilogue keeps one accumulator row of N columns in registers");
    constexpr uint32_t kStateHalfBytes = N * kTmKChunk * 4;
    constexpr uint32_t kPowerBytes = 2 * kPowerHalfBytes, kStateBytes = 2 * kStateHalfBytes;
    constexpr uint32_t kStageBytes = kPowerBytes + kStateBytes;
    constexpr uint32_t kRawBytes = N * kTmKChunk * 4; // a stage's FP32 state rows: they ARE the head operand (the tensor core ignores the low 13 mantissa bits)
    constexpr uint32_t kSteps = kTmKChunk / 8; // MMAs of K = 8 per stage and operand pair
    static_assert(kTmKChunk == 16, "a state row is one 64-byte swizzle atom");
    extern __shared__ __align__(1024) uint8_t stage_storage_raw[];
    // The swizzle is a function of the shared-memory address: the ring starts on a 1024-byte boundary.
    uint8_t *stage_storage = stage_storage_raw + ((1024u - (SmemAddr(stage_storage_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full_bar[Stages], raw_bar[Stages], empty_bar[Stages], accum_full[2], accum_empty[2];
    __shared__ uint32_t tmem_base_slot;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Tiles of one group are adjacent in launch order, so the CTAs streaming the same power stages run together and
    // share them through L2.
    const uint32_t tile = blockIdx.x % plan.Tiles, row_index = blockIdx.x / plan.Tiles;
    const uint32_t n_stages = plan.StagesPerRow, first_stage = row_index * plan.StagesPerRow; // in the group-major stage sequence

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < Stages; ++s) BarrierInit(&full_bar[s], 3), BarrierInit(&raw_bar[s], 1), BarrierInit(&empty_bar[s], 1);
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

    if (warp == 0) {
        if (lane == 0) {
            const uint8_t *powers = reinterpret_cast<const uint8_t *>(plan.Powers) + size_t(first_stage) * kPowerBytes;
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&states_map)) : "memory");
            for (uint32_t k = 0; k < n_stages; ++k) {
                const uint32_t s = k % Stages, round = k / Stages;
                if (round) BarrierWait(&empty_bar[s], (round - 1) & 1);
                uint8_t *stage = stage_storage + size_t(s) * kStageBytes;
                BarrierExpectTx(&raw_bar[s], kRawBytes);
                TensorCopy3(stage, &states_map, ((first_stage + k) % kTmStagesPerGroup) * kTmKChunk, 0, tile * plan.Groups + (first_stage + k) / kTmStagesPerGroup, &raw_bar[s]);
                BarrierExpectTx(&full_bar[s], kPowerBytes);
                BulkCopy(stage + kStateBytes, powers + size_t(k) * kPowerBytes, kPowerBytes, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // The states are the A operand (M = N time blocks = TMEM lanes), the powers the B operand (N = 256 frames =
            // accumulator columns): one MMA covers the whole time block, so neither operand is read twice per K step.
            static_assert(N == 128 && kTmBlock == 256, "M = 128 time blocks, N = 256 frames");
            constexpr uint32_t idesc = InstructionDescriptor(N, kTmBlock), idesc16 = InstructionDescriptorBf16(N, kTmBlock);
            // Power images: the 16-byte K pieces are 4096 bytes apart and the 8-row groups 128 bytes.
            constexpr uint32_t lbo_p = kTmBlock * 16, sbo = 128;
            for (uint32_t k = 0; k < n_stages; ++k) {
                const uint32_t s = k % Stages, round = k / Stages;
                const uint32_t fold = k / kFoldStages, buffer = fold & 1;
                const bool opens = k % kFoldStages == 0;
                if (opens && fold >= 2) {
                    BarrierWait(&accum_empty[buffer], ((fold >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                BarrierWait(&full_bar[s], round & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t stage = SmemAddr(stage_storage + size_t(s) * kStageBytes);
                const uint32_t tmem_d = tmem_base + buffer * 2 * N;
                // Stage layout: [FP32 state rows 8 KB, swizzled][BF16 states 4 KB][BF16 state tails 4 KB]
                //               [TF32 power heads 16 KB][BF16 powers 8 KB][BF16 power tails 8 KB]
                const uint32_t w_value16 = stage + kRawBytes, w_tail16 = w_value16 + kRawBytes / 2;
                const uint32_t p_head32 = stage + kStateBytes, p_value16 = p_head32 + kTmPowerHeadBytes, p_tail16 = p_value16 + kTmPowerBf16Bytes;
                // The two cross products first (small terms), one K = 16 BF16 MMA each: state x power tail, state tail x power.
                MmaBf16(tmem_d, MatrixDescriptor(w_value16, N * 16, sbo), MatrixDescriptor(p_tail16, lbo_p, sbo), idesc16, !opens);
                MmaBf16(tmem_d, MatrixDescriptor(w_tail16, N * 16, sbo), MatrixDescriptor(p_value16, lbo_p, sbo), idesc16, 1);
#pragma unroll
                for (uint32_t kk = 0; kk < kSteps; ++kk) // head x head in TF32; inside the swizzle atom a K step of 8 is 32 bytes along the row
                    MmaTf32(tmem_d, SwizzledDescriptor(stage + kk * 32), MatrixDescriptor(p_head32 + kk * 2 * lbo_p, lbo_p, sbo), idesc, 1);
                MmaCommit(&empty_bar[s]); // arrives when the MMAs above have read the stage
                if (k % kFoldStages == kFoldStages - 1) MmaCommit(&accum_full[buffer]);
            }
        }
    } else if (warp < 4) {
        // Splitters: the TMA copy lands the FP32 state rows at the front of the stage, already in the swizzled K-major layout.
        // kind::tf32 reads only the top 19 bits of each element, i.e. the head is the TRUNCATED value; the two cross products
        // take BF16 copies of x and of the tail x - trunc(x), written in the canonical no-swizzle layout (8-element pieces).
        // Thread wt owns rows wt and wt + 64; the piece order keeps a quarter-warp's loads in different banks.
        static_assert(N == 128, "two rows per splitter thread");
        const uint32_t wt = (warp - 2) * 32 + lane;
        for (uint32_t k = 0; k < n_stages; ++k) {
            const uint32_t s = k % Stages, round = k / Stages;
            BarrierWait(&raw_bar[s], round & 1); // (the producer refilled this slot only after the MMAs of its last use)
            uint8_t *stage = stage_storage + size_t(s) * kStageBytes;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t n = wt + 64 * r, q = (n >> 1) & 3;
                uint2 value[4], tail[4]; // logical 4-element pieces 0..3 of the row
#pragma unroll
                for (uint32_t piece = 0; piece < 4; ++piece) {
                    const float4 v = *reinterpret_cast<const float4 *>(stage + n * 64 + ((piece ^ q) << 4)); // 64-byte swizzle: logical piece ^ row bits
                    value[piece] = PackBf16x4(v.x, v.y, v.z, v.w);
                    tail[piece] = PackBf16x4(v.x - Tf32Truncated(v.x), v.y - Tf32Truncated(v.y), v.z - Tf32Truncated(v.z), v.w - Tf32Truncated(v.w));
                }
                const uint32_t at = (n >> 3) * 128 + (n & 7) * 16; // + 2048 for elements 8..15
                *reinterpret_cast<uint4 *>(stage + kRawBytes + at) = {value[0].x, value[0].y, value[1].x, value[1].y};
                *reinterpret_cast<uint4 *>(stage + kRawBytes + N * 16 + at) = {value[2].x, value[2].y, value[3].x, value[3].y};
                *reinterpret_cast<uint4 *>(stage + kRawBytes + kRawBytes / 2 + at) = {tail[0].x, tail[0].y, tail[1].x, tail[1].y};
                *reinterpret_cast<uint4 *>(stage + kRawBytes + kRawBytes / 2 + N * 16 + at) = {tail[2].x, tail[2].y, tail[3].x, tail[3].y};
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(SmemAddr(&full_bar[s])) : "memory");
        }
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
            BarrierWait(&accum_full[buffer], (fold >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (uint32_t c0 = 0; c0 < N; c0 += 16) {
                uint32_t v[16];
                TmemLoad16(tmem_base + ((quarter * 32) << 16) + (buffer * 2 + half) * N + c0, v);
#pragma unroll
                for (uint32_t c = 0; c < 16; ++c) acc[c0 + c] += __uint_as_float(v[c]);
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
    constexpr uint32_t bytes = Stages * (2 * kPowerHalfBytes + 2 * N * kTmKChunk * 4) + 1024; // stage ring + alignment slack
    // States[tile*group][block][4096] as a 3-D tensor, innermost first; one box = the raw rows of one stage.
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

void LaunchTensorMixKernel(const TensorMixPlan &plan, cudaStream_t stream) {
    if (plan.Groups == 0 || plan.Tiles == 0) return;
    if (plan.StagesPerRow == 0 || plan.StagesPerRow % kFoldStagesHost != 0 || (uint64_t(plan.Groups) * kTmStagesPerGroup) % plan.StagesPerRow != 0)
        Fail(ME_BAD_ARG, "tensor mix: stages per row must be a multiple of %u that divides the stage count", kFoldStagesHost);
    if (plan.BlocksPerTile == 128) Launch<128, 4>(plan, stream);
    else Fail(ME_BAD_ARG, "tensor mix: blocks per tile must be 128");
}

} // namespace me
