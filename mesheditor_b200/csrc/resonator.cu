// sm_100a kernels of the modal resonator bank.
//
// What is computed (reference: src/audio/ModalAudio.cpp:86-147, 486-590): every mode is a damped complex
// one-pole  z[n] = c z[n-1] + e[n],  e[n] = sum over the object's live impacts of force[n] * gain, and the
// object's sample is  sum_modes (pIm*Im z + pRe*Re z) * OutGain*ListenerGain.
//
// Two forms share this file's kernels: the FP32 sample loop below, and the tensor-core form, in which
// the same ResonatorKernel (template parameter Walk) advances a whole 256-frame time block per step and only writes
// the block-start states that tensor_mix.cu turns into samples (plus PowerTableKernel for its constant operand).
//
// How the sample loop is laid out for B200 (this form is bound by FP32 issue, not by HBM: SURVEY.md F9):
//   * ResonatorKernel: one thread owns one 8-mode chunk (the reference's Lanes) in registers for a whole time
//     segment, so a warp covers 256 modes and reads each per-mode column as two coalesced float4 per thread.
//   * The output rotation is folded into the state: w = z*(pIm + i pRe)*gain obeys the same recurrence and the
//     sample is just sum Im w.
//   * K samples are advanced at a time: the K-1 intermediate samples are read off the current state with the
//     precomputed powers c^j (Im(c^j w), two FMAs chained into a running sum, no separate add), and the state
//     jumps by c^K. That is (2K+3)/K FP32 lane-operations per mode-sample instead of the reference's 7
//     (2.75 at K = 4), all issued as Blackwell packed FFMA2/FMUL2/FADD2 (fma.rn.f32x2) on mode pairs.
//   * Modes are summed across the warp 32 samples at a time through a padded shared-memory transpose; each warp
//     writes one coalesced 128-byte segment of its own partial row, and MixKernel sums the rows in fixed order,
//     so the mix is deterministic (no float atomics) and no CTA barrier sits in the sample loop.
//   * Excitation is rare (a contact pulse lasts ~1/PulseStep samples) and the system is linear, so PulseKernel
//     renders each pulse's zero-state response on its own (samples + end-of-pulse state increment) and the main
//     kernel only adds that increment when the pulse ends. The main loop carries no excitation term at all.
//   * The same increments give the state at any later frame in closed form (FP64 powers of c), which is the
//     block-parallel scan along time: SegmentScanKernel seeds segments 1.. of a window so that different threads
//     render different time segments of the same chunk.
//   * The reference's audibility culling (LiveModeCount prefix, SilenceObject, :139-146) is reproduced at the
//     RenderModal block boundaries with one CTA-wide exchange per block; objects never straddle a CTA.
#include "resonator.cuh"
#include "tensor_mix.cuh"

#include "common.h"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace me {
namespace {

constexpr uint32_t kRowPad = 68; // floats per transposed row: 32 float2 partial sums + 4 keeps float4 alignment and spreads banks
constexpr float kSilentEnergy = 1e-12f; // ModalAudio.cpp:20

__device__ __forceinline__ uint64_t AsBits(float2 v) { return reinterpret_cast<uint64_t &>(v); }
__device__ __forceinline__ float2 Fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t &>(d)) : "l"(AsBits(a)), "l"(AsBits(b)), "l"(AsBits(c)));
    return d;
}
__device__ __forceinline__ float2 Mul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t &>(d)) : "l"(AsBits(a)), "l"(AsBits(b)));
    return d;
}
__device__ __forceinline__ float2 Add2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t &>(d)) : "l"(AsBits(a)), "l"(AsBits(b)));
    return d;
}

// Eight complex values held as four mode pairs.
struct Chunk {
    float2 Re[4], Im[4];
};

__device__ __forceinline__ void Load8(const float *p, float2 (&v)[4]) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = {a.x, a.y}, v[1] = {a.z, a.w}, v[2] = {b.x, b.y}, v[3] = {b.z, b.w};
}
__device__ __forceinline__ void Store8(float *p, const float2 (&v)[4]) {
    reinterpret_cast<float4 *>(p)[0] = {v[0].x, v[0].y, v[1].x, v[1].y};
    reinterpret_cast<float4 *>(p)[1] = {v[2].x, v[2].y, v[3].x, v[3].y};
}

// c^1 .. c^K of the chunk's eight coefficients.
template<int K>
struct Powers {
    float2 Re[K][4], Im[K][4];
    float2 NegImK[4];
};

template<int K>
__device__ __forceinline__ void MakePowers(const float2 (&cre)[4], const float2 (&cim)[4], Powers<K> &p) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // Powers are formed in FP64 from the float coefficient and rounded once.
        const double ax = cre[i].x, bx = cim[i].x, ay = cre[i].y, by = cim[i].y;
        double rx = ax, ix = bx, ry = ay, iy = by;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            p.Re[j][i] = {float(rx), float(ry)};
            p.Im[j][i] = {float(ix), float(iy)};
            const double nrx = rx * ax - ix * bx, nry = ry * ay - iy * by;
            ix = rx * bx + ix * ax, iy = ry * by + iy * ay;
            rx = nrx, ry = nry;
        }
        p.NegImK[i] = {-p.Im[K - 1][i].x, -p.Im[K - 1][i].y};
    }
}

// The state jumps by p = float(c^K), not by c^K: every jump is off by the same factor c^K / p = 1 + e, |e| <= 2^-24. Unlike the
// rounding of the arithmetic (random, a square-root walk) this error has a fixed sign per mode and adds up LINEARLY in the
// number of jumps: 3.6e-4 per second of audio per mode at K = 4 - invisible on a decaying strike, far outside the 1e-5 gate on
// a bank that rings for seconds (measured on BASELINE.json configs[4]: 2.9e-5 of peak inside one 15-block segment). So e is
// kept (FP64 quotient, rounded once) and every kDriftJumps jumps the state is multiplied by 1 + n e, which leaves n^2 e^2.
constexpr uint32_t kDriftJumps = 128;
template<int K>
__device__ __forceinline__ void MakeDrift(const float2 (&cre)[4], const float2 (&cim)[4], const Powers<K> &p, float *drift) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double a = h ? cre[i].y : cre[i].x, b = h ? cim[i].y : cim[i].x;
            double r = a, q = b;
#pragma unroll
            for (int j = 1; j < K; ++j) { // the recurrence of MakePowers
                const double nr = r * a - q * b;
                q = r * b + q * a;
                r = nr;
            }
            const double pr = h ? p.Re[K - 1][i].y : p.Re[K - 1][i].x, pi = h ? p.Im[K - 1][i].y : p.Im[K - 1][i].x;
            const double d = pr * pr + pi * pi;
            // e = (c^K - p) / p
            const double er = d > 0 ? ((r - pr) * pr + (q - pi) * pi) / d : 0.0, ei = d > 0 ? ((q - pi) * pr - (r - pr) * pi) / d : 0.0;
            drift[(4 * i + 2 * h) * kBlockThreads] = float(er);
            drift[(4 * i + 2 * h + 1) * kBlockThreads] = float(ei);
        }
    }
}
// w <- w (1 + n e): e as MakeDrift left it ([component][thread], this thread's column).
__device__ __forceinline__ void ApplyDrift(Chunk &w, const float *drift, uint32_t jumps) {
    const float n = float(jumps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float erx = n * drift[(4 * i) * kBlockThreads], eix = n * drift[(4 * i + 1) * kBlockThreads];
        const float ery = n * drift[(4 * i + 2) * kBlockThreads], eiy = n * drift[(4 * i + 3) * kBlockThreads];
        const float rx = w.Re[i].x, ix = w.Im[i].x, ry = w.Re[i].y, iy = w.Im[i].y;
        w.Re[i] = {rx + fmaf(rx, erx, -ix * eix), ry + fmaf(ry, ery, -iy * eiy)};
        w.Im[i] = {ix + fmaf(rx, eix, ix * erx), iy + fmaf(ry, eiy, iy * ery)};
    }
}

// K samples: y[j] = sum Im(c^(j+1) w) for j < K-1 straight off the state, then w <- c^K w and y[K-1] = sum Im w.
// Each sample is left as a float2 of two partial sums (the row sum adds them). Instructions that share a state
// operand are kept adjacent so the operand-reuse cache can serve it: FFMA2 with three fresh register pairs is
// bound by register-file bandwidth, not by the FMA pipe.
template<int K>
__device__ __forceinline__ void StepK(Chunk &w, const Powers<K> &p, float2 *column) {
    float2 acc[K > 1 ? K - 1 : 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j + 1 < K; ++j) acc[j] = i == 0 ? Mul2(p.Re[j][0], w.Im[0]) : Fma2(p.Re[j][i], w.Im[i], acc[j]);
#pragma unroll
        for (int j = 0; j + 1 < K; ++j) acc[j] = Fma2(p.Im[j][i], w.Re[i], acc[j]);
    }
#pragma unroll
    for (int j = 0; j + 1 < K; ++j) column[j * (kRowPad / 2)] = acc[j];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t2 = Mul2(w.Im[i], p.Re[K - 1][i]);
        const float2 t1 = Mul2(w.Re[i], p.Re[K - 1][i]);
        const float2 im = Fma2(w.Re[i], p.Im[K - 1][i], t2);
        w.Re[i] = Fma2(w.Im[i], p.NegImK[i], t1);
        w.Im[i] = im;
    }
    column[(K - 1) * (kRowPad / 2)] = Add2(Add2(w.Im[0], w.Im[1]), Add2(w.Im[2], w.Im[3]));
}

// One sample with c^1 (remainders of a tile that is not a multiple of K).
template<int K>
__device__ __forceinline__ void Step1(Chunk &w, const Powers<K> &p, float2 *column) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float rx = fmaf(-w.Im[i].x, p.Im[0][i].x, w.Re[i].x * p.Re[0][i].x);
        const float ry = fmaf(-w.Im[i].y, p.Im[0][i].y, w.Re[i].y * p.Re[0][i].y);
        w.Im[i].x = fmaf(w.Re[i].x, p.Im[0][i].x, w.Im[i].x * p.Re[0][i].x);
        w.Im[i].y = fmaf(w.Re[i].y, p.Im[0][i].y, w.Im[i].y * p.Re[0][i].y);
        w.Re[i].x = rx, w.Re[i].y = ry;
        sum += w.Im[i].x + w.Im[i].y;
    }
    column[0] = float2{sum, 0.f};
}

// Lane s sums sample s over the warp's 32 columns (64 partial sums).
__device__ __forceinline__ float SumRow(const float *rows, uint32_t lane) {
    const float4 *r = reinterpret_cast<const float4 *>(rows + lane * kRowPad);
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        const float4 x = r[j], y = r[j + 1];
        a += (x.x + x.y) + (x.z + x.w);
        b += (y.x + y.y) + (y.z + y.w);
    }
    return a + b;
}

// c^m for the chunk's eight modes from the FP64 polar form.
__device__ __forceinline__ void PolarPower(const BankView &b, uint32_t mode0, uint32_t m, float2 (&re)[4], float2 (&im)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float r[2], q[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double mag = exp(double(m) * b.LogRho[mode0 + 2 * i + h]); // ln 0 = -inf -> 0
            double sn, cs;
            sincos(double(m) * b.Theta[mode0 + 2 * i + h], &sn, &cs);
            r[h] = float(mag * cs), q[h] = float(mag * sn);
        }
        re[i] = {r[0], r[1]}, im[i] = {q[0], q[1]};
    }
}

// Walk == false: the sample loop (K samples per step). Walk == true: the producer of the tensor-core form, which
// advances one kTmBlock-frame time block per step and writes the block-start states as state stages (tensor_mix.cuh).
// Everything around the inner loop (RenderModal blocks, culling, increments, final state) is shared.
template<int K, int MinBlocks, bool Walk>
__global__ void __launch_bounds__(kBlockThreads, MinBlocks) ResonatorKernel(const BankView b, const RenderPlan plan) {
    extern __shared__ __align__(16) float transposed_storage[];
    float (*transposed)[kTile][kRowPad] = reinterpret_cast<float (*)[kTile][kRowPad]>(transposed_storage);
    __shared__ float cull_energy[kBlockThreads];
    __shared__ uint8_t cull_audible[kBlockThreads];

    if (plan.OnlyIf && (*reinterpret_cast<const volatile uint32_t *>(plan.Speculation) & plan.OnlyIf) == 0) return; // (grid-uniform: every CTA reads the same word, written by an earlier kernel)
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kBlockThreads + threadIdx.x;
    const uint32_t seg = blockIdx.y;
    const uint32_t mode0 = chunk * kLanes;
    const uint32_t object = chunk < b.NChunks ? b.ChunkObject[chunk] : kNoObject;
    const bool valid = object != kNoObject;

    const uint32_t seg_begin = seg * plan.SegmentFrames;                      // relative to the window
    const uint32_t seg_end = min(plan.Frames, seg_begin + plan.SegmentFrames);
    const bool last_segment = seg + 1 == plan.NSegments;

    Powers<K> p;
    Chunk w;
    float2 jump_re[4], jump_im[4];       // Walk: float(c^kTmBlock)
    float2 jump_lo_re[4], jump_lo_im[4]; // Walk: c^kTmBlock - float(c^kTmBlock): the jump is applied as hi + lo (see MakeDrift for why)
    float2 jump_nim[4], jump_lo_nim[4];  // Walk: the imaginary parts negated, so that every product of the jump is a packed FFMA2
#pragma unroll
    for (int i = 0; i < 4; ++i) jump_re[i] = jump_im[i] = jump_lo_re[i] = jump_lo_im[i] = jump_nim[i] = jump_lo_nim[i] = float2{0.f, 0.f};
    // Sample loop: this thread's drift terms e (16 floats), [component][thread], behind the transpose tiles.
    float *drift = transposed_storage + kWarpsPerBlock * kTile * kRowPad + threadIdx.x;
    uint32_t jumps = 0; // K-sample jumps since the drift was last taken out
    float gain = 1.f, out_scale = 0.f, energy_scale = 0.f;
    uint32_t my_chunk = 0, obj_chunks = 0, obj_first_local = 0, obj_tuned = 0;
    bool in_tuned = false, cull = false, live = true, ringing = true;
    uint32_t inj = 0, inj_hi = 0, exc = 0, exc_hi = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) w.Re[i] = w.Im[i] = float2{0.f, 0.f};
    if (valid) {
        float2 cre[4], cim[4];
        Load8(b.CoeffRe + mode0, cre), Load8(b.CoeffIm + mode0, cim);
        MakePowers<K>(cre, cim, p);
        if constexpr (Walk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { // c^kTmBlock by FP64 squarings of the float coefficient, rounded once
                double rx = cre[i].x, ix = cim[i].x, ry = cre[i].y, iy = cim[i].y;
#pragma unroll
                for (uint32_t q = 1; q < kTmBlock; q <<= 1) {
                    const double nrx = rx * rx - ix * ix, nry = ry * ry - iy * iy;
                    ix = 2.0 * rx * ix, iy = 2.0 * ry * iy;
                    rx = nrx, ry = nry;
                }
                jump_re[i] = {float(rx), float(ry)}, jump_im[i] = {float(ix), float(iy)};
                jump_lo_re[i] = {float(rx - double(jump_re[i].x)), float(ry - double(jump_re[i].y))};
                jump_lo_im[i] = {float(ix - double(jump_im[i].x)), float(iy - double(jump_im[i].y))};
                jump_nim[i] = {-jump_im[i].x, -jump_im[i].y}, jump_lo_nim[i] = {-jump_lo_im[i].x, -jump_lo_im[i].y};
            }
        } else {
            MakeDrift<K>(cre, cim, p, drift);
        }
        const uint32_t first = b.ObjFirstChunk[object];
        my_chunk = chunk - first;
        obj_chunks = b.ObjStride[object] / kLanes;
        obj_first_local = first - blockIdx.x * kBlockThreads; // only meaningful when the object fits this CTA
        obj_tuned = b.ObjTunedChunks[object];
        in_tuned = my_chunk < obj_tuned;
        cull = b.ObjCull[object] != 0;
        const float mix = b.ObjMixGain[object];
        const bool muted = mix == 0.f;
        gain = muted ? 1.f : mix;
        out_scale = muted ? 0.f : 1.f;
        energy_scale = b.ObjEnergyScale[object];
        if (seg == 0) {
            float2 zr[4], zi[4], qr[4], qi[4];
            Load8(b.PhaseIm + mode0, qr), Load8(b.PhaseRe + mode0, qi);
            Load8(b.StateRe + mode0, zr), Load8(b.StateIm + mode0, zi);
#pragma unroll
            for (int i = 0; i < 4; ++i) { // w = z * q * gain
                w.Re[i] = {(zr[i].x * qr[i].x - zi[i].x * qi[i].x) * gain, (zr[i].y * qr[i].y - zi[i].y * qi[i].y) * gain};
                w.Im[i] = {(zr[i].x * qi[i].x + zi[i].x * qr[i].x) * gain, (zr[i].y * qi[i].y + zi[i].y * qr[i].y) * gain};
            }
            live = b.ChunkLive[chunk] != 0;
            ringing = b.ObjRinging[object] != 0;
        } else {
            const size_t off = size_t(seg - 1) * b.NChunks * kLanes + mode0;
            Load8(plan.SegStateRe + off, w.Re), Load8(plan.SegStateIm + off, w.Im);
        }
        const uint32_t begin_abs = plan.FrameBegin + seg_begin;
        inj = plan.ObjInjectPtr[object], inj_hi = plan.ObjInjectPtr[object + 1];
        while (inj < inj_hi && plan.InjectFrame[inj] < begin_abs) ++inj;
        exc = plan.ObjExcitePtr[object], exc_hi = plan.ObjExcitePtr[object + 1];
    } else {
#pragma unroll
        for (int j = 0; j < K; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) p.Re[j][i] = p.Im[j][i] = float2{0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) p.NegImK[i] = float2{0.f, 0.f};
    }
    uint32_t inj_frame = inj < inj_hi ? plan.InjectFrame[inj] : 0xFFFFFFFFu;

    const auto inject = [&] {
        const uint32_t off = plan.InjectDelta[inj] + my_chunk * kLanes;
        float2 dr[4], di[4];
        Load8(plan.DeltaRe + off, dr), Load8(plan.DeltaIm + off, di);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            w.Re[i].x += dr[i].x, w.Re[i].y += dr[i].y;
            w.Im[i].x += di[i].x, w.Im[i].y += di[i].y;
        }
        ++inj;
        inj_frame = inj < inj_hi ? plan.InjectFrame[inj] : 0xFFFFFFFFu;
    };

    float *rows = &transposed[warp][0][0];
    float2 *column = reinterpret_cast<float2 *>(rows) + lane;
    float *partial = plan.Partial + size_t(blockIdx.x * kWarpsPerBlock + warp) * plan.Frames;

    // Walk: row-major states (tensor_mix.cuh). This chunk owns 16 consecutive reduction elements of every time-block row: 64 bytes,
    // their FP16 hi values then their lo values; a warp's 32 chunks own 2 KB. They pass through a swizzled shared-memory transpose
    // so that every store instruction covers 512 contiguous bytes. The row pointer is stepped, not recomputed: every step of the
    // walk starts on a time-block boundary and fills exactly one row (the last one may be ragged).
    uint4 *walk_at = nullptr; // this lane's 16 bytes of the first quarter of the warp's 2 KB of the row being written
    float *scale_at = nullptr; // this warp's entry of the scale table for the row being written (tensor_mix.cuh)
    uint32_t walk_row = 0;
    size_t walk_tile_skip = 0;
    if constexpr (Walk) {
        const uint32_t nb = plan.WalkBlocksPerTile;
        const uint32_t block_index = seg_begin / kTmBlock;
        const size_t tile_stride = size_t(gridDim.x) * TmStateTileFloats(nb) / 4;
        walk_row = block_index % nb;
        walk_at = reinterpret_cast<uint4 *>(plan.WalkStates + size_t(blockIdx.x) * TmStateTileFloats(nb)) + warp * 128 + lane + size_t(block_index / nb) * tile_stride + size_t(walk_row) * (kTmGroupK / 4);
        walk_tile_skip = tile_stride - size_t(nb) * (kTmGroupK / 4);
        scale_at = plan.WalkScales + (size_t(block_index / nb) * gridDim.x + blockIdx.x) * TmScaleTileFloats(nb) + size_t(warp) * nb + walk_row;
    }

    uint32_t pos = seg_begin;
    uint32_t next_block_abs = ((plan.FrameBegin + seg_begin) / plan.BlockFrames + 1) * plan.BlockFrames; // end of the RenderModal block holding `pos` (stepped, not divided for, per block)
    while (pos < seg_end) {
        // One RenderModal block (or what is left of it inside this segment).
        const uint32_t pos_abs = plan.FrameBegin + pos;
        const uint32_t block_end_abs = min(min(next_block_abs, plan.SpanFrames), plan.FrameBegin + seg_end);
        const uint32_t block_end = block_end_abs - plan.FrameBegin;
        if (block_end_abs == next_block_abs) next_block_abs += plan.BlockFrames;
        // Does the object hold a live impact in this block (:90 `impacts.empty()`)?
        while (exc < exc_hi && plan.ExciteEnd[exc] <= pos_abs) ++exc;
        const bool excited = exc < exc_hi && plan.ExciteBegin[exc] <= pos_abs;
        if (excited) ringing = true; // ActivateImpact :50
        while (inj_frame < pos_abs) { // increments that landed while this chunk was not rendered are all zero
            ++inj;
            inj_frame = inj < inj_hi ? plan.InjectFrame[inj] : 0xFFFFFFFFu;
        }
        const bool rendered = valid && in_tuned && (!cull || excited || (live && ringing));
        if (valid && in_tuned && !rendered) {
            // A chunk that holds state but sits out the block breaks the linear evolution the scan along time assumed.
            bool holds = false;
#pragma unroll
            for (int i = 0; i < 4; ++i) holds |= w.Re[i].x != 0.f || w.Re[i].y != 0.f || w.Im[i].x != 0.f || w.Im[i].y != 0.f;
            if (holds) {
                if (plan.Debug && atomicOr(plan.Speculation, seg == 0 ? 1u : 2u) == 0) printf("[me] frozen chunk holds state: object %u chunk %u seg %u frame %u live %d ringing %d excited %d\n", object, my_chunk, seg, pos_abs, int(live), int(ringing), int(excited));
                atomicOr(plan.Speculation, seg == 0 ? 1u : 2u);
            }
        }
        const bool warp_renders = __ballot_sync(0xFFFFFFFFu, rendered) != 0;

        if constexpr (Walk) {
            // With one segment the walk is sequential in time over the whole window and culling needs no speculation; small
            // banks (few chunk groups) are walked in several seeded segments like the sample loop.
            uint4 *exchange = reinterpret_cast<uint4 *>(transposed_storage) + warp * 128;
            const uint32_t put = lane * 4, put_swizzle = (lane >> 1) & 3;
            const bool audible = rendered && out_scale != 0.f;
            for (uint32_t t = pos; t < block_end;) {
                const uint32_t t_abs = plan.FrameBegin + t;
                const uint32_t step = min(kTmBlock, block_end - t);
                if (rendered) {
                    while (inj_frame == t_abs) inject();
                    // Increments land on time-block boundaries in a span planned for this form (DevImpact::RenderLen).
                    if (inj_frame - t_abs < step) atomicOr(plan.Speculation, 8u);
                }
                // The row as the mix kernel's tensor-core operand (tensor_mix.cuh): the warp's 512 entries scaled by the power of
                // two that brings the largest into FP16 range, split into FP16 hi + lo. Reduction elements 4p .. 4p+3 of the chunk
                // are (Im, Im, Re, Re) of its mode pair p - the register pairs the packed FP32 instructions work on; PowerTableKernel
                // writes the powers in the same order.
                float largest = 0.f;
                if (audible) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) largest = fmaxf(fmaxf(largest, fmaxf(fabsf(w.Im[i].x), fabsf(w.Im[i].y))), fmaxf(fabsf(w.Re[i].x), fabsf(w.Re[i].y)));
                }
                // (non-negative floats order like their bit patterns: one warp-wide integer max instead of five dependent shuffles)
                const float scale = TmStateScale(__uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(largest))));
                if (lane == 0) *scale_at = scale;
                if (audible) {
                    const float2 scale2 = {scale, scale};
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 im = Mul2(w.Im[i], scale2), re = Mul2(w.Re[i], scale2);
                        const __half2 hi_im = __float22half2_rn(im), hi_re = __float22half2_rn(re);
                        const float2 back_im = __half22float2(hi_im), back_re = __half22float2(hi_re);
                        const __half2 lo_im = __floats2half2_rn(im.x - back_im.x, im.y - back_im.y), lo_re = __floats2half2_rn(re.x - back_re.x, re.y - back_re.y);
                        hi[2 * i] = reinterpret_cast<const uint32_t &>(hi_im), hi[2 * i + 1] = reinterpret_cast<const uint32_t &>(hi_re);
                        lo[2 * i] = reinterpret_cast<const uint32_t &>(lo_im), lo[2 * i + 1] = reinterpret_cast<const uint32_t &>(lo_re);
                    }
                    // the chunk's 64 bytes of the row: FP16 hi x 16, then lo x 16
                    exchange[put + (0 ^ put_swizzle)] = uint4{hi[0], hi[1], hi[2], hi[3]};
                    exchange[put + (1 ^ put_swizzle)] = uint4{hi[4], hi[5], hi[6], hi[7]};
                    exchange[put + (2 ^ put_swizzle)] = uint4{lo[0], lo[1], lo[2], lo[3]};
                    exchange[put + (3 ^ put_swizzle)] = uint4{lo[4], lo[5], lo[6], lo[7]};
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) exchange[put + i] = uint4{0u, 0u, 0u, 0u};
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t from = 8 * j + (lane >> 2); // the chunk-lane whose piece lands at position 32*j + lane
                    walk_at[32 * j] = exchange[from * 4 + ((lane & 3) ^ ((from >> 1) & 3))];
                }
                __syncwarp();
                // next row of the tile, or the first row of the next tile
                walk_at += kTmGroupK / 4, ++scale_at;
                if (++walk_row == plan.WalkBlocksPerTile) {
                    walk_row = 0, walk_at += walk_tile_skip;
                    scale_at += size_t(gridDim.x) * TmScaleTileFloats(plan.WalkBlocksPerTile) - plan.WalkBlocksPerTile;
                }
                if (rendered) {
                    if (step == kTmBlock) {
                        // w <- (hi + lo) w, the small products first
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 lr = Fma2(w.Im[i], jump_lo_nim[i], Mul2(w.Re[i], jump_lo_re[i]));
                            const float2 li = Fma2(w.Re[i], jump_lo_im[i], Mul2(w.Im[i], jump_lo_re[i]));
                            const float2 re = Fma2(w.Im[i], jump_nim[i], Fma2(w.Re[i], jump_re[i], lr));
                            w.Im[i] = Fma2(w.Re[i], jump_im[i], Fma2(w.Im[i], jump_re[i], li));
                            w.Re[i] = re;
                        }
                    } else {
                        float2 sr[4], si[4];
                        PolarPower(b, mode0, step, sr, si); // ragged end of the span: a single step
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 re = {fmaf(-w.Im[i].x, si[i].x, w.Re[i].x * sr[i].x), fmaf(-w.Im[i].y, si[i].y, w.Re[i].y * sr[i].y)};
                            w.Im[i] = {fmaf(w.Re[i].x, si[i].x, w.Im[i].x * sr[i].x), fmaf(w.Re[i].y, si[i].y, w.Im[i].y * sr[i].y)};
                            w.Re[i] = re;
                        }
                    }
                }
                t += step;
            }
        } else {
            for (uint32_t tile = pos; tile < block_end; tile += kTile) {
                const uint32_t nv = min(kTile, block_end - tile);
                if (!warp_renders) {
                    if (lane < nv) partial[tile + lane] = 0.f;
                    continue;
                }
                if (rendered) {
                    const uint32_t tile_abs = plan.FrameBegin + tile;
                    uint32_t s = 0;
                    while (true) {
                        uint32_t lim = nv;
                        if (inj_frame - tile_abs < nv) lim = inj_frame - tile_abs; // inj_frame >= tile_abs + s
                        for (; s + K <= lim; s += K, ++jumps) StepK<K>(w, p, column + s * (kRowPad / 2));
                        for (; s < lim; ++s) Step1<K>(w, p, column + s * (kRowPad / 2));
                        if (s == nv) break;
                        while (inj_frame == tile_abs + s) inject();
                    }
                    if (K > 1 && (jumps >= kDriftJumps || tile + kTile >= block_end)) {
                        ApplyDrift(w, drift, jumps);
                        jumps = 0;
                    }
                }
                // A chunk sitting the block out adds nothing; a muted object (mix gain 0) evolves but its samples are discarded.
                if (!rendered || out_scale == 0.f)
                    for (uint32_t s = 0; s < nv; ++s) column[s * (kRowPad / 2)] = float2{0.f, 0.f};
                __syncwarp();
                const float total = lane < nv ? SumRow(rows, lane) : 0.f;
                if (lane < nv) partial[tile + lane] = total;
                __syncwarp();
            }

        }

        // End of the block: chunk energies decide the audible prefix, or silence the object (:132-146).
        float energy = 0.f;
        if (rendered) {
#pragma unroll
            for (int i = 0; i < 4; ++i) energy += (w.Re[i].x * w.Re[i].x + w.Im[i].x * w.Im[i].x) + (w.Re[i].y * w.Re[i].y + w.Im[i].y * w.Im[i].y);
        }
        const bool audible_chunk = rendered && energy * energy_scale >= kSilentEnergy;
        cull_energy[threadIdx.x] = energy;
        cull_audible[threadIdx.x] = audible_chunk;
        // Common case: every tuned chunk of every culled object in the CTA is audible. Then each object's audible prefix
        // is its whole tuned set and its total energy is above the threshold, so the per-object scan below is skipped.
        const bool settled = !(valid && cull) || (in_tuned ? audible_chunk : obj_tuned != 0);
        if (__syncthreads_and(settled)) {
            if (valid && cull) {
                ringing = true;
                live = excited || in_tuned;
            }
        } else if (valid && cull) {
            float total = 0.f;
            int last = -1;
            for (uint32_t k = 0; k < obj_chunks; ++k) {
                total += cull_energy[obj_first_local + k];
                if (cull_audible[obj_first_local + k]) last = int(k);
            }
            // A decision that removes state from the linear evolution invalidates the scan along time for every later
            // segment; the window's very last block only feeds the next window, which starts from the real flags.
            const bool feeds_scan = !(last_segment && block_end == seg_end);
            if (!excited && total * energy_scale < kSilentEnergy) { // SilenceObject :53-64
                if (total > 0.f && feeds_scan) atomicOr(plan.Speculation, 4u);
#pragma unroll
                for (int i = 0; i < 4; ++i) w.Re[i] = w.Im[i] = float2{0.f, 0.f};
                live = true;
                ringing = false;
            } else {
                ringing = true;
                live = excited || int(my_chunk) <= last;
                if (!live && in_tuned && energy > 0.f && feeds_scan) {
                    if (plan.Debug && atomicOr(plan.Speculation, 2u) == 0) printf("[me] chunk frozen with state: object %u chunk %u seg %u frame %u last %d total %g energy %g scale %g\n", object, my_chunk, seg, block_end_abs, last, total, energy, energy_scale);
                    atomicOr(plan.Speculation, 2u);
                }
            }
        }
        __syncthreads();
        pos = block_end;
    }

    if (valid && last_segment) {
        // Increments of pulses that end exactly with the span belong to the state the next call adopts.
        if (plan.FrameBegin + seg_end == plan.SpanFrames && in_tuned)
            while (inj_frame == plan.SpanFrames) inject();
        float2 zr[4], zi[4], qr[4], qi[4];
        Load8(b.PhaseIm + mode0, qr), Load8(b.PhaseRe + mode0, qi);
        const float inv_gain = 1.f / gain;
#pragma unroll
        for (int i = 0; i < 4; ++i) { // z = w * conj(q) / gain
            zr[i] = {(w.Re[i].x * qr[i].x + w.Im[i].x * qi[i].x) * inv_gain, (w.Re[i].y * qr[i].y + w.Im[i].y * qi[i].y) * inv_gain};
            zi[i] = {(w.Im[i].x * qr[i].x - w.Re[i].x * qi[i].x) * inv_gain, (w.Im[i].y * qr[i].y - w.Re[i].y * qi[i].y) * inv_gain};
        }
        Store8(b.StateOutRe + mode0, zr), Store8(b.StateOutIm + mode0, zi);
        b.ChunkLiveOut[chunk] = live;
        if (my_chunk == 0) b.ObjRingingOut[object] = ringing;
    }
}

// The impulse of an impact projected onto the chunk's 8 mode shapes (ImpactGainRow, ModalAudio.h:182-188), in the
// reference's operation order, then rotated by the output phase and scaled by the mix gain.
__device__ __forceinline__ void ImpactGain(const BankView &b, const DevImpact &im, uint32_t mode0, uint32_t first_mode, const float2 (&qr)[4], const float2 (&qi)[4], float scale, Chunk &g) {
    const uint32_t base = b.ObjShapeOffset[im.Object] + im.ExPos * b.ObjStride[im.Object] + first_mode;
    float2 sx[4], sy[4], sz[4], rg[4];
    Load8(b.ShapeX + base, sx), Load8(b.ShapeY + base, sy), Load8(b.ShapeZ + base, sz), Load8(b.RadiationGain + mode0, rg);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float gx = __fmul_rn(rg[i].x, __fadd_rn(__fadd_rn(__fmul_rn(sx[i].x, im.Jx), __fmul_rn(sy[i].x, im.Jy)), __fmul_rn(sz[i].x, im.Jz)));
        const float gy = __fmul_rn(rg[i].y, __fadd_rn(__fadd_rn(__fmul_rn(sx[i].y, im.Jx), __fmul_rn(sy[i].y, im.Jy)), __fmul_rn(sz[i].y, im.Jz)));
        g.Re[i] = {gx * qr[i].x * scale, gy * qr[i].y * scale};
        g.Im[i] = {gx * qi[i].x * scale, gy * qi[i].y * scale};
    }
}

// One warp per (impact, 32 chunks of its object): the zero-state response to the force pulse, v <- c v + f g.
constexpr uint32_t kPulseWarps = 4;
__global__ void __launch_bounds__(kPulseWarps * 32) PulseKernel(const BankView b, const PulsePlan plan) {
    __shared__ __align__(16) float transposed[kPulseWarps][kTile][kRowPad];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t pw = blockIdx.x * kPulseWarps + warp;
    if (pw >= plan.NPulseWarps) return;
    const PulseWarp job = plan.Warps[pw];
    const DevImpact im = plan.Impacts[job.Impact];
    const uint32_t my_chunk = job.Chunk0 + lane;
    const bool valid = my_chunk < b.ObjStride[im.Object] / kLanes;
    const uint32_t mode0 = (b.ObjFirstChunk[im.Object] + my_chunk) * kLanes;

    float2 cre[4], cim[4];
    Chunk g, v;
    float out_scale = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) cre[i] = cim[i] = g.Re[i] = g.Im[i] = v.Re[i] = v.Im[i] = float2{0.f, 0.f};
    if (valid) {
        float2 qr[4], qi[4];
        Load8(b.CoeffRe + mode0, cre), Load8(b.CoeffIm + mode0, cim);
        Load8(b.PhaseIm + mode0, qr), Load8(b.PhaseRe + mode0, qi);
        const float mix = b.ObjMixGain[im.Object];
        out_scale = mix == 0.f ? 0.f : 1.f;
        ImpactGain(b, im, mode0, my_chunk * kLanes, qr, qi, mix == 0.f ? 1.f : mix, g);
    }
    float *rows = &transposed[warp][0][0];
    const float *force = plan.Force + im.ForceOff;
    // v <- c v + f g on mode pairs (packed FFMA2): three packed operations per component pair and sample.
    float2 ncim[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ncim[i] = {-cim[i].x, -cim[i].y};
    const float2 scale2 = {out_scale, out_scale};
    // Past the pulse (tensor spans: up to the frame the increment is injected at) the response rings freely: those samples
    // take the sample loop's K-step form, 2.75 lane-operations per mode-sample instead of 7. The kernel runs in two phases -
    // tiles with force samples, then pure ringing - so that the powers of the K-step form (72 registers) are not alive while the
    // forced recurrence runs: 5 CTAs fit an SM instead of 3, and two of them fit beside a CTA of the state walk.
    constexpr int K = 4;
    const auto flush = [&](uint32_t tile, uint32_t nv) { // the warp's nv samples of this tile: summed over its chunks, stored
        __syncwarp();
        if (lane < nv) plan.Rows[job.RowOff + tile + lane] = SumRow(rows, lane);
        __syncwarp();
    };
    uint32_t tile = 0;
    for (; tile < im.Len; tile += kTile) { // tiles holding force samples (the last one may already ring for part of its length)
        const uint32_t nv = min(kTile, im.RenderLen - tile);
        const uint32_t forced = min(nv, im.Len - tile);
        for (uint32_t s = 0; s < forced; ++s) {
            const float f = __ldg(force + tile + s);
            const float2 f2 = {f, f};
            float2 sum = {0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 re = Fma2(f2, g.Re[i], Fma2(v.Re[i], cre[i], Mul2(v.Im[i], ncim[i])));
                v.Im[i] = Fma2(f2, g.Im[i], Fma2(v.Re[i], cim[i], Mul2(v.Im[i], cre[i])));
                v.Re[i] = re;
                sum = Add2(sum, v.Im[i]);
            }
            reinterpret_cast<float2 *>(rows)[s * (kRowPad / 2) + lane] = Mul2(sum, scale2);
        }
        for (uint32_t s = forced; s < nv; ++s) { // the rest of the transition tile rings freely, one sample at a time
            float2 sum = {0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 re = Fma2(v.Re[i], cre[i], Mul2(v.Im[i], ncim[i]));
                v.Im[i] = Fma2(v.Re[i], cim[i], Mul2(v.Im[i], cre[i]));
                v.Re[i] = re;
                sum = Add2(sum, v.Im[i]);
            }
            reinterpret_cast<float2 *>(rows)[s * (kRowPad / 2) + lane] = Mul2(sum, scale2);
        }
        flush(tile, nv);
    }
    if (tile < im.RenderLen) {
        Powers<K> p;
        MakePowers<K>(cre, cim, p);
        for (; tile < im.RenderLen; tile += kTile) {
            const uint32_t nv = min(kTile, im.RenderLen - tile);
            float2 *column = reinterpret_cast<float2 *>(rows) + lane;
            uint32_t s = 0;
            for (; s + K <= nv; s += K) StepK<K>(v, p, column + s * (kRowPad / 2));
            for (; s < nv; ++s) Step1<K>(v, p, column + s * (kRowPad / 2));
            if (out_scale == 0.f) // a muted object evolves but its samples are discarded
                for (s = 0; s < nv; ++s) column[s * (kRowPad / 2)] = float2{0.f, 0.f};
            flush(tile, nv);
        }
    }
    if (valid) {
        const uint32_t off = im.DeltaOff + my_chunk * kLanes;
        Store8(plan.DeltaRe + off, v.Re), Store8(plan.DeltaIm + off, v.Im);
    }
}

// Block-parallel scan along time: one thread per mode walks the window's segment boundaries, carrying the rotated
// state in FP64 and jumping between events with c^m = exp(m ln|c|) (cos m*arg c + i sin m*arg c).
__global__ void __launch_bounds__(128) SegmentScanKernel(const BankView b, const RenderPlan plan, float *__restrict__ seg_re, float *__restrict__ seg_im) {
    const uint32_t mode = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode >= b.NChunks * kLanes) return;
    const uint32_t chunk = mode / kLanes, lane_in_chunk = mode % kLanes;
    const uint32_t object = b.ChunkObject[chunk];
    const size_t stride = size_t(b.NChunks) * kLanes;
    if (object == kNoObject) {
        for (uint32_t s = 1; s < plan.NSegments; ++s) seg_re[size_t(s - 1) * stride + mode] = seg_im[size_t(s - 1) * stride + mode] = 0.f;
        return;
    }
    const uint32_t my_chunk = chunk - b.ObjFirstChunk[object];
    const float mix = b.ObjMixGain[object];
    const double gain = mix == 0.f ? 1.0 : double(mix);
    const double zr = b.StateRe[mode], zi = b.StateIm[mode], qr = b.PhaseIm[mode], qi = b.PhaseRe[mode];
    double wr = (zr * qr - zi * qi) * gain, wi = (zr * qi + zi * qr) * gain;
    const double log_rho = b.LogRho[mode], theta = b.Theta[mode];
    const auto jump = [&](uint32_t m) {
        if (m == 0) return;
        const double mag = exp(double(m) * log_rho); // ln 0 = -inf -> 0
        double sn, cs;
        sincos(double(m) * theta, &sn, &cs);
        const double re = (wr * cs - wi * sn) * mag, im = (wr * sn + wi * cs) * mag;
        wr = re, wi = im;
    };
    uint32_t inj = plan.ObjInjectPtr[object];
    const uint32_t inj_hi = plan.ObjInjectPtr[object + 1];
    while (inj < inj_hi && plan.InjectFrame[inj] < plan.FrameBegin) ++inj;
    uint32_t cursor = plan.FrameBegin;
    for (uint32_t s = 1; s < plan.NSegments; ++s) {
        const uint32_t boundary = plan.FrameBegin + s * plan.SegmentFrames;
        // Increments landing strictly before the boundary; one landing on it is added by the segment's own thread.
        for (; inj < inj_hi && plan.InjectFrame[inj] < boundary; ++inj) {
            jump(plan.InjectFrame[inj] - cursor);
            cursor = plan.InjectFrame[inj];
            const uint32_t off = plan.InjectDelta[inj] + my_chunk * kLanes + lane_in_chunk;
            wr += double(plan.DeltaRe[off]), wi += double(plan.DeltaIm[off]);
        }
        jump(boundary - cursor);
        cursor = boundary;
        seg_re[size_t(s - 1) * stride + mode] = float(wr), seg_im[size_t(s - 1) * stride + mode] = float(wi);
    }
}

// out[n] = sum of the warp rows in fixed order + the pulse rows overlapping n (sorted by start, so also fixed order).
// A CTA owns 32 consecutive frames; its 8 warps each sum a fixed residue class of the rows, coalesced along n, and
// the eight partial sums are combined in warp order.
constexpr uint32_t kMixWarps = 8;
__global__ void __launch_bounds__(kMixWarps * 32) MixKernel(const float *__restrict__ partial, uint32_t rows, const RenderPlan plan, const PulsePlan pulses, float *__restrict__ out) {
    __shared__ float sums[kMixWarps][32], pulse_sums[kMixWarps][32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = blockIdx.x * 32 + lane;
    const bool in_range = n < plan.Frames;
    float a = 0.f, b2 = 0.f, c = 0.f, d = 0.f;
    if (in_range) {
        uint32_t r = warp;
        for (; r + 3 * kMixWarps < rows; r += 4 * kMixWarps) {
            a += partial[size_t(r) * plan.Frames + n];
            b2 += partial[size_t(r + kMixWarps) * plan.Frames + n];
            c += partial[size_t(r + 2 * kMixWarps) * plan.Frames + n];
            d += partial[size_t(r + 3 * kMixWarps) * plan.Frames + n];
        }
        for (; r < rows; r += kMixWarps) a += partial[size_t(r) * plan.Frames + n];
    }
    // The pulse rows overlapping n, sorted by start: warp k adds every eighth one, again a fixed order.
    float extra = 0.f;
    if (in_range && pulses.NPulseWarps) {
        const uint32_t n_abs = plan.FrameBegin + n;
        // First pulse-warp whose start could still cover n_abs.
        const uint32_t earliest = n_abs >= pulses.MaxLen ? n_abs - pulses.MaxLen + 1 : 0;
        uint32_t lo = 0, hi = pulses.NPulseWarps;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (pulses.Warps[mid].Start < earliest) lo = mid + 1;
            else hi = mid;
        }
        for (uint32_t i = lo + warp; i < pulses.NPulseWarps; i += kMixWarps) {
            const PulseWarp pw = pulses.Warps[i];
            if (pw.Start > n_abs) break;
            if (n_abs - pw.Start < pw.RenderLen) extra += pulses.Rows[pw.RowOff + (n_abs - pw.Start)];
        }
    }
    sums[warp][lane] = (a + b2) + (c + d);
    pulse_sums[warp][lane] = extra;
    __syncthreads();
    if (warp != 0 || !in_range) return;
    float sum = 0.f, pulse_sum = 0.f;
#pragma unroll
    for (uint32_t k = 0; k < kMixWarps; ++k) sum += sums[k][lane], pulse_sum += pulse_sums[k][lane];
    out[n] = sum + pulse_sum;
}

// One thread per impact: the raised-cosine force curve from a unit-circle rotor, with the reference's own float
// operation order and no FMA contraction (ModalAudio.cpp:517-526), so the curve is bit-identical.
__global__ void ForceKernel(const DevImpact *__restrict__ impacts, const DevImpactTail *__restrict__ tails, uint32_t n, float *__restrict__ force) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevImpact im = impacts[i];
    const float half_gamma = __fmul_rn(tails[i].Gamma, 0.5f);
    float pr = im.PhaseRe, pi = im.PhaseIm;
    float *f = force + im.ForceOff;
    for (uint32_t s = 0; s < im.Len; ++s) {
        const float re = __fsub_rn(__fmul_rn(pr, im.RotRe), __fmul_rn(pi, im.RotIm));
        pi = __fadd_rn(__fmul_rn(pr, im.RotIm), __fmul_rn(pi, im.RotRe));
        pr = re;
        f[s] = __fmul_rn(half_gamma, __fsub_rn(1.f, pr));
    }
}

// One thread per click-carrying impact: the coupled recoil biquad driven by the force pulse (ModalAudio.cpp:527-531).
__global__ void ClickKernel(const DevImpact *__restrict__ impacts, const DevImpactTail *__restrict__ tails, uint32_t n, float *__restrict__ out, uint32_t frames) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevImpact im = impacts[i];
    if (!im.HasClick) return;
    const DevImpactTail t = tails[i];
    const float half_gamma = __fmul_rn(t.Gamma, 0.5f);
    float pr = im.PhaseRe, pi = im.PhaseIm, z1 = t.ClickZ1, z2 = t.ClickZ2;
    const uint32_t end = min(im.End, frames);
    for (uint32_t s = im.Start; s < end; ++s) {
        float cur = 0.f;
        if (s - im.Start < im.Len) {
            const float re = __fsub_rn(__fmul_rn(pr, im.RotRe), __fmul_rn(pi, im.RotIm));
            pi = __fadd_rn(__fmul_rn(pr, im.RotIm), __fmul_rn(pi, im.RotRe));
            pr = re;
            cur = __fmul_rn(half_gamma, __fsub_rn(1.f, pr));
        }
        const float u = __fmul_rn(t.AccelAmp, cur);
        const float y = __fadd_rn(__fmul_rn(t.ClickB0, u), z1);
        z1 = __fadd_rn(__fmul_rn(-t.ClickA1, y), z2);
        z2 = __fsub_rn(__fmul_rn(-t.ClickB0, u), __fmul_rn(t.ClickA2, y));
        atomicAdd(out + s, __fmul_rn(y, t.ClickGain));
    }
}

// One thread per mode pair: rows j = 1..kTmBlock of the power stages, c^j by FP64 products of the float coefficient
// (the same arithmetic as MakePowers). A stage (one chunk) holds two FP16 images of its 256 x 16 block (tensor_mix.cuh):
// hi = fp16(value) and lo = fp16(value - hi).
__global__ void __launch_bounds__(128) PowerTableKernel(const BankView b, float *__restrict__ powers) {
    static_assert(kTmKChunk == 16, "one chunk per stage");
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; // chunk * 4 + pair
    if (idx >= b.NChunks * 4) return;
    const uint32_t chunk = idx >> 2, pair = idx & 3;
    const uint32_t mode = chunk * kLanes + pair * 2;
    const double ax = b.CoeffRe[mode], bx = b.CoeffIm[mode], ay = b.CoeffRe[mode + 1], by = b.CoeffIm[mode + 1];
    uint8_t *stage = reinterpret_cast<uint8_t *>(powers + size_t(chunk) * TmPowerStageFloats());
    uint8_t *hi16 = stage + size_t(pair >> 1) * kTmBlock * 16 + (pair & 1) * 8; // pieces of 8 elements, 4096 bytes apart
    uint8_t *lo16 = hi16 + kTmPowerImageBytes;
    double rx = ax, ix = bx, ry = ay, iy = by;
    for (uint32_t j = 0; j < kTmBlock; ++j) {
        const float4 v = {float(rx), float(ry), float(ix), float(iy)}; // against the state's (Im, Im, Re, Re) of the pair
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
        const uint32_t row = (j >> 3) * 128 + (j & 7) * 16;
        *reinterpret_cast<uint2 *>(hi16 + row) = {reinterpret_cast<const uint32_t &>(h0), reinterpret_cast<const uint32_t &>(h1)};
        *reinterpret_cast<uint2 *>(lo16 + row) = {reinterpret_cast<const uint32_t &>(l0), reinterpret_cast<const uint32_t &>(l1)};
        const double nrx = rx * ax - ix * bx, nry = ry * ay - iy * by;
        ix = rx * bx + ix * ax, iy = ry * by + iy * ay;
        rx = nrx, ry = nry;
    }
}

// ---- FP32 issue-rate micro-benchmark --------------------------------------------------------------------------

// Mode 0: scalar FFMA. 1: packed FFMA2. 2: FFMA2 and FFMA interleaved 1:1 (instructions). 3: FFMA2:FFMA 1:2. 4: scalar FADD.
template<int Mode>
__global__ void __launch_bounds__(256) FmaRateKernel(float *sink, int inner) {
    float2 a[8], b = {1.0000001f, 0.9999999f}, c = {1e-9f, -1e-9f};
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = {Mode >= 5 ? 1e-3f * float(threadIdx.x + i) : float(threadIdx.x + i), Mode >= 5 ? 1e-3f * float(i) : float(i)}, d[i] = float(i) * 0.5f;
    for (int it = 0; it < inner; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if constexpr (Mode == 1) {
                a[i] = Fma2(a[i], b, c);
            } else if constexpr (Mode == 0) {
                a[i].x = fmaf(a[i].x, b.x, c.x);
                a[i].y = fmaf(a[i].y, b.y, c.y);
            } else if constexpr (Mode == 2) {
                if (i < 4) a[i] = Fma2(a[i], b, c);
                else d[i] = fmaf(d[i], b.x, c.x);
            } else if constexpr (Mode == 3) {
                if (i < 4) a[i] = Fma2(a[i], b, c);
                else d[i] = fmaf(d[i], b.x, c.x), d[i - 4] = fmaf(d[i - 4], b.y, c.y);
            } else if constexpr (Mode == 4) {
                a[i].x = a[i].x + c.x;
                a[i].y = a[i].y + c.y;
            } else if constexpr (Mode == 5) { // three fresh register pairs per FFMA2, no operand reuse possible
                a[i] = Fma2(a[(i + 3) & 7], a[(i + 5) & 7], a[i]);
            } else if constexpr (Mode == 6) { // two fresh pairs, one shared with the previous instruction's slot
                a[i] = Fma2(b, a[(i + 5) & 7], a[i]);
            } else { // scalar FFMA with three fresh registers
                a[i].x = fmaf(a[(i + 3) & 7].x, a[(i + 5) & 7].y, a[i].x);
                a[i].y = fmaf(a[(i + 3) & 7].y, a[(i + 5) & 7].x, a[i].y);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + d[i];
    if (s == 123.456f) sink[0] = s;
}

} // namespace

uint32_t ResonatorRows(uint32_t n_chunks) { return (n_chunks + kBlockThreads - 1) / kBlockThreads * kWarpsPerBlock; }

void LaunchForceKernel(const DevImpact *impacts, const DevImpactTail *tails, uint32_t n, float *force, cudaStream_t stream, LaunchCounter &counter) {
    if (n == 0) return;
    ForceKernel<<<(n + 127) / 128, 128, 0, stream>>>(impacts, tails, n, force);
    ME_CUDA(cudaGetLastError());
    ++counter.Launches;
}

// Dynamic shared memory that leaves room for exactly `per_sm` CTAs of a kernel on an SM (0: no limit). The pulse kernels run
// beside the state walk: how many CTAs of each an SM takes decides how much of one hides behind the other.
static size_t OccupancyPad(const char *env, int fallback, size_t static_bytes) {
    const char *v = std::getenv(env);
    const int per_sm = v ? std::atoi(v) : fallback;
    if (per_sm <= 0) return 0;
    const size_t per_cta = (size_t(227) << 10) / size_t(per_sm + 1) + 1024; // more than a (per_sm + 1)-th of the SM's 227 KB
    return per_cta > static_bytes ? std::min<size_t>(per_cta - static_bytes, (size_t(227) << 10) - static_bytes) : 0;
}

void LaunchPulseKernel(const BankView &bank, const PulsePlan &plan, cudaStream_t stream, LaunchCounter &counter) {
    if (plan.NPulseWarps == 0) return;
    static const size_t pad = [] {
        const size_t bytes = OccupancyPad("ME_PULSE_CTAS_PER_SM", 0, sizeof(float) * kPulseWarps * kTile * kRowPad);
        if (bytes) ME_CUDA(cudaFuncSetAttribute(PulseKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
        return bytes;
    }();
    PulseKernel<<<(plan.NPulseWarps + kPulseWarps - 1) / kPulseWarps, kPulseWarps * 32, pad, stream>>>(bank, plan);
    ME_CUDA(cudaGetLastError());
    ++counter.Launches;
}

void LaunchSegmentScan(const BankView &bank, const RenderPlan &plan, float *seg_re, float *seg_im, cudaStream_t stream, LaunchCounter &counter) {
    if (plan.NSegments <= 1 || bank.NChunks == 0) return;
    SegmentScanKernel<<<(bank.NChunks * kLanes + 127) / 128, 128, 0, stream>>>(bank, plan, seg_re, seg_im);
    ME_CUDA(cudaGetLastError());
    ++counter.Launches;
}

void LaunchResonatorKernel(const BankView &bank, const RenderPlan &plan, int steps, cudaStream_t stream, LaunchCounter &counter) {
    if (bank.NChunks == 0 || plan.Frames == 0) return;
    const dim3 grid((bank.NChunks + kBlockThreads - 1) / kBlockThreads, plan.NSegments);
    constexpr size_t smem = sizeof(float) * (kWarpsPerBlock * kTile * kRowPad + 16 * kBlockThreads); // transpose tiles + the drift terms
    static const bool wide = [] {
        const char *occ = std::getenv("ME_RESONATOR_OCC");
        return occ && occ[0] == '1';
    }();
    static bool configured = false;
    if (!configured) {
        ME_CUDA(cudaFuncSetAttribute(ResonatorKernel<1, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        ME_CUDA(cudaFuncSetAttribute(ResonatorKernel<2, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        ME_CUDA(cudaFuncSetAttribute(ResonatorKernel<4, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        ME_CUDA(cudaFuncSetAttribute(ResonatorKernel<4, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = true;
    }
    switch (steps) {
        case 1: ResonatorKernel<1, 2, false><<<grid, kBlockThreads, smem, stream>>>(bank, plan); break;
        case 2: ResonatorKernel<2, 2, false><<<grid, kBlockThreads, smem, stream>>>(bank, plan); break;
        default:
            if (wide) ResonatorKernel<4, 1, false><<<grid, kBlockThreads, smem, stream>>>(bank, plan);
            else ResonatorKernel<4, 2, false><<<grid, kBlockThreads, smem, stream>>>(bank, plan);
            break;
    }
    ME_CUDA(cudaGetLastError());
    ++counter.Launches;
}

void LaunchStateWalkKernel(const BankView &bank, const RenderPlan &plan, cudaStream_t stream, LaunchCounter &counter) {
    if (bank.NChunks == 0 || plan.Frames == 0) return;
    const dim3 grid(bank.NChunks / kBlockThreads, plan.NSegments);
    static const size_t smem = [] {
        const size_t used = kWarpsPerBlock * 128 * sizeof(float4), fixed = kBlockThreads * 5;
        const size_t bytes = used + OccupancyPad("ME_WALK_CTAS_PER_SM", 0, used + fixed);
        if (bytes > used) ME_CUDA(cudaFuncSetAttribute(ResonatorKernel<1, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
        return bytes;
    }();
    ResonatorKernel<1, 2, true><<<grid, kBlockThreads, smem, stream>>>(bank, plan);
    ME_CUDA(cudaGetLastError());
    ++counter.Launches;
}

void LaunchPowerTableKernel(const BankView &bank, float *powers, cudaStream_t stream, LaunchCounter &counter) {
    if (bank.NChunks == 0) return;
    PowerTableKernel<<<(bank.NChunks * 4 + 127) / 128, 128, 0, stream>>>(bank, powers);
    ME_CUDA(cudaGetLastError());
    ++counter.Launches;
}

void LaunchMixKernel(const float *partial, uint32_t rows, const RenderPlan &plan, const PulsePlan &pulses, float *out, cudaStream_t stream, LaunchCounter &counter) {
    if (plan.Frames == 0) return;
    // (A four-frames-per-lane form with 16-byte loads was measured 0.3 ms SLOWER on configs[4]: the pulse rows of four frames walked
    // one after the other by each lane cost more than the wider row loads saved.)
    MixKernel<<<(plan.Frames + 31) / 32, kMixWarps * 32, 0, stream>>>(partial, rows, plan, pulses, out);
    ME_CUDA(cudaGetLastError());
    ++counter.Launches;
}

void LaunchClickKernel(const DevImpact *impacts, const DevImpactTail *tails, uint32_t n, float *out, uint32_t frames, cudaStream_t stream, LaunchCounter &counter) {
    if (n == 0) return;
    ClickKernel<<<(n + 127) / 128, 128, 0, stream>>>(impacts, tails, n, out, frames);
    ME_CUDA(cudaGetLastError());
    ++counter.Launches;
}

double MeasureFmaRate(int mode, int iters) {
    int device = 0, sms = 0;
    ME_CUDA(cudaGetDevice(&device));
    ME_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float *sink = nullptr;
    ME_CUDA(cudaMalloc(&sink, sizeof(float)));
    const int inner = 4096, blocks = sms * 8, threads = 256;
    cudaEvent_t start, stop;
    ME_CUDA(cudaEventCreate(&start));
    ME_CUDA(cudaEventCreate(&stop));
    auto launch = [&] {
        switch (mode) {
            case 0: FmaRateKernel<0><<<blocks, threads>>>(sink, inner); break;
            case 1: FmaRateKernel<1><<<blocks, threads>>>(sink, inner); break;
            case 2: FmaRateKernel<2><<<blocks, threads>>>(sink, inner); break;
            case 3: FmaRateKernel<3><<<blocks, threads>>>(sink, inner); break;
            case 4: FmaRateKernel<4><<<blocks, threads>>>(sink, inner); break;
            case 5: FmaRateKernel<5><<<blocks, threads>>>(sink, inner); break;
            case 6: FmaRateKernel<6><<<blocks, threads>>>(sink, inner); break;
            default: FmaRateKernel<7><<<blocks, threads>>>(sink, inner); break;
        }
    };
    for (int i = 0; i < 3; ++i) launch();
    ME_CUDA(cudaEventRecord(start));
    for (int i = 0; i < iters; ++i) launch();
    ME_CUDA(cudaEventRecord(stop));
    ME_CUDA(cudaEventSynchronize(stop));
    float ms = 0.f;
    ME_CUDA(cudaEventElapsedTime(&ms, start, stop));
    cudaEventDestroy(start), cudaEventDestroy(stop), cudaFree(sink);
    // Lane-operations per thread per inner iteration.
    const double per_iter = mode == 2 ? 12.0 : 16.0;
    return double(iters) * blocks * threads * double(inner) * per_iter / (double(ms) * 1e-3);
}

} // namespace me
