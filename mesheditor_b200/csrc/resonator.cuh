// Device-side views and launchers of the resonator bank kernels (resonator.cu).
// Reference for every kernel: src/audio/ModalAudio.cpp (RenderModal :486-590, RenderObjectFast :86-147).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace me {

constexpr uint32_t kLanes = 8;          // modes per chunk == the reference's Lanes (ModalAudio.h:169)
constexpr uint32_t kTile = 32;          // samples reduced across a warp at a time
constexpr uint32_t kBlockThreads = 256; // chunk-threads per CTA; an object of <= 256 chunks never straddles CTAs
constexpr uint32_t kWarpsPerBlock = kBlockThreads / 32;
constexpr uint32_t kNoObject = 0xFFFFFFFFu;

// The installed bank in HBM. Per-mode columns are padded per object to whole 8-mode chunks (padding: coefficient
// 0, gain 0, state 0) and objects are placed so that they do not straddle a CTA's 256 chunk slots; unused slots
// carry ChunkObject == kNoObject.
struct BankView {
    uint32_t NChunks;                       // chunk slots, padding included
    uint32_t NObjects;
    const float *CoeffRe, *CoeffIm;         // [NChunks*8]
    const float *StateRe, *StateIm;         // [NChunks*8] resonator state z at the start of the window
    float *StateOutRe, *StateOutIm;         // [NChunks*8] state at the end of the window (ping-pong with the above)
    const float *PhaseIm, *PhaseRe;         // OutPhaseIm / OutPhaseRe (the output rotation q = PhaseIm + i*PhaseRe)
    const float *RadiationGain;             // [NChunks*8]
    const double *LogRho, *Theta;           // [NChunks*8] polar form of the coefficient in FP64 (for c^m)
    const uint32_t *ChunkObject;            // [NChunks] owning object or kNoObject
    const float *ShapeX, *ShapeY, *ShapeZ;  // [object][ex_pos][stride], stride = padded mode count
    const uint32_t *ObjShapeOffset;         // [NObjects]
    const uint32_t *ObjStride;              // [NObjects] padded mode count
    const uint32_t *ObjFirstChunk;          // [NObjects] slot of the object's first chunk
    const uint32_t *ObjTunedChunks;         // [NObjects] chunks holding TunedModeCount modes
    const float *ObjMixGain;                // [NObjects] OutGain*ListenerGain
    const float *ObjEnergyScale;            // [NObjects] OutGain^2 / MixGain^2 (0 when muted): chunk energy -> audibility
    const uint8_t *ObjCull;                 // [NObjects] 1 when the audibility culling of :139-146 applies (object fits one CTA)
    const uint8_t *ChunkLive;               // [NChunks] chunk is inside LiveModeCount
    const uint8_t *ObjRinging;              // [NObjects] ModalBank::Ringing
    uint8_t *ChunkLiveOut, *ObjRingingOut;  // the same at the end of the window
};

// One contact pulse as the force / pulse / click kernels see it.
struct alignas(16) DevImpact {
    uint32_t Start;      // first frame of the pulse, relative to the span
    uint32_t Len;        // force samples generated in this span (<= SamplesLeft)
    uint32_t ForceOff;   // offset of its force curve in the force buffer
    uint32_t ExPos;
    float Jx, Jy, Jz;
    uint32_t Object;
    float PhaseRe, PhaseIm, RotRe, RotIm; // rotor state at Start and its per-sample rotation
    uint32_t End;        // frame at which the impact is retired, or the span's end
    uint32_t DeltaOff;   // offset of its end-of-pulse state increment in the delta buffers
    uint32_t HasClick;
    uint32_t RenderLen;  // samples the pulse kernel renders on its own (>= Len): up to the frame its increment is injected at
};
struct alignas(16) DevImpactTail {
    float Gamma, AccelAmp, ClickB0, ClickA1;
    float ClickA2, ClickZ1, ClickZ2, ClickGain; // ClickGain = ModalAudio::ClickGain * ListenerGain[object]
};
// One warp of the pulse kernel: 32 chunks of one impact's object.
struct PulseWarp {
    uint32_t Impact;     // index into the DevImpact array
    uint32_t Chunk0;     // first chunk (inside the object) this warp covers
    uint32_t RowOff;     // offset of its RenderLen output samples in the pulse rows
    uint32_t Start;      // == Impacts[Impact].Start and .RenderLen, so the mix kernel reads one record per pulse-warp
    uint32_t RenderLen;
    uint32_t Pad;
};

struct RenderPlan {
    uint32_t SpanFrames;       // frames of the span (block grid and impact frames are relative to its start)
    uint32_t BlockFrames;      // RenderModal block length: culling decisions fall on multiples of it
    uint32_t FrameBegin;       // first frame of this launch window inside the span
    uint32_t Frames;           // frames in this launch window
    uint32_t NSegments;        // block-parallel scan along time: segments per chunk-thread inside the window
    uint32_t SegmentFrames;    // frames per segment
    // Per object: injections of end-of-pulse state increments, sorted by frame; and the merged intervals during
    // which the object holds a live impact ("excited": the whole tuned set renders, :90,:146).
    const uint32_t *ObjInjectPtr;  // [NObjects+1]
    const uint32_t *InjectFrame;   // frame before which the increment is added
    const uint32_t *InjectDelta;   // offset into DeltaRe/DeltaIm of the object's chunk 0
    const uint32_t *ObjExcitePtr;  // [NObjects+1]
    const uint32_t *ExciteBegin, *ExciteEnd;
    const float *DeltaRe, *DeltaIm;
    float *Partial;                // [warp rows][Frames] per-warp partial mixes
    const float *SegStateRe, *SegStateIm; // [NSegments-1][NChunks*8] rotated state at the start of segments 1..
    // Set to 1 when a segment met a culling decision the scan along time could not foresee (a frozen chunk or a
    // silenced object with state left): the window is then rendered again sequentially in time.
    uint32_t *Speculation;
    // Launch-time condition: when non-zero the kernel runs only if (*Speculation & OnlyIf) != 0 and returns at once otherwise.
    // The sequential repeat of a seeded walk is launched this way right behind it, so that no host round trip sits between the
    // walk and the tcgen05 mix: the device decides.
    uint32_t OnlyIf;
    uint32_t Debug;
    // Tensor-core form (tensor_mix.cuh): the walk kernel (one segment: sequential in time, so culling is exact) writes
    // the block-start state of every 256-frame time block (kTmBlock) as rows of States[tile][chunk group][head,tail][block][4096].
    float *WalkStates;
    float *WalkScales;         // [tile][chunk group][walk warp (8)][time block]: the FP16 scale of each warp's part of each row
    uint32_t WalkBlocksPerTile;
};

struct PulsePlan {
    uint32_t NPulseWarps;
    const PulseWarp *Warps;        // sorted by the impact's Start
    const DevImpact *Impacts;
    const float *Force;
    float *Rows;                   // per pulse-warp output samples
    float *DeltaRe, *DeltaIm;
    uint32_t MaxLen;               // longest pulse of the span
};

struct LaunchCounter {
    uint32_t Launches{0};
};

// Force curves of every impact (ModalAudio.cpp:506-526), bit-exact with the reference's float recurrence.
void LaunchForceKernel(const DevImpact *impacts, const DevImpactTail *tails, uint32_t n_impacts, float *force, cudaStream_t, LaunchCounter &);
// Zero-state response of every contact pulse: its samples (per pulse-warp rows) and its end-of-pulse state increment.
void LaunchPulseKernel(const BankView &, const PulsePlan &, cudaStream_t, LaunchCounter &);
// States at the start of segments 1.. of the window from the bank state and the pulses' increments (FP64 powers).
void LaunchSegmentScan(const BankView &, const RenderPlan &, float *seg_re, float *seg_im, cudaStream_t, LaunchCounter &);
// The free-running resonator bank over the window: one thread per 8-mode chunk and time segment.
void LaunchResonatorKernel(const BankView &, const RenderPlan &, int steps, cudaStream_t, LaunchCounter &);
uint32_t ResonatorRows(uint32_t n_chunks);
// Tensor-core form, producer side: the same per-chunk walk over RenderModal blocks (culling, increments, final state)
// as the resonator kernel, but advancing one 256-frame time block per step with c^256 and writing the state stages instead of samples.
void LaunchStateWalkKernel(const BankView &, const RenderPlan &, cudaStream_t, LaunchCounter &);
// Power stages of the installed tuning: c^1..c^256 of every mode (FP64 products of the float coefficient), split into
// FP16 hi + lo, in the stage layout of tensor_mix.cuh. powers: [NChunks/256][256 stages][hi, lo][256 x 16] FP16.
void LaunchPowerTableKernel(const BankView &, float *powers, cudaStream_t, LaunchCounter &);
// out[n] = sum of the partial rows in fixed order (the reference sums renderer buffers in a fixed order, :553-555)
// plus the pulse rows overlapping n.
void LaunchMixKernel(const float *partial, uint32_t rows, const RenderPlan &, const PulsePlan &, float *out, cudaStream_t, LaunchCounter &);
// The acceleration-noise click of each click-carrying impact, added into out (ModalAudio.cpp:527-531).
void LaunchClickKernel(const DevImpact *impacts, const DevImpactTail *tails, uint32_t n_impacts, float *out, uint32_t frames, cudaStream_t, LaunchCounter &);

// FP32 issue-rate micro-benchmark; returns lane-operations per second.
// mode 0 scalar FFMA, 1 packed FFMA2, 2/3 FFMA2+FFMA interleaved (1:1, 1:2 instructions), 4 scalar FADD.
double MeasureFmaRate(int mode, int iters);

} // namespace me
