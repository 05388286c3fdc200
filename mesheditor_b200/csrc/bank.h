// Host side of the resonator bank: the reference's ModalBank / ModalAudio bookkeeping (object slots, tuning,
// the SPSC event ring, impact lifetimes) in C++, with the per-mode columns resident in HBM and every sample
// computed by the kernels in resonator.cu. Reference: src/audio/ModalAudio.{h,cpp}.
#pragma once

#include "common.h"
#include "resonator.cuh"
#include "tensor_mix.cuh"

#include <array>
#include <cstring>
#include <atomic>
#include <cstdint>
#include <functional>
#include <vector>

namespace me {

// ModalBank::ActiveImpact (ModalAudio.h:147-160).
struct HostImpact {
    uint32_t Object, ExPos, SamplesLeft;
    float Jx, Jy, Jz;
    float PhaseRe, PhaseIm, RotRe, RotIm;
    float Gamma, AccelAmp;
    float ClickB0, ClickA1, ClickA2, ClickZ1, ClickZ2;
};

// An impact as scheduled inside one span of the timeline (a stretch without Silence events inside).
struct ScheduledImpact {
    HostImpact AtStart; // state at Start
    HostImpact AtEnd;   // state the next span adopts when it survives
    uint32_t Start;     // relative to the span
    uint32_t End;       // frame at which it is retired, or the span's end when it survives
    bool Survives;
};

// The per-span plan (impact records, pulse jobs, per-object increment and excitation lists, gains) goes to the device through
// ONE pinned staging buffer and one device buffer, in at most two asynchronous copies: a copy from pageable memory is staged by
// the driver synchronously, ~10-20 us each, and the plan used to be eleven std::vectors - 0.15 ms of an 8-GPU rank's 1.4 ms step.
struct PlanArena {
    PinnedBuffer<uint8_t> Host;
    DeviceBuffer<uint8_t> Dev;
    size_t Cursor{0}, Flushed{0};
    cudaEvent_t Copied{nullptr}; // the last flush has left the pinned buffer
    ~PlanArena() {
        if (Copied) cudaEventDestroy(Copied);
    }
    // Starts a new plan of at most `capacity` bytes (the pointers Stage returns stay valid until the next Begin).
    void Begin(size_t capacity) {
        if (!Copied) ME_CUDA(cudaEventCreateWithFlags(&Copied, cudaEventDisableTiming));
        else ME_CUDA(cudaEventSynchronize(Copied));
        Host.Reserve(capacity), Dev.Reserve(capacity);
        Cursor = Flushed = 0;
    }
    template<typename T>
    const T *Stage(const T *src, size_t count) {
        const size_t at = (Cursor + 255) & ~size_t(255), bytes = count * sizeof(T);
        if (at + bytes > Host.Capacity) Fail(ME_CUDA_ERROR, "internal: plan arena overflow (%zu + %zu of %zu bytes)", at, bytes, Host.Capacity);
        if (bytes) std::memcpy(Host.Ptr + at, src, bytes);
        Cursor = at + bytes;
        return reinterpret_cast<const T *>(Dev.Ptr + at);
    }
    template<typename T>
    const T *Stage(const std::vector<T> &v) { return Stage(v.data(), v.size()); }
    static size_t Room(size_t bytes) { return bytes + 256; }
    // A region filled in place, piece by piece, and copied with FlushRange (the arrays a span's pulse batches append to). Regions
    // are reserved right after Begin, before anything is staged; SkipRegions then moves the flush cursor past them.
    template<typename T>
    struct Region {
        T *Host;
        const T *Dev;
    };
    template<typename T>
    Region<T> Reserve(size_t count) {
        const size_t at = (Cursor + 255) & ~size_t(255), bytes = count * sizeof(T);
        if (at + bytes > Host.Capacity) Fail(ME_CUDA_ERROR, "internal: plan arena overflow (%zu + %zu of %zu bytes)", at, bytes, Host.Capacity);
        Cursor = at + bytes;
        return {reinterpret_cast<T *>(Host.Ptr + at), reinterpret_cast<const T *>(Dev.Ptr + at)};
    }
    void SkipRegions() { Flushed = Cursor; }
    template<typename T>
    void FlushRange(const Region<T> &region, size_t first, size_t count, cudaStream_t stream) {
        if (count == 0) return;
        const size_t at = size_t(reinterpret_cast<const uint8_t *>(region.Host + first) - Host.Ptr);
        ME_CUDA(cudaMemcpyAsync(Dev.Ptr + at, Host.Ptr + at, count * sizeof(T), cudaMemcpyHostToDevice, stream));
    }
    // Every copy of this plan was issued on `stream` or on streams it has since waited for.
    void Seal(cudaStream_t stream) { ME_CUDA(cudaEventRecord(Copied, stream)); }
    // Everything staged since the last flush, in one copy.
    void Flush(cudaStream_t stream) {
        if (Cursor > Flushed) ME_CUDA(cudaMemcpyAsync(Dev.Ptr + Flushed, Host.Ptr + Flushed, Cursor - Flushed, cudaMemcpyHostToDevice, stream));
        Flushed = Cursor;
        ME_CUDA(cudaEventRecord(Copied, stream));
    }
};

class Bank {
public:
    Bank(float sample_rate, int device);
    ~Bank();
    Bank(const Bank &) = delete;
    Bank &operator=(const Bank &) = delete;

    uint32_t AddObject(uint32_t n_modes, uint32_t n_points, const float *shapes_xyz, const float *positions_xyz, const uint32_t *indices, uint32_t n_indices);
    void TuneObject(uint32_t slot, const float *freqs, const float *t60s, uint32_t n, float radius_scale);
    void SetObjectShapes(uint32_t slot, uint32_t n_modes, uint32_t n_points, const float *shapes_xyz);
    void SetGain(uint32_t slot, float out_gain, float listener_gain);
    void SetOutGain(uint32_t slot, float out_gain); // SetModalOutGain (AudioSystem.cpp:227-230): the listener gain stays
    void SetClickGain(float g) { ClickGain = g; }
    void SetMaxImpacts(uint32_t n) { MaxImpacts = n; }
    void SetTimeSegments(uint32_t n) { RequestedSegments = n; }
    void SetRenderPath(uint32_t path) { RenderPath = path; }

    void Install();
    MeStatus Enqueue(const MeModalEvent &);
    // `out` is a host buffer that is added into when `out_is_device` is false, else a device buffer that is overwritten.
    void RenderTimeline(const MeModalEvent *events, const uint64_t *event_frames, uint32_t n_events, uint64_t total_frames, uint32_t block_frames, float *out, bool out_is_device, cudaStream_t stream, bool use_own_stream);

    uint32_t ObjectCount() const { return uint32_t(ModeOffset.size()); }
    uint32_t ModeTotal() const { return uint32_t(CoeffRe.size()); }
    uint32_t ActiveImpacts() const { return uint32_t(Impacts.size()); }
    uint64_t EventsDroppedCount() const { return EventsDropped; }
    void GetModeColumn(MeModeColumn which, float *out);
    void GetObjectLayout(uint32_t slot, uint32_t *mode_offset, uint32_t *mode_count, uint32_t *tuned, float *radius) const;
    void GetObjectStatus(uint32_t slot, uint32_t *live_mode_count, uint32_t *ringing);
    const MeRenderStats &LastStats();

private:
    void CheckSlot(uint32_t slot) const;
    void RequireInstalled() const;
    void UploadTuning(cudaStream_t);
    void ResetObjectOnDevice(uint32_t object, bool clear_state, cudaStream_t);
    // Upper bounds of what a span's impacts need in the plan and in the pulse buffers (sized before the first impact is admitted).
    struct SpanBounds {
        uint64_t Impacts{0}, Force{0}, Delta{0}, Rows{0}, Warps{0};
    };
    // `admit_before(limit)` appends to the schedule, in start order, every impact of the span that starts before frame `limit`
    // (relative to the span) and has not been admitted yet.
    void RenderSpan(uint32_t frames, uint32_t block_frames, std::vector<ScheduledImpact> &, const std::function<void(uint32_t)> &admit_before, const SpanBounds &, float *out_dev, cudaStream_t);
    cudaEvent_t NextEvent();
    cudaEvent_t NextJoin();
    // Brackets `launch` with an event pair of the given kind.
    template<typename F>
    void Timed(uint8_t kind, cudaStream_t stream, F &&launch) {
        cudaEvent_t a = NextEvent(), b = NextEvent();
        EventKind.push_back(kind);
        ME_CUDA(cudaEventRecord(a, stream));
        launch();
        ME_CUDA(cudaEventRecord(b, stream));
    }
    BankView View() const;

    float SampleRate;
    int Device;
    cudaStream_t OwnStream{nullptr};
    cudaStream_t PulseStream{nullptr}; // tensor-core form: the force + pulse kernels of a sub-window run beside the state walk of the one before
    cudaStream_t ForceStream{nullptr}; // ... and the record copies + force kernel of a batch beside the pulse kernel of the batch before
    std::vector<cudaEvent_t> JoinPool; // untimed events ordering the streams
    uint32_t JoinsUsed{0};
    uint32_t SubWindowTiles{2};        // tiles per sub-window of the pulse / walk pipeline (0: one batch per launch window, on the render stream)
    cudaEvent_t EvBegin{nullptr}, EvEnd{nullptr};
    std::vector<cudaEvent_t> EventPool; // begin/end pairs around every resonator kernel launch of the last call
    std::vector<uint8_t> EventKind;     // per pair: 0 the whole resonator stage of a window, 1 the walk kernel, 2 the tcgen05 mix kernel, 3 force + pulse kernels
    uint32_t EventsUsed{0};
    bool StatsResolved{true};

    // Host columns, objects concatenated exactly like ModalBank (unpadded).
    std::vector<float> CoeffRe, CoeffIm, RadiationGain, RadiationArea, DeflectionGain, OutPhaseIm, OutPhaseRe, QuadCompliance, QuadDriveScale;
    std::vector<float> ShapeX, ShapeY, ShapeZ;
    std::vector<uint32_t> ModeOffset, ModeCount, ShapeOffset, ShapePoints, TunedModeCount;
    std::vector<float> OutGain, ListenerGain, RadiantRadius, DeflectionScale;

    // ModalAudio side.
    float ClickGain{1.f};
    uint32_t MaxImpacts{1024};
    static constexpr uint32_t EventCapacity{256};
    std::array<MeModalEvent, EventCapacity> Events{};
    std::atomic<uint32_t> EventWrite{0}, EventRead{0};
    bool FlushEvents{false};
    uint64_t EventsDropped{0};
    std::vector<HostImpact> Impacts; // in flight between render calls

    // Device residency.
    bool Installed{false};
    bool Installing{false};
    bool NeedsReinstall{false};  // AddObject after Install: the device layout is stale until the next Install
    uint32_t InstalledObjects{0};
    bool TuningDirty{false};
    uint32_t NChunks{0};
    std::vector<uint32_t> ObjFirstChunk, ObjStride, ObjPaddedShapeOffset; // per object, padded layout
    std::vector<uint32_t> RetunedObjects; // LiveModeCount resets owed to the device (TuneModalObject :392)
    int Side{0};                          // which of the ping-pong state buffers holds the current state
    DeviceBuffer<float> DCoeffRe, DCoeffIm, DStateRe[2], DStateIm[2], DPhaseIm, DPhaseRe, DRadiationGain, DShapeX, DShapeY, DShapeZ;
    PlanArena Plan;                                             // the span's plan arrays (see PlanArena)
    const float *PlanMixGain{nullptr}, *PlanEnergyScale{nullptr}; // inside Plan.Dev, for View()
    DeviceBuffer<uint32_t> DChunkObject, DObjShapeOffset, DObjStride, DObjFirstChunk, DObjTunedChunks;
    DeviceBuffer<uint8_t> DObjCull, DChunkLive[2], DObjRinging[2];
    DeviceBuffer<uint32_t> DSpeculation;
    DeviceBuffer<float> DForce, DPartial, DOut, DSegRe, DSegIm, DDeltaRe, DDeltaIm, DPulseRows;
    DeviceBuffer<float> DSnapRe, DSnapIm; // state at the start of a launch window walked in several sub-windows (for the sequential fallback)
    DeviceBuffer<uint8_t> DSnapLive, DSnapRinging;
    DeviceBuffer<double> DLogRho, DTheta;
    PinnedBuffer<float> POut;

    // Per-call scratch.
    std::vector<uint32_t> CallInjectPtr, CallInjectFrame, CallInjectDelta, CallExcitePtr, CallExciteBegin, CallExciteEnd;
    std::vector<float> MixGain, EnergyScale;

    // Tensor-core form (tensor_mix.cuh): power stages of the installed tuning, state stages and group mixes of a window.
    DeviceBuffer<float> DPowers, DWalkStates, DWalkScales, DGroupMix;
    uint64_t StateBudgetBytes{0}; // HBM the state stages of one launch window may take (half of what was free at first use, at most 40 GB)
    bool PowersDirty{true};
    uint64_t TuningVersion{1}, PowersVersion{0}; // the power stages follow the coefficients, not the install
    uint32_t RenderPath{0}; // 0 automatic, 1 sample loop (FP32 pipe), 2 tensor-core form wherever the span allows it

    uint32_t RequestedSegments{0};
    int Steps{4};
    bool SpeculationFailed{false};
    bool SeededWalkFailed{false}; // tensor-core form: a culling decision fell inside a seeded walk; later windows walk sequentially
    MeRenderStats Stats{};
    LaunchCounter Counter;

    friend struct BankAccess;
};

// DealObjects (ModalAudio.cpp:430-461): heaviest first onto the least-loaded renderer; see me_deal_objects.
void DealObjects(const uint64_t *costs, uint32_t n_objects, uint32_t n_renderers, uint32_t *owner, uint32_t *local_slot);

} // namespace me
