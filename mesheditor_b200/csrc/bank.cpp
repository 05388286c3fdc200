// Host side of the resonator bank. See bank.h.
#include "bank.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <numbers>
#include <numeric>
#include <queue>

namespace me {
namespace {
// ModalAudio.h:41-46.
constexpr float AirDensity{1.204f}, SpeedOfSound{343.f}, ListenerDistance{1.f};
constexpr float Ln1000 = 3 * std::numbers::ln10_v<float>;
constexpr float Pi = std::numbers::pi_v<float>;

constexpr size_t PartialBudgetBytes = size_t(8) << 30; // per-warp partial mixes kept in HBM per launch window (of 180 GB)
// Tensor-core form: a tile is 128 time blocks of 256 frames; the state stages of a launch window stay under this budget.
constexpr uint32_t TensorBlocksPerTile = 128;
constexpr uint32_t TensorTileFrames = TensorBlocksPerTile * kTmBlock;
constexpr size_t TensorStateBudgetBytes = size_t(40) << 30; // 10 s of 1024 x 500 modes is 31.5 GB: one window (of 180 GB)

uint32_t PaddedModes(uint32_t count) { return (count + kLanes - 1) / kLanes * kLanes; }

struct Vec3 {
    float x, y, z;
};
Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
Vec3 Cross(Vec3 a, Vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
float Dot(Vec3 a, Vec3 b) {
    const float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    return px + py + pz;
}
Vec3 At(const float *xyz, size_t i) { return {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}; }

bool HasClick(const HostImpact &im) { return im.AccelAmp != 0.f && im.ClickB0 != 0.f; }

// One sample of the contact pulse and its click filter, in the reference's float operation order (ModalAudio.cpp:518-531).
inline void StepImpact(HostImpact &im) {
    float cur = 0.f;
    if (im.SamplesLeft > 0) {
        const float re = im.PhaseRe * im.RotRe - im.PhaseIm * im.RotIm;
        im.PhaseIm = im.PhaseRe * im.RotIm + im.PhaseIm * im.RotRe;
        im.PhaseRe = re;
        cur = im.Gamma * 0.5f * (1.f - im.PhaseRe);
        --im.SamplesLeft;
    }
    const float u = im.AccelAmp * cur;
    const float y = im.ClickB0 * u + im.ClickZ1;
    im.ClickZ1 = -im.ClickA1 * y + im.ClickZ2;
    im.ClickZ2 = -im.ClickB0 * u - im.ClickA2 * y;
}
} // namespace

Bank::Bank(float sample_rate, int device) : SampleRate(sample_rate), Device(device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) Fail(ME_CUDA_ERROR, "no CUDA device is available (there is no CPU fallback)");
    if (device < 0 || device >= count) Fail(ME_BAD_ARG, "device %d out of range (%d visible)", device, count);
    ME_CUDA(cudaSetDevice(Device));
    // The render stream outranks the pulse stream: when both have CTAs pending, the state walk (which the tcgen05 mix waits for) is
    // placed first and the pulse kernels of the next sub-window fill what is left of the SMs. A caller's stream has the default
    // (lowest) priority, so renders always run on the bank's own stream, forked from the caller's and joined back into it.
    int least = 0, greatest = 0;
    ME_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    ME_CUDA(cudaStreamCreateWithPriority(&OwnStream, cudaStreamNonBlocking, greatest));
    ME_CUDA(cudaStreamCreateWithPriority(&PulseStream, cudaStreamNonBlocking, least));
    ME_CUDA(cudaStreamCreateWithPriority(&ForceStream, cudaStreamNonBlocking, greatest));
    ME_CUDA(cudaEventCreate(&EvBegin));
    ME_CUDA(cudaEventCreate(&EvEnd));
    // Diagnostic overrides: samples advanced per state jump (1, 2 or 4) and segments of the scan along time.
    if (const char *steps = std::getenv("ME_RESONATOR_STEPS")) Steps = std::atoi(steps);
    if (const char *segments = std::getenv("ME_RESONATOR_SEGMENTS")) RequestedSegments = uint32_t(std::atoi(segments));
    if (const char *tiles = std::getenv("ME_WALK_SUBWINDOW_TILES")) SubWindowTiles = uint32_t(std::max(0, std::atoi(tiles)));
}

Bank::~Bank() {
    cudaSetDevice(Device);
    if (PulseStream) cudaStreamSynchronize(PulseStream), cudaStreamDestroy(PulseStream);
    if (ForceStream) cudaStreamSynchronize(ForceStream), cudaStreamDestroy(ForceStream);
    if (OwnStream) cudaStreamSynchronize(OwnStream), cudaStreamDestroy(OwnStream);
    for (auto e : {EvBegin, EvEnd})
        if (e) cudaEventDestroy(e);
    for (auto e : EventPool) cudaEventDestroy(e);
    for (auto e : JoinPool) cudaEventDestroy(e);
}

void Bank::CheckSlot(uint32_t slot) const {
    if (slot >= ObjectCount()) Fail(ME_BAD_ARG, "object slot %u out of range (%u objects)", slot, ObjectCount());
}
void Bank::RequireInstalled() const {
    if (!Installed) Fail(ME_BAD_ARG, "the bank has not been installed (me_bank_install)");
    if (NeedsReinstall) Fail(ME_BAD_ARG, "objects were added after me_bank_install: install the bank again before rendering (%u slots, %u installed)", ObjectCount(), InstalledObjects);
}

// AddModalObject, ModalAudio.cpp:291-338.
uint32_t Bank::AddObject(uint32_t count, uint32_t n_points, const float *shapes, const float *positions, const uint32_t *indices, uint32_t n_indices) {
    if ((count && n_points && !shapes) || (n_indices && (!indices || !positions))) Fail(ME_BAD_ARG, "null shape/position/index array");
    for (uint32_t t = 0; t < n_indices; ++t)
        if (indices[t] >= n_points) Fail(ME_BAD_ARG, "triangle index %u out of range (%u points)", indices[t], n_points);
    const auto slot = ObjectCount();
    const size_t k0 = CoeffRe.size();
    ++TuningVersion;
    // The reference never grows a live bank: RebuildModalBank builds the next one and InstallModalBank swaps it in
    // (ModalAudio.cpp:277-289). A slot added after me_bank_install changes the padded layout the device buffers were sized
    // for, so rendering is refused until the bank is installed again.
    if (Installed) NeedsReinstall = true;
    ModeOffset.push_back(uint32_t(k0));
    ModeCount.push_back(count);
    TunedModeCount.push_back(count);
    ShapeOffset.push_back(uint32_t(ShapeX.size()));
    ShapePoints.push_back(n_points);
    OutGain.push_back(0.f);
    ListenerGain.push_back(1.f);
    DeflectionScale.push_back(1.f);
    for (auto *col : {&CoeffRe, &CoeffIm, &RadiationGain, &DeflectionGain, &QuadCompliance, &QuadDriveScale}) col->resize(k0 + count, 0.f);
    OutPhaseIm.resize(k0 + count, 1.f);
    OutPhaseRe.resize(k0 + count, 0.f);
    const size_t n_shapes = size_t(n_points) * count;
    for (size_t i = 0; i < n_shapes; ++i) {
        ShapeX.push_back(shapes[3 * i]);
        ShapeY.push_back(shapes[3 * i + 1]);
        ShapeZ.push_back(shapes[3 * i + 2]);
    }
    // Radiating strength: surface integral of the squared normal shape by centroid quadrature.
    RadiationArea.resize(k0 + count, 0.f);
    float total_area = 0.f;
    for (size_t t = 0; t + 2 < n_indices; t += 3) {
        const auto i = indices[t], j = indices[t + 1], l = indices[t + 2];
        const Vec3 cr = Cross(At(positions, j) - At(positions, i), At(positions, l) - At(positions, i));
        const float doubled = std::sqrt(Dot(cr, cr));
        if (doubled <= 0.f) continue;
        const Vec3 n{cr.x / doubled, cr.y / doubled, cr.z / doubled};
        const float area = doubled / 2;
        total_area += area;
        for (uint32_t k = 0; k < count; ++k) {
            const Vec3 a = At(shapes, size_t(i) * count + k), b = At(shapes, size_t(j) * count + k), c = At(shapes, size_t(l) * count + k);
            const Vec3 shape{(a.x + b.x + c.x) / 3.f, (a.y + b.y + c.y) / 3.f, (a.z + b.z + c.z) / 3.f};
            const float normal = Dot(shape, n);
            RadiationArea[k0 + k] += area * normal * normal;
        }
    }
    RadiantRadius.push_back(std::sqrt(total_area / (4 * Pi)));
    return slot;
}

// TuneModalObject, ModalAudio.cpp:340-393. Runs on the host with the same libm calls as the reference, so the
// coefficients that reach HBM are the reference's, bit for bit.
void Bank::TuneObject(uint32_t object, const float *freqs, const float *t60s, uint32_t n, float radius_scale) {
    CheckSlot(object);
    if (n && (!freqs || !t60s)) Fail(ME_BAD_ARG, "null frequency/T60 array");
    ++TuningVersion;
    const auto k0 = ModeOffset[object];
    const auto count = std::min(ModeCount[object], n);
    const float sr = SampleRate;
    const float radius = RadiantRadius[object] * radius_scale;
    DeflectionScale[object] = 1.f / (radius_scale * radius_scale * radius_scale);
    for (uint32_t k = 0; k < count; ++k) {
        const float freq = freqs[k], t60 = t60s[k];
        const size_t m = size_t(k0) + k;
        if (!std::isfinite(freq) || !std::isfinite(t60) || freq <= 0.f || freq >= sr / 2 - 1 || t60 <= 0.f) {
            CoeffRe[m] = CoeffIm[m] = RadiationGain[m] = DeflectionGain[m] = QuadCompliance[m] = QuadDriveScale[m] = 0.f;
            OutPhaseIm[m] = 1.f;
            OutPhaseRe[m] = 0.f;
            continue;
        }
        const auto omega = 2 * Pi * freq / sr;
        const float omega_si = 2 * Pi * freq;
        const float ka = omega_si * radius / SpeedOfSound;
        const float sigma = ka * ka / (1 + ka * ka);
        const float area = RadiationArea[m] / radius_scale;
        const float radiation_rate = AirDensity * SpeedOfSound * sigma * area * 0.5f;
        const auto decay = std::exp(-(Ln1000 / t60 + radiation_rate) / sr);
        CoeffRe[m] = decay * std::cos(omega);
        CoeffIm[m] = decay * std::sin(omega);
        const float gain = AirDensity * SpeedOfSound * std::sqrt(sigma * RadiationArea[m] / (4 * Pi)) / ListenerDistance;
        RadiationGain[m] = gain;
        const float spread = sigma * Pi * (2.f * std::fmod(0.6180339887f * float(k + 1), 1.0f) - 1.f);
        OutPhaseIm[m] = std::cos(spread);
        OutPhaseRe[m] = std::sin(spread);
        DeflectionGain[m] = gain > 0.f ? 1.f / (gain * omega_si) : 0.f;
        const float dt = 1.f / sr;
        const float central = dt * (1 + decay * decay + 2 * decay * std::cos(omega)) / 4;
        QuadCompliance[m] = central;
        QuadDriveScale[m] = central * omega_si / (decay * std::sin(omega));
    }
    uint32_t live = ModeCount[object];
    while (live > 0 && CoeffRe[k0 + live - 1] == 0.f && CoeffIm[k0 + live - 1] == 0.f) --live;
    TunedModeCount[object] = live;
    if (Installed) TuningDirty = true, RetunedObjects.push_back(object);
}

// SetModalObjectShapes, ModalAudio.cpp:395-410.
void Bank::SetObjectShapes(uint32_t object, uint32_t n_modes, uint32_t n_points, const float *shapes) {
    CheckSlot(object);
    if (ModeCount[object] != n_modes || ShapePoints[object] != n_points) Fail(ME_BAD_ARG, "shape layout differs from the object's (%u modes x %u points)", ModeCount[object], ShapePoints[object]);
    const size_t begin = ShapeOffset[object], n = size_t(n_modes) * n_points;
    for (size_t i = 0; i < n; ++i) {
        ShapeX[begin + i] = shapes[3 * i];
        ShapeY[begin + i] = shapes[3 * i + 1];
        ShapeZ[begin + i] = shapes[3 * i + 2];
    }
    if (Installed) TuningDirty = true;
}

void Bank::SetGain(uint32_t slot, float out_gain, float listener_gain) {
    CheckSlot(slot);
    OutGain[slot] = out_gain;
    ListenerGain[slot] = listener_gain;
}

void Bank::SetOutGain(uint32_t slot, float out_gain) {
    CheckSlot(slot);
    OutGain[slot] = out_gain;
}

// Pads every object to whole chunks, places objects so that none of at most 256 chunks straddles a CTA, and uploads
// the per-mode columns. Also used for live retunes.
void Bank::UploadTuning(cudaStream_t stream) {
    const uint32_t n_obj = ObjectCount();
    ObjFirstChunk.assign(n_obj, 0);
    ObjStride.assign(n_obj, 0);
    ObjPaddedShapeOffset.assign(n_obj, 0);
    std::vector<uint32_t> tuned_chunks(n_obj);
    std::vector<uint8_t> obj_cull(n_obj);
    uint64_t cursor = 0;
    size_t shape_total = 0;
    for (uint32_t o = 0; o < n_obj; ++o) {
        ObjStride[o] = PaddedModes(ModeCount[o]);
        const uint32_t chunks = ObjStride[o] / kLanes;
        const bool fits = chunks <= kBlockThreads;
        if (!fits || cursor % kBlockThreads + chunks > kBlockThreads) cursor = (cursor + kBlockThreads - 1) / kBlockThreads * kBlockThreads;
        ObjFirstChunk[o] = uint32_t(cursor);
        cursor += chunks;
        obj_cull[o] = fits;
        tuned_chunks[o] = (TunedModeCount[o] + kLanes - 1) / kLanes;
        ObjPaddedShapeOffset[o] = uint32_t(shape_total);
        shape_total += size_t(ObjStride[o]) * ShapePoints[o];
    }
    cursor = (cursor + kBlockThreads - 1) / kBlockThreads * kBlockThreads;
    if (cursor * kLanes >= (uint64_t(1) << 31) || shape_total + kLanes >= (size_t(1) << 32)) Fail(ME_BAD_ARG, "bank too large for 32-bit mode indexing");
    const uint32_t chunks = uint32_t(cursor);
    const size_t padded = size_t(chunks) * kLanes;
    const bool layout_changed = chunks != NChunks;
    NChunks = chunks;

    std::vector<float> cre(padded, 0.f), cim(padded, 0.f), pim(padded, 1.f), pre(padded, 0.f), rg(padded, 0.f);
    std::vector<float> sx(shape_total, 0.f), sy(shape_total, 0.f), sz(shape_total, 0.f);
    std::vector<uint32_t> chunk_obj(chunks, kNoObject);
    for (uint32_t o = 0; o < n_obj; ++o) {
        const size_t dst = size_t(ObjFirstChunk[o]) * kLanes, src = ModeOffset[o];
        // Modes past TunedModeCount are muted (coefficient 0 already), so the whole ModeCount is copied.
        std::copy_n(CoeffRe.begin() + src, ModeCount[o], cre.begin() + dst);
        std::copy_n(CoeffIm.begin() + src, ModeCount[o], cim.begin() + dst);
        std::copy_n(OutPhaseIm.begin() + src, ModeCount[o], pim.begin() + dst);
        std::copy_n(OutPhaseRe.begin() + src, ModeCount[o], pre.begin() + dst);
        std::copy_n(RadiationGain.begin() + src, ModeCount[o], rg.begin() + dst);
        for (uint32_t p = 0; p < ShapePoints[o]; ++p) {
            const size_t s_src = ShapeOffset[o] + size_t(p) * ModeCount[o], s_dst = ObjPaddedShapeOffset[o] + size_t(p) * ObjStride[o];
            std::copy_n(ShapeX.begin() + s_src, ModeCount[o], sx.begin() + s_dst);
            std::copy_n(ShapeY.begin() + s_src, ModeCount[o], sy.begin() + s_dst);
            std::copy_n(ShapeZ.begin() + s_src, ModeCount[o], sz.begin() + s_dst);
        }
        for (uint32_t c = 0; c < ObjStride[o] / kLanes; ++c) chunk_obj[ObjFirstChunk[o] + c] = o;
    }
    // Polar form of every coefficient in FP64, for the powers c^m the scan along time needs.
    std::vector<double> log_rho(padded), theta(padded);
    for (size_t m = 0; m < padded; ++m) {
        const double re = cre[m], im = cim[m];
        const double rho = std::hypot(re, im);
        log_rho[m] = rho > 0 ? std::log(rho) : -std::numeric_limits<double>::infinity();
        theta[m] = rho > 0 ? std::atan2(im, re) : 0.0;
    }
    // A live retune (TuneModalObject on an installed bank) never changes the slot layout; only Install() may.
    if (Installed && !Installing && layout_changed) Fail(ME_BAD_ARG, "the padded layout changed under an installed bank (%u chunks); install the bank again", chunks);
    DCoeffRe.Upload(cre, stream), DCoeffIm.Upload(cim, stream), DPhaseIm.Upload(pim, stream), DPhaseRe.Upload(pre, stream), DRadiationGain.Upload(rg, stream);
    DShapeX.Upload(sx, stream), DShapeY.Upload(sy, stream), DShapeZ.Upload(sz, stream);
    DChunkObject.Upload(chunk_obj, stream);
    DObjShapeOffset.Upload(ObjPaddedShapeOffset, stream), DObjStride.Upload(ObjStride, stream), DObjFirstChunk.Upload(ObjFirstChunk, stream);
    DObjTunedChunks.Upload(tuned_chunks, stream), DObjCull.Upload(obj_cull, stream);
    DLogRho.Upload(log_rho, stream), DTheta.Upload(theta, stream);
    // The uploads read pageable host vectors that die at the end of this scope.
    ME_CUDA(cudaStreamSynchronize(stream));
    TuningDirty = false;
}

// InstallModalBank, ModalAudio.cpp:277-289: a freshly built bank starts silent, and events queued against the old
// slot layout are dropped by the next render.
void Bank::Install() {
    ME_CUDA(cudaSetDevice(Device));
    Installing = true;
    try {
        UploadTuning(OwnStream);
    } catch (...) {
        Installing = false;
        throw;
    }
    Installing = false;
    const size_t padded = std::max<size_t>(size_t(NChunks) * kLanes, 1);
    const uint32_t n_obj = ObjectCount();
    for (int side = 0; side < 2; ++side) {
        DStateRe[side].Reserve(padded), DStateIm[side].Reserve(padded);
        DChunkLive[side].Reserve(std::max<uint32_t>(NChunks, 1)), DObjRinging[side].Reserve(std::max<uint32_t>(n_obj, 1));
    }
    Side = 0;
    ME_CUDA(cudaMemsetAsync(DStateRe[0].Ptr, 0, padded * sizeof(float), OwnStream));
    ME_CUDA(cudaMemsetAsync(DStateIm[0].Ptr, 0, padded * sizeof(float), OwnStream));
    ME_CUDA(cudaMemsetAsync(DChunkLive[0].Ptr, 1, std::max<uint32_t>(NChunks, 1), OwnStream));
    ME_CUDA(cudaMemsetAsync(DObjRinging[0].Ptr, 0, std::max<uint32_t>(n_obj, 1), OwnStream));
    DSpeculation.Reserve(64); // one word per sub-window of a launch window
    ME_CUDA(cudaStreamSynchronize(OwnStream));
    Impacts.clear();
    RetunedObjects.clear();
    FlushEvents = true;
    Installed = true;
    NeedsReinstall = false;
    InstalledObjects = n_obj;
}

// EnqueueModalEvent, ModalAudio.cpp:417-425.
MeStatus Bank::Enqueue(const MeModalEvent &e) {
    const auto write = EventWrite.load(std::memory_order_relaxed);
    if (write - EventRead.load(std::memory_order_acquire) >= EventCapacity) {
        ++EventsDropped;
        SetLastError("event queue full (%u entries); event dropped", EventCapacity);
        return ME_QUEUE_FULL;
    }
    Events[write % EventCapacity] = e;
    EventWrite.store(write + 1, std::memory_order_release);
    return ME_OK;
}

BankView Bank::View() const {
    return {
        .NChunks = NChunks,
        .NObjects = ObjectCount(),
        .CoeffRe = DCoeffRe.Ptr,
        .CoeffIm = DCoeffIm.Ptr,
        .StateRe = DStateRe[Side].Ptr,
        .StateIm = DStateIm[Side].Ptr,
        .StateOutRe = DStateRe[Side ^ 1].Ptr,
        .StateOutIm = DStateIm[Side ^ 1].Ptr,
        .PhaseIm = DPhaseIm.Ptr,
        .PhaseRe = DPhaseRe.Ptr,
        .RadiationGain = DRadiationGain.Ptr,
        .LogRho = DLogRho.Ptr,
        .Theta = DTheta.Ptr,
        .ChunkObject = DChunkObject.Ptr,
        .ShapeX = DShapeX.Ptr,
        .ShapeY = DShapeY.Ptr,
        .ShapeZ = DShapeZ.Ptr,
        .ObjShapeOffset = DObjShapeOffset.Ptr,
        .ObjStride = DObjStride.Ptr,
        .ObjFirstChunk = DObjFirstChunk.Ptr,
        .ObjTunedChunks = DObjTunedChunks.Ptr,
        .ObjMixGain = PlanMixGain,
        .ObjEnergyScale = PlanEnergyScale,
        .ObjCull = DObjCull.Ptr,
        .ChunkLive = DChunkLive[Side].Ptr,
        .ObjRinging = DObjRinging[Side].Ptr,
        .ChunkLiveOut = DChunkLive[Side ^ 1].Ptr,
        .ObjRingingOut = DObjRinging[Side ^ 1].Ptr,
    };
}

// The device half of SilenceObject (ModalAudio.cpp:53-64) and of TuneModalObject's LiveModeCount reset (:392).
void Bank::ResetObjectOnDevice(uint32_t o, bool clear_state, cudaStream_t stream) {
    const size_t first = size_t(ObjFirstChunk[o]) * kLanes, count = size_t(ObjStride[o]);
    if (clear_state) {
        ME_CUDA(cudaMemsetAsync(DStateRe[Side].Ptr + first, 0, count * sizeof(float), stream));
        ME_CUDA(cudaMemsetAsync(DStateIm[Side].Ptr + first, 0, count * sizeof(float), stream));
        ME_CUDA(cudaMemsetAsync(DObjRinging[Side].Ptr + o, 0, 1, stream));
    }
    if (count) ME_CUDA(cudaMemsetAsync(DChunkLive[Side].Ptr + ObjFirstChunk[o], 1, count / kLanes, stream));
}

namespace {
// End of the render block holding `frame` (frames relative to the span, which starts on a block boundary).
uint32_t BlockEnd(uint32_t frame, uint32_t block_frames, uint32_t span_frames) {
    const uint64_t end = (uint64_t(frame) / block_frames + 1) * block_frames;
    return uint32_t(std::min<uint64_t>(end, span_frames));
}

// Walks one impact from its start to its retirement or the end of the span (RenderModal :506-538, :557-561).
// `s.AtEnd` becomes the state the next span adopts when it survives.
void Schedule(ScheduledImpact &s, uint32_t block_frames, uint32_t span_frames) {
    s.AtEnd = s.AtStart;
    if (!HasClick(s.AtStart)) {
        // The click filter stays at rest, so the impact is retired at the end of the block its pulse ends in.
        const uint64_t pulse_end = uint64_t(s.Start) + s.AtStart.SamplesLeft;
        if (pulse_end <= span_frames) {
            s.End = BlockEnd(s.AtStart.SamplesLeft ? uint32_t(pulse_end - 1) : s.Start, block_frames, span_frames);
            s.Survives = false;
        } else {
            for (uint32_t f = s.Start; f < span_frames; ++f) StepImpact(s.AtEnd);
            s.End = span_frames;
            s.Survives = true;
        }
        return;
    }
    uint32_t f = s.Start;
    while (f < span_frames) {
        const uint32_t end = BlockEnd(f, block_frames, span_frames);
        for (; f < end; ++f) StepImpact(s.AtEnd);
        if (s.AtEnd.SamplesLeft == 0 && std::abs(s.AtEnd.ClickZ1) + std::abs(s.AtEnd.ClickZ2) < 1e-12f) {
            s.End = end;
            s.Survives = false;
            return;
        }
    }
    s.End = span_frames;
    s.Survives = true;
}

// Samples of a contact pulse (ActivateImpact, ModalAudio.cpp:36).
uint32_t PulseSamples(float pulse_step) {
    const float steps = std::ceil(1.f / pulse_step);
    return steps >= 4294967040.f ? 4294967040u : uint32_t(steps);
}

// ActivateImpact, ModalAudio.cpp:28-51.
HostImpact MakeImpact(const MeModalEvent &e) {
    const auto theta = 2 * Pi * e.pulse_step;
    return {
        .Object = e.object,
        .ExPos = e.ex_pos,
        .SamplesLeft = PulseSamples(e.pulse_step),
        .Jx = e.jx,
        .Jy = e.jy,
        .Jz = e.jz,
        .PhaseRe = 1.f,
        .PhaseIm = 0.f,
        .RotRe = std::cos(theta),
        .RotIm = std::sin(theta),
        .Gamma = e.pulse_gamma,
        .AccelAmp = e.accel_amp,
        .ClickB0 = e.click_b0,
        .ClickA1 = e.click_a1,
        .ClickA2 = e.click_a2,
        .ClickZ1 = 0.f,
        .ClickZ2 = 0.f,
    };
}
} // namespace

void Bank::RenderSpan(uint32_t frames, uint32_t block_frames, std::vector<ScheduledImpact> &scheduled, const std::function<void(uint32_t)> &admit_before, const SpanBounds &bounds, float *out_dev, cudaStream_t stream) {
    using Clock = std::chrono::steady_clock;
    const auto since = [](Clock::time_point t0) { return std::chrono::duration<float, std::milli>(Clock::now() - t0).count(); };
    const uint32_t n_obj = ObjectCount();
    // The tensor-core form needs RenderModal blocks made of whole 256-frame time blocks and tiles made of whole
    // RenderModal blocks, and pays off once (chunk groups x tiles) fills the SMs.
    const uint32_t groups = NChunks / kTmGroupChunks;
    const bool tensor_possible = groups > 0 && block_frames % kTmBlock == 0 && TensorTileFrames % block_frames == 0 && !SpeculationFailed;
    const uint64_t tensor_units = uint64_t(groups) * ((frames + TensorTileFrames - 1) / TensorTileFrames);
    const bool tensor_span = tensor_possible && (RenderPath == 2 || (RenderPath == 0 && tensor_units >= 128 && frames >= 2 * TensorTileFrames));
    // Tensor spans are rendered as a pipeline over sub-windows of a few tiles: while the walk kernel (HBM-write bound) steps the
    // bank through sub-window k on the render stream, the force + pulse kernels (FP32-issue bound) of sub-window k + 1 run beside
    // it on the pulse stream, and the host admits and plans sub-window k + 2. Only the first sub-window's planning is exposed.
    const bool piped = tensor_span && SubWindowTiles > 0;
    static const bool small_banks_piped = std::getenv("ME_SMALL_BANKS_PIPED") != nullptr; // measured: no gain (the chain pulses -> walk -> mix of a 128-voice bank is serial either way)
    const bool small_bank = groups < 128;
    const cudaStream_t pulse_stream = piped ? PulseStream : stream;

    auto host_begin = Clock::now();
    const uint64_t window_bound = tensor_span ? (uint64_t(frames) + TensorTileFrames - 1) / TensorTileFrames + 1 : 2; // list sets staged at most
    Plan.Begin(PlanArena::Room(bounds.Impacts * sizeof(DevImpact)) + PlanArena::Room(bounds.Impacts * sizeof(DevImpactTail)) + PlanArena::Room(bounds.Warps * sizeof(PulseWarp)) + 2 * PlanArena::Room(n_obj * sizeof(float)) +
               window_bound * (2 * PlanArena::Room((size_t(n_obj) + 1) * 4) + 4 * PlanArena::Room(size_t(bounds.Impacts) * 4)) + 4096);
    const auto impacts = Plan.Reserve<DevImpact>(bounds.Impacts);
    const auto tails = Plan.Reserve<DevImpactTail>(bounds.Impacts);
    const auto pulse_warps = Plan.Reserve<PulseWarp>(bounds.Warps);
    Plan.SkipRegions();
    DForce.Reserve(std::max<uint64_t>(bounds.Force, 1)), DDeltaRe.Reserve(std::max<uint64_t>(bounds.Delta, 1)), DDeltaIm.Reserve(std::max<uint64_t>(bounds.Delta, 1)), DPulseRows.Reserve(std::max<uint64_t>(bounds.Rows, 1));
    // Slots of the delta buffers the pulse kernel does not write (chunks past the object in a warp's tail are skipped,
    // every chunk inside the object is written) need no clearing.
    {
        // The mix gains are read by the pulse kernel too (and the view below must see their final addresses).
        MixGain.resize(n_obj), EnergyScale.resize(n_obj);
        for (uint32_t o = 0; o < n_obj; ++o) {
            MixGain[o] = OutGain[o] * ListenerGain[o];
            EnergyScale[o] = MixGain[o] != 0.f ? (OutGain[o] * OutGain[o]) / (MixGain[o] * MixGain[o]) : 0.f;
        }
        PlanMixGain = Plan.Stage(MixGain), PlanEnergyScale = Plan.Stage(EnergyScale);
        Plan.Flush(stream);
        Stats.h2d_bytes += 2 * n_obj * sizeof(float);
    }
    const cudaStream_t force_stream = piped ? ForceStream : stream;
    if (piped) {
        // The pulse and force streams pick up behind everything the render stream has been given so far (the gains, earlier spans).
        const cudaEvent_t fork = NextJoin();
        ME_CUDA(cudaEventRecord(fork, stream));
        ME_CUDA(cudaStreamWaitEvent(pulse_stream, fork, 0));
        ME_CUDA(cudaStreamWaitEvent(force_stream, fork, 0));
    }
    BankView view = View();

    // ---- pulse batches: impacts in start order, planned and launched as they are admitted -------------------------------
    uint32_t n = 0, n_warps = 0, max_len = 0;
    uint64_t force_total = 0, delta_total = 0, row_total = 0;
    bool any_click = false;
    cudaEvent_t pulses_done = nullptr; // last batch's kernels have finished (pulse stream)
    std::vector<uint32_t> order;
    const auto current_pulses = [&] {
        return PulsePlan{.NPulseWarps = n_warps, .Warps = pulse_warps.Dev, .Impacts = impacts.Dev, .Force = DForce.Ptr, .Rows = DPulseRows.Ptr, .DeltaRe = DDeltaRe.Ptr, .DeltaIm = DDeltaIm.Ptr, .MaxLen = max_len};
    };
    const auto plan_pulses_before = [&](uint32_t limit) {
        admit_before(limit);
        const uint32_t n0 = n, n1 = uint32_t(scheduled.size()), w0 = n_warps;
        if (n1 == n0) return;
        if (n1 > bounds.Impacts) Fail(ME_CUDA_ERROR, "internal: %u impacts admitted, %llu planned for", n1, (unsigned long long)bounds.Impacts);
        // Sorted by (start, object): the order of the pulse rows in the mix. Batches are disjoint in start, so each is sorted on its own.
        order.resize(n1 - n0);
        std::iota(order.begin(), order.end(), n0);
        const auto by_start_then_object = [&](uint32_t a, uint32_t b) {
            const auto &x = scheduled[a], &y = scheduled[b];
            return x.Start != y.Start ? x.Start < y.Start : x.AtStart.Object < y.AtStart.Object;
        };
        if (!std::is_sorted(order.begin(), order.end(), by_start_then_object)) std::stable_sort(order.begin(), order.end(), by_start_then_object);
        for (uint32_t i = n0; i < n1; ++i) {
            const auto &s = scheduled[order[i - n0]];
            const auto &im = s.AtStart;
            const uint32_t len = uint32_t(std::min<uint64_t>(im.SamplesLeft, frames - s.Start));
            // In a tensor span the pulse kernel keeps rendering the pulse's free ringing up to the next time-block boundary,
            // where its state increment joins the bank (increments must land on block-start states).
            uint32_t render_len = len;
            if (tensor_span && len) render_len = uint32_t(std::min<uint64_t>((uint64_t(s.Start) + len + kTmBlock - 1) / kTmBlock * kTmBlock, frames)) - s.Start;
            const uint32_t stride = ObjStride[im.Object];
            const uint32_t warps = len ? (stride / kLanes + 31) / 32 : 0;
            if (force_total + len >= (uint64_t(1) << 32) || delta_total + stride >= (uint64_t(1) << 32) || row_total + uint64_t(warps) * render_len >= (uint64_t(1) << 32))
                Fail(ME_BAD_ARG, "impact buffers exceed 2^32 entries in one span");
            if (force_total + len > DForce.Capacity || delta_total + (len ? stride : 0) > DDeltaRe.Capacity || row_total + uint64_t(warps) * render_len > DPulseRows.Capacity || n_warps + warps > bounds.Warps)
                Fail(ME_CUDA_ERROR, "internal: impact %u outgrows the span's pulse buffers", i);
            impacts.Host[i] = {.Start = s.Start, .Len = len, .ForceOff = uint32_t(force_total), .ExPos = im.ExPos, .Jx = im.Jx, .Jy = im.Jy, .Jz = im.Jz, .Object = im.Object, .PhaseRe = im.PhaseRe, .PhaseIm = im.PhaseIm, .RotRe = im.RotRe, .RotIm = im.RotIm, .End = s.End, .DeltaOff = uint32_t(delta_total), .HasClick = HasClick(im) ? 1u : 0u, .RenderLen = render_len};
            tails.Host[i] = {.Gamma = im.Gamma, .AccelAmp = im.AccelAmp, .ClickB0 = im.ClickB0, .ClickA1 = im.ClickA1, .ClickA2 = im.ClickA2, .ClickZ1 = im.ClickZ1, .ClickZ2 = im.ClickZ2, .ClickGain = ClickGain * ListenerGain[im.Object]};
            for (uint32_t w = 0; w < warps; ++w) {
                pulse_warps.Host[n_warps++] = {.Impact = i, .Chunk0 = w * 32, .RowOff = uint32_t(row_total), .Start = s.Start, .RenderLen = render_len, .Pad = 0};
                row_total += render_len;
            }
            any_click |= HasClick(im);
            max_len = std::max(max_len, render_len);
            force_total += len;
            if (len) delta_total += stride;
        }
        n = n1;
        // The batch's records and its force curves (one thread per impact walking its rotor sample by sample: ~30 us of latency,
        // a handful of CTAs) go through the force stream, so that they run beside the pulse kernel of the batch before instead
        // of behind it.
        Plan.FlushRange(impacts, n0, n1 - n0, force_stream), Plan.FlushRange(tails, n0, n1 - n0, force_stream), Plan.FlushRange(pulse_warps, w0, n_warps - w0, force_stream);
        Stats.h2d_bytes += (n1 - n0) * (sizeof(DevImpact) + sizeof(DevImpactTail)) + (n_warps - w0) * sizeof(PulseWarp);
        LaunchForceKernel(impacts.Dev + n0, tails.Dev + n0, n1 - n0, DForce.Ptr, force_stream, Counter);
        if (force_stream != pulse_stream) {
            const cudaEvent_t forces_done = NextJoin();
            ME_CUDA(cudaEventRecord(forces_done, force_stream));
            ME_CUDA(cudaStreamWaitEvent(pulse_stream, forces_done, 0));
        }
        Timed(3, pulse_stream, [&] {
            PulsePlan batch = current_pulses();
            batch.NPulseWarps = n_warps - w0, batch.Warps = pulse_warps.Dev + w0;
            LaunchPulseKernel(view, batch, pulse_stream, Counter);
        });
        if (piped) {
            pulses_done = NextJoin();
            ME_CUDA(cudaEventRecord(pulses_done, pulse_stream));
        }
    };
    // The render stream may read what the pulse batches planned so far have written.
    const auto join_pulses = [&] {
        if (pulses_done) ME_CUDA(cudaStreamWaitEvent(stream, pulses_done, 0));
        pulses_done = nullptr;
    };

    // ---- per-object lists of a stretch [begin, end) of the span ------------------------------------------------------------
    // Per object: increments in landing order, and the merged intervals during which it holds a live impact. `candidates` (impact
    // indices, ascending = start order) must hold every impact whose increment lands in [begin, end] or whose lifetime meets
    // [begin, end); a counting sort by object leaves each object's lists in start order.
    struct Lists {
        const uint32_t *InjectPtr, *InjectFrame, *InjectDelta, *ExcitePtr, *ExciteBegin, *ExciteEnd;
    };
    std::vector<std::pair<uint32_t, uint32_t>> excite;
    std::vector<uint32_t> inject_at, excite_at, merged_ptr;
    const auto build_lists = [&](uint32_t begin, uint32_t end, const std::vector<uint32_t> &candidates) {
        const auto lands_here = [&](const DevImpact &im) { return im.Len && im.Start + im.RenderLen >= begin && im.Start + im.RenderLen <= end; };
        const auto lives_here = [&](const DevImpact &im) { return im.End > im.Start && im.End > begin && im.Start < end; };
        CallInjectPtr.assign(n_obj + 1, 0), CallExcitePtr.assign(n_obj + 1, 0);
        for (const uint32_t i : candidates) {
            const auto &im = impacts.Host[i];
            if (lands_here(im)) ++CallInjectPtr[im.Object + 1];
            if (lives_here(im)) ++CallExcitePtr[im.Object + 1];
        }
        for (uint32_t o = 0; o < n_obj; ++o) CallInjectPtr[o + 1] += CallInjectPtr[o], CallExcitePtr[o + 1] += CallExcitePtr[o];
        CallInjectFrame.resize(CallInjectPtr[n_obj]), CallInjectDelta.resize(CallInjectPtr[n_obj]);
        excite.resize(CallExcitePtr[n_obj]);
        inject_at.assign(CallInjectPtr.begin(), CallInjectPtr.end() - 1), excite_at.assign(CallExcitePtr.begin(), CallExcitePtr.end() - 1);
        for (const uint32_t i : candidates) {
            const auto &im = impacts.Host[i];
            if (lands_here(im)) {
                const uint32_t at = inject_at[im.Object]++;
                CallInjectFrame[at] = im.Start + im.RenderLen, CallInjectDelta[at] = im.DeltaOff;
            }
            if (lives_here(im)) excite[excite_at[im.Object]++] = {im.Start, im.End};
        }
        CallExciteBegin.clear(), CallExciteEnd.clear();
        merged_ptr.assign(n_obj + 1, 0);
        for (uint32_t o = 0; o < n_obj; ++o) {
            // Increments sorted by the frame they land on (pulse lengths differ, so start order is not landing order).
            const uint32_t i0 = CallInjectPtr[o], i1 = CallInjectPtr[o + 1];
            bool sorted = true;
            for (uint32_t i = i0 + 1; i < i1 && sorted; ++i) sorted = CallInjectFrame[i - 1] <= CallInjectFrame[i];
            if (!sorted) {
                std::vector<std::pair<uint32_t, uint32_t>> tmp(i1 - i0);
                for (uint32_t i = i0; i < i1; ++i) tmp[i - i0] = {CallInjectFrame[i], CallInjectDelta[i]};
                std::stable_sort(tmp.begin(), tmp.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
                for (uint32_t i = i0; i < i1; ++i) CallInjectFrame[i] = tmp[i - i0].first, CallInjectDelta[i] = tmp[i - i0].second;
            }
            // Merged intervals during which the object holds a live impact.
            const auto e0 = excite.begin() + CallExcitePtr[o], e1 = excite.begin() + CallExcitePtr[o + 1];
            if (!std::is_sorted(e0, e1)) std::sort(e0, e1);
            const size_t first = CallExciteBegin.size();
            for (auto it = e0; it != e1; ++it) {
                if (CallExciteBegin.size() > first && it->first <= CallExciteEnd.back()) CallExciteEnd.back() = std::max(CallExciteEnd.back(), it->second);
                else CallExciteBegin.push_back(it->first), CallExciteEnd.push_back(it->second);
            }
            merged_ptr[o + 1] = uint32_t(CallExciteBegin.size());
        }
        CallExcitePtr = merged_ptr;
        const Lists lists{Plan.Stage(CallInjectPtr), Plan.Stage(CallInjectFrame), Plan.Stage(CallInjectDelta), Plan.Stage(CallExcitePtr), Plan.Stage(CallExciteBegin), Plan.Stage(CallExciteEnd)};
        Plan.Flush(stream);
        Stats.h2d_bytes += (CallInjectPtr.size() + CallExcitePtr.size() + 2 * CallInjectFrame.size() + 2 * CallExciteBegin.size()) * 4;
        return lists;
    };
    std::vector<uint32_t> candidates, carried; // impacts a stretch has to look at; those reaching past its end
    const auto lists_for = [&](uint32_t begin, uint32_t end, uint32_t first_new) {
        candidates = carried;
        for (uint32_t i = first_new; i < n; ++i) candidates.push_back(i);
        const Lists lists = build_lists(begin, end, candidates);
        carried.clear();
        for (const uint32_t i : candidates) {
            const auto &im = impacts.Host[i];
            if (std::max(im.Len ? im.Start + im.RenderLen : 0u, im.End) >= end) carried.push_back(i);
        }
        return lists;
    };
    const auto all_lists = [&](uint32_t begin, uint32_t end) { // after a fallback: every impact planned so far is looked at
        candidates.resize(n);
        std::iota(candidates.begin(), candidates.end(), 0u);
        return build_lists(begin, end, candidates);
    };
    const auto plan_for = [&](const Lists &lists, uint32_t begin, uint32_t wf, uint32_t blocks) {
        return RenderPlan{
            .SpanFrames = frames,
            .BlockFrames = block_frames,
            .FrameBegin = begin,
            .Frames = wf,
            .NSegments = 1,
            .SegmentFrames = blocks * block_frames,
            .ObjInjectPtr = lists.InjectPtr,
            .InjectFrame = lists.InjectFrame,
            .InjectDelta = lists.InjectDelta,
            .ObjExcitePtr = lists.ExcitePtr,
            .ExciteBegin = lists.ExciteBegin,
            .ExciteEnd = lists.ExciteEnd,
            .DeltaRe = DDeltaRe.Ptr,
            .DeltaIm = DDeltaIm.Ptr,
            .Partial = nullptr,
            .SegStateRe = nullptr,
            .SegStateIm = nullptr,
            .Speculation = DSpeculation.Ptr,
            .OnlyIf = 0u,
            .Debug = std::getenv("ME_RESONATOR_DEBUG") ? 1u : 0u,
            .WalkStates = nullptr,
            .WalkScales = nullptr,
            .WalkBlocksPerTile = TensorBlocksPerTile,
        };
    };

    Lists span_lists{};
    if (!tensor_span) {
        // Sample loop: every impact of the span is planned at once; the per-object lists are built while the pulse kernels run.
        plan_pulses_before(frames);
        span_lists = all_lists(0, frames);
        Stats.host_plan_ms += since(host_begin);
    }

    const uint32_t rows = ResonatorRows(NChunks);
    Stats.time_segments = 1;
    if (rows == 0) {
        plan_pulses_before(frames);
        join_pulses();
        ME_CUDA(cudaMemsetAsync(out_dev, 0, size_t(frames) * sizeof(float), stream));
    } else {
        // Launch windows bound the partial-row buffer (or the state stages); they start on block boundaries.
        const uint64_t blocks_in_budget = std::max<uint64_t>(1, PartialBudgetBytes / (uint64_t(rows) * sizeof(float)) / block_frames);
        uint32_t window = uint32_t(std::min<uint64_t>(blocks_in_budget * block_frames, frames));
        if (tensor_span) {
            const uint64_t tile_bytes = uint64_t(groups) * TmStateTileFloats(TensorBlocksPerTile) * sizeof(float);
            // Asked once: the query takes a driver lock that monitoring tools (nvidia-smi) contend for.
            if (StateBudgetBytes == 0) {
                size_t free_bytes = 0, total_bytes = 0;
                ME_CUDA(cudaMemGetInfo(&free_bytes, &total_bytes));
                StateBudgetBytes = std::min<uint64_t>(TensorStateBudgetBytes, free_bytes / 2);
            }
            const uint64_t budget = std::max<uint64_t>(StateBudgetBytes, DWalkStates.Capacity * sizeof(float));
            const uint64_t tiles_in_budget = std::max<uint64_t>(1, budget / tile_bytes);
            window = uint32_t(std::min<uint64_t>(tiles_in_budget * TensorTileFrames, (uint64_t(frames) + TensorTileFrames - 1) / TensorTileFrames * TensorTileFrames));
            if (PowersDirty || PowersVersion != TuningVersion) {
                DPowers.Reserve(size_t(groups) * kTmStagesPerGroup * TmPowerStageFloats());
                LaunchPowerTableKernel(view, DPowers.Ptr, stream, Counter);
                PowersDirty = false;
                PowersVersion = TuningVersion;
            }
        }
        // Stages reduced in one CTA of the tensor-core kernel (one partial mix row each): whole groups for large banks,
        // fractions of a group for small ones, aiming at six waves or more (a CTA runs for tens of microseconds, so a
        // short grid leaves a long tail).
        uint32_t stages_per_row = kTmStagesPerGroup;
        if (tensor_span) {
            const uint64_t tiles_in_window = (std::min(window, frames) + TensorTileFrames - 1) / TensorTileFrames;
            const uint64_t total_stages = uint64_t(groups) * kTmStagesPerGroup;
            // coarser rows (fewer partial rows to mix) only while twelve waves remain; finer ones until there are six
            while (stages_per_row < 16 * kTmStagesPerGroup && total_stages % (stages_per_row * 2) == 0 && total_stages / (stages_per_row * 2) * tiles_in_window >= 12 * 148) stages_per_row *= 2;
            while (stages_per_row > 64 && total_stages / stages_per_row * tiles_in_window < 6 * 148) stages_per_row /= 2;
        }
        const uint32_t mix_rows = groups ? uint32_t(uint64_t(groups) * kTmStagesPerGroup / stages_per_row) : 0;
        const uint32_t ctas = NChunks / kBlockThreads;
        bool lists_cover_span = !tensor_span; // span_lists hold every impact (sample loop, or after a fallback)
        for (uint32_t begin = 0; begin < frames; begin += window) {
            const uint32_t wf = std::min(window, frames - begin);
            const uint32_t blocks = (wf + block_frames - 1) / block_frames;
            if (!lists_cover_span && !(tensor_span && !SpeculationFailed)) {
                // A tensor span that fell back to the sample loop: the remaining windows take lists over everything.
                host_begin = Clock::now();
                plan_pulses_before(frames);
                span_lists = all_lists(begin, frames);
                lists_cover_span = true;
                Stats.host_plan_ms += since(host_begin);
            }
            RenderPlan plan = plan_for(span_lists, begin, wf, blocks);
            const auto seed_segments = [&](RenderPlan &p, uint32_t segments) {
                if (segments <= 1) return;
                const size_t seg_floats = size_t(segments - 1) * NChunks * kLanes;
                DSegRe.Reserve(seg_floats), DSegIm.Reserve(seg_floats);
                p.SegStateRe = DSegRe.Ptr, p.SegStateIm = DSegIm.Ptr;
                LaunchSegmentScan(view, p, DSegRe.Ptr, DSegIm.Ptr, stream, Counter);
            };
            // OR of the window's speculation words (one per sub-window).
            const auto speculation_failed = [&](uint32_t words) -> uint32_t {
                uint32_t failed[64] = {};
                ME_CUDA(cudaMemcpyAsync(failed, DSpeculation.Ptr, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
                ME_CUDA(cudaStreamSynchronize(stream));
                uint32_t any = 0;
                for (uint32_t i = 0; i < words; ++i) any |= failed[i];
                if (any && std::getenv("ME_RESONATOR_DEBUG")) fprintf(stderr, "[me] speculation failed: code %u (window %u+%u, %u segments)\n", any, begin, wf, plan.NSegments);
                return any;
            };
            // A culling decision inside the window (or an increment off the time-block grid): the window is rendered
            // again sequentially in time by the sample loop (still on the GPU).
            const auto render_sequentially = [&] {
                SpeculationFailed = true;
                ++Stats.scan_fallbacks;
                DPartial.Reserve(size_t(rows) * wf);
                plan.Partial = DPartial.Ptr;
                plan.NSegments = 1;
                plan.SegmentFrames = blocks * block_frames;
                plan.SegStateRe = plan.SegStateIm = nullptr;
                plan.Speculation = DSpeculation.Ptr;
                LaunchResonatorKernel(view, plan, Steps, stream, Counter);
            };
            cudaEvent_t k0 = NextEvent(), k1 = NextEvent();
            EventKind.push_back(0);
            ME_CUDA(cudaEventRecord(k0, stream));
            bool mixed = false; // the window's MixKernel has been launched (tensor-core form launches it ahead of its flag check)
            bool tensor_window = tensor_span && !SpeculationFailed;
            uint32_t segments = 1;
            if (tensor_window) {
                // The walk kernel steps every chunk through the window one 256-frame time block at a time (culling applied exactly as
                // it goes) and writes the block-start states; the tcgen05 kernel turns them into per-group-set mixes.
                const uint32_t tiles = (wf + TensorTileFrames - 1) / TensorTileFrames;
                DWalkStates.Reserve(size_t(tiles) * groups * TmStateTileFloats(TensorBlocksPerTile));
                DWalkScales.Reserve(size_t(tiles) * groups * TmScaleTileFloats(TensorBlocksPerTile));
                DGroupMix.Reserve(size_t(mix_rows) * wf);
                // Sub-windows of the pipeline, in tiles: the first ones are short (1, 2, .. tiles) so that the device has work after a
                // fraction of the host's planning; the rest take SubWindowTiles each.
                std::vector<uint32_t> sub_begin{0};
                if (piped && !small_bank) {
                    for (uint32_t at = 0, size = 1; at < tiles; size = std::min(size + 1, SubWindowTiles)) {
                        at = std::min(tiles, at + size);
                        if (at < tiles) sub_begin.push_back(at * TensorTileFrames);
                    }
                } else if (piped && small_banks_piped && tiles > 4) {
                    // A small bank is launch-bound (every sub-window costs a seeded walk, its scan and its conditional repeat): two
                    // sub-windows, so that all but a fraction of the host's planning runs beside the device.
                    sub_begin.push_back(2 * TensorTileFrames);
                }
                const uint32_t subs = uint32_t(sub_begin.size());
                sub_begin.push_back(wf);
                if (subs > 64) Fail(ME_BAD_ARG, "launch window of %u sub-windows (at most 64): raise ME_WALK_SUBWINDOW_TILES", subs);
                ME_CUDA(cudaMemsetAsync(DSpeculation.Ptr, 0, 64 * sizeof(uint32_t), stream));
                const int side_at_begin = Side;
                if (subs > 1) {
                    // The window's start state, should the whole window have to be rendered again by the sample loop.
                    const size_t modes = size_t(NChunks) * kLanes;
                    DSnapRe.Reserve(modes), DSnapIm.Reserve(modes), DSnapLive.Reserve(NChunks), DSnapRinging.Reserve(std::max(n_obj, 1u));
                    ME_CUDA(cudaMemcpyAsync(DSnapRe.Ptr, DStateRe[Side].Ptr, modes * sizeof(float), cudaMemcpyDeviceToDevice, stream));
                    ME_CUDA(cudaMemcpyAsync(DSnapIm.Ptr, DStateIm[Side].Ptr, modes * sizeof(float), cudaMemcpyDeviceToDevice, stream));
                    ME_CUDA(cudaMemcpyAsync(DSnapLive.Ptr, DChunkLive[Side].Ptr, NChunks, cudaMemcpyDeviceToDevice, stream));
                    ME_CUDA(cudaMemcpyAsync(DSnapRinging.Ptr, DObjRinging[Side].Ptr, n_obj, cudaMemcpyDeviceToDevice, stream));
                }
                bool seeded = false;
                uint32_t walked_segments = 1; // most segments any sub-window's walk was split into
                for (uint32_t sub = 0; sub < subs; ++sub) {
                    const uint32_t sb = begin + sub_begin[sub], sf = sub_begin[sub + 1] - sub_begin[sub];
                    const uint32_t sub_blocks = (sf + block_frames - 1) / block_frames;
                    host_begin = Clock::now();
                    const uint32_t first_new = n;
                    // The very first batch is cut short - the strikes of the span's first block alone (an offline timeline starts
                    // with one per voice) - so that the device has pulse kernels to run while the rest of the first sub-window is
                    // still being admitted and planned: of the ~0.24 ms the host used to spend ahead of the first launch, half.
                    // (Same-box A/B over 10 steps: 6.76-6.98 ms against 6.97-7.02.)
                    if (sub == 0 && piped && !small_bank && sf > block_frames) plan_pulses_before(sb + block_frames);
                    plan_pulses_before(sb + sf);
                    const Lists lists = lists_for(sb, sb + sf, first_new);
                    Stats.host_plan_ms += since(host_begin);
                    join_pulses();
                    RenderPlan walk = plan_for(lists, sb, sf, sub_blocks);
                    walk.Speculation = DSpeculation.Ptr + sub;
                    walk.WalkStates = DWalkStates.Ptr + size_t((sb - begin) / TensorTileFrames) * groups * TmStateTileFloats(TensorBlocksPerTile);
                    walk.WalkScales = DWalkScales.Ptr + size_t((sb - begin) / TensorTileFrames) * groups * TmScaleTileFloats(TensorBlocksPerTile);
                    // A bank of few chunk groups cannot fill the SMs with one CTA per group: its walk is split into seeded
                    // segments along time (the scan of the sample loop). A culling decision inside the window invalidates
                    // the seeds; the walk is then repeated sequentially (it is cheap next to the mix).
                    segments = RequestedSegments ? RequestedSegments : (SeededWalkFailed || groups >= 148 ? 1u : std::min<uint32_t>(16, 296 / groups)); // two walk CTAs fit an SM: one wave
                    segments = std::max(1u, std::min(segments, sub_blocks));
                    const uint32_t seg_blocks = (sub_blocks + segments - 1) / segments;
                    segments = (sub_blocks + seg_blocks - 1) / seg_blocks;
                    walk.NSegments = segments;
                    walk.SegmentFrames = seg_blocks * block_frames;
                    seed_segments(walk, segments);
                    Timed(1, stream, [&] { LaunchStateWalkKernel(view, walk, stream, Counter); });
                    if (segments > 1) {
                        // The sequential repeat, launched right behind the seeded walk and run only if that one raised a flag
                        // (RenderPlan::OnlyIf): the device decides, the host does not wait to find out.
                        seeded = true;
                        RenderPlan repeat = walk;
                        repeat.NSegments = 1;
                        repeat.SegmentFrames = sub_blocks * block_frames;
                        repeat.SegStateRe = repeat.SegStateIm = nullptr;
                        repeat.OnlyIf = 7u;
                        Timed(1, stream, [&] { LaunchStateWalkKernel(view, repeat, stream, Counter); });
                    }
                    walked_segments = std::max(walked_segments, segments);
                    if (sub + 1 < subs) {
                        Side ^= 1;
                        view = View();
                    }
                }
                plan.Speculation = DSpeculation.Ptr;
                segments = walked_segments;
                Timed(2, stream, [&] {
                    LaunchTensorMixKernel({.Groups = groups, .StagesPerRow = stages_per_row, .Tiles = tiles, .BlocksPerTile = TensorBlocksPerTile, .Frames = wf, .Powers = DPowers.Ptr, .States = DWalkStates.Ptr, .Scales = DWalkScales.Ptr, .Partial = DGroupMix.Ptr}, stream);
                });
                ++Counter.Launches;
                // The fixed-order mix goes out before the flags are read: one host round trip per window, behind its last launch.
                ME_CUDA(cudaEventRecord(k1, stream));
                LaunchMixKernel(DGroupMix.Ptr, mix_rows, plan, current_pulses(), out_dev + begin, stream, Counter);
                mixed = true;
                const uint32_t flags = speculation_failed(subs);
                if (seeded && (flags & 7u)) {
                    SeededWalkFailed = true;
                    ++Stats.scan_fallbacks;
                    segments = 1;
                }
                // Only code 8 (an increment off the time-block grid) can invalidate a sequential walk.
                if (flags & 8u) {
                    tensor_window = false;
                    mixed = false;
                    if (subs > 1) {
                        const size_t modes = size_t(NChunks) * kLanes;
                        Side = side_at_begin;
                        ME_CUDA(cudaMemcpyAsync(DStateRe[Side].Ptr, DSnapRe.Ptr, modes * sizeof(float), cudaMemcpyDeviceToDevice, stream));
                        ME_CUDA(cudaMemcpyAsync(DStateIm[Side].Ptr, DSnapIm.Ptr, modes * sizeof(float), cudaMemcpyDeviceToDevice, stream));
                        ME_CUDA(cudaMemcpyAsync(DChunkLive[Side].Ptr, DSnapLive.Ptr, NChunks, cudaMemcpyDeviceToDevice, stream));
                        ME_CUDA(cudaMemcpyAsync(DObjRinging[Side].Ptr, DSnapRinging.Ptr, n_obj, cudaMemcpyDeviceToDevice, stream));
                        view = View();
                    }
                    const Lists lists = all_lists(begin, begin + wf);
                    plan = plan_for(lists, begin, wf, blocks);
                    render_sequentially();
                } else {
                    ++Stats.tensor_windows;
                }
            } else {
                join_pulses();
                // Segments of the block-parallel scan along time: enough (chunk-CTA x segment) units to even out the
                // 148 SMs, each a whole number of blocks.
                segments = RequestedSegments;
                if (segments == 0) segments = SpeculationFailed ? 1 : std::min<uint32_t>(64, (16 * 296 + ctas - 1) / ctas);
                segments = std::max(1u, std::min(segments, blocks));
                const uint32_t seg_blocks = (blocks + segments - 1) / segments;
                segments = (blocks + seg_blocks - 1) / seg_blocks;
                DPartial.Reserve(size_t(rows) * wf);
                plan.Partial = DPartial.Ptr;
                plan.NSegments = segments;
                plan.SegmentFrames = seg_blocks * block_frames;
                if (segments > 1) {
                    ME_CUDA(cudaMemsetAsync(DSpeculation.Ptr, 0, 64 * sizeof(uint32_t), stream));
                    seed_segments(plan, segments);
                    LaunchResonatorKernel(view, plan, Steps, stream, Counter);
                    if (speculation_failed(1)) {
                        render_sequentially();
                        segments = 1;
                    }
                } else {
                    LaunchResonatorKernel(view, plan, Steps, stream, Counter);
                }
            }
            if (!mixed) ME_CUDA(cudaEventRecord(k1, stream));
            Stats.time_segments = std::max(Stats.time_segments, segments);
            Stats.partial_rows = tensor_window ? mix_rows : rows;
            if (!mixed) {
                if (tensor_window) LaunchMixKernel(DGroupMix.Ptr, mix_rows, plan, current_pulses(), out_dev + begin, stream, Counter);
                else LaunchMixKernel(DPartial.Ptr, rows, plan, current_pulses(), out_dev + begin, stream, Counter);
            }
            Side ^= 1;
            view = View();
        }
    }
    // Whatever the windows did not reach (nothing, unless the span had no chunks to render).
    plan_pulses_before(frames);
    join_pulses();
    if (any_click) LaunchClickKernel(impacts.Dev, tails.Dev, n, out_dev, frames, stream, Counter);
    Plan.Seal(stream);
    for (uint32_t o = 0; o < n_obj; ++o) Stats.mode_samples += uint64_t(TunedModeCount[o]) * frames;
}

cudaEvent_t Bank::NextEvent() {
    if (EventsUsed == EventPool.size()) {
        cudaEvent_t e;
        ME_CUDA(cudaEventCreate(&e));
        EventPool.push_back(e);
    }
    return EventPool[EventsUsed++];
}

cudaEvent_t Bank::NextJoin() {
    if (JoinsUsed == JoinPool.size()) {
        cudaEvent_t e;
        ME_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        JoinPool.push_back(e);
    }
    return JoinPool[JoinsUsed++];
}

void Bank::RenderTimeline(const MeModalEvent *events, const uint64_t *event_frames, uint32_t n_events, uint64_t total_frames, uint32_t block_frames, float *out, bool out_is_device, cudaStream_t stream, bool use_own_stream) {
    RequireInstalled();
    if (total_frames == 0) return;
    if (!out) Fail(ME_BAD_ARG, "null output buffer");
    if (block_frames == 0) Fail(ME_BAD_ARG, "block_frames must be positive");
    if (total_frames >= (uint64_t(1) << 31)) Fail(ME_BAD_ARG, "total_frames must be below 2^31 per call");
    if (n_events && (!events || !event_frames)) Fail(ME_BAD_ARG, "null event arrays");
    ME_CUDA(cudaSetDevice(Device));
    const cudaStream_t caller = use_own_stream ? nullptr : stream;
    JoinsUsed = 0;
    if (!use_own_stream) {
        const cudaEvent_t fork = NextJoin();
        ME_CUDA(cudaEventRecord(fork, caller));
        ME_CUDA(cudaStreamWaitEvent(OwnStream, fork, 0));
    }
    stream = OwnStream;
    if (TuningDirty) {
        UploadTuning(stream);
        for (const auto o : RetunedObjects) ResetObjectOnDevice(o, false, stream);
        RetunedObjects.clear();
    }
    SpeculationFailed = false;
    SeededWalkFailed = false;

    Stats = {};
    Counter = {};
    EventsUsed = 0;
    EventKind.clear();
    StatsResolved = false;
    float *out_dev = out;
    if (!out_is_device) {
        DOut.Reserve(total_frames);
        out_dev = DOut.Ptr;
    }
    ME_CUDA(cudaEventRecord(EvBegin, stream));

    // The callback's own prologue (RenderModal :496-499): adopt the flush flag, then drain the ring.
    std::vector<MeModalEvent> ring;
    if (FlushEvents) {
        FlushEvents = false;
        EventRead.store(EventWrite.load(std::memory_order_relaxed), std::memory_order_relaxed);
    }
    {
        auto read = EventRead.load(std::memory_order_relaxed);
        const auto write = EventWrite.load(std::memory_order_acquire);
        for (; read != write; ++read) ring.push_back(Events[read % EventCapacity]);
        EventRead.store(read, std::memory_order_release);
    }

    // What the impacts of any one span can need of the plan and the pulse buffers, from the events alone (a span's impacts are
    // events of the timeline or remainders of earlier ones): the buffers are sized before the first impact is admitted, because
    // the kernels of early impacts are already running while later ones are still being planned.
    const uint32_t total = uint32_t(total_frames);
    const uint32_t n_obj = ObjectCount();
    SpanBounds bounds;
    const auto bound = [&](uint32_t object, uint64_t samples_left) {
        if (object >= n_obj) return;
        const uint64_t len = std::min<uint64_t>(samples_left, total);
        const uint32_t stride = ObjStride[object], warps = (stride / kLanes + 31) / 32;
        ++bounds.Impacts, bounds.Force += len, bounds.Delta += stride, bounds.Warps += warps, bounds.Rows += warps * std::min<uint64_t>(len + kTmBlock, total);
    };
    for (const auto &im : Impacts) bound(im.Object, im.SamplesLeft);
    for (const auto &e : ring)
        if (e.kind == 0 && e.pulse_step > 0) bound(e.object, PulseSamples(e.pulse_step));
    for (uint32_t i = 0; i < n_events; ++i) { // (one pass: the events' validity and their bounds)
        if (event_frames[i] % block_frames != 0 || event_frames[i] >= total_frames) Fail(ME_BAD_ARG, "event %u: frame %llu is not a block boundary inside the timeline", i, (unsigned long long)event_frames[i]);
        if (i && event_frames[i] < event_frames[i - 1]) Fail(ME_BAD_ARG, "event frames must be ascending");
        if (events[i].kind == 0 && events[i].pulse_step > 0) bound(events[i].object, PulseSamples(events[i].pulse_step));
    }

    if (std::max({bounds.Force, bounds.Delta, bounds.Rows, bounds.Warps}) >= (uint64_t(1) << 32)) Fail(ME_BAD_ARG, "impact buffers exceed 2^32 entries in one call");

    // A Silence event rewrites resonator state, so the timeline is rendered in spans cut at silences.
    uint32_t next_event = 0;
    uint32_t span_begin = 0;
    std::vector<ScheduledImpact> scheduled;
    while (span_begin < total) {
        uint32_t span_end = total;
        for (uint32_t i = next_event; i < n_events; ++i) {
            if (events[i].kind == 1 && event_frames[i] > span_begin) {
                span_end = uint32_t(event_frames[i]);
                break;
            }
        }
        const uint32_t span_frames = span_end - span_begin;
        scheduled.clear();
        // In-flight impacts for the MaxImpacts cap (ActivateImpact :29). Impacts retire on block boundaries and events arrive
        // in frame order, so a count per retirement block and a cursor replace a heap of retirement frames.
        std::vector<uint32_t> retire_in_block(size_t(span_frames) / block_frames + 2, 0);
        uint32_t retire_cursor = 0, in_flight = 0;
        scheduled.reserve(bounds.Impacts);
        const auto note_retirement = [&](const ScheduledImpact &s) {
            ++in_flight;
            if (!s.Survives) ++retire_in_block[(uint64_t(s.End) + block_frames - 1) / block_frames];
        };
        const auto admit = [&](const HostImpact &im, uint32_t start) {
            ScheduledImpact s{.AtStart = im, .AtEnd = im, .Start = start, .End = 0, .Survives = false};
            Schedule(s, block_frames, span_frames);
            scheduled.push_back(s);
            note_retirement(s);
        };
        const auto apply = [&](const MeModalEvent &e, uint32_t start) {
            if (e.object >= n_obj) return; // DrainEvents :71
            if (e.kind == 0) {
                if (!(e.pulse_step > 0)) return;
                // TriggerModalStrike refuses a strike outside the object's excitable points before it becomes an event
                // (AudioSystem.cpp: `excitable_index >= min(Vertices, Positions)` returns early); the C ABI takes raw events,
                // so the same guard sits here: such an event is dropped like one aimed at a missing slot (DrainEvents :71), never turned
                // into a shape read out of bounds.
                if (e.ex_pos >= ShapePoints[e.object]) return;
                for (; retire_cursor <= start / block_frames; ++retire_cursor) in_flight -= retire_in_block[retire_cursor]; // retired at or before `start`
                if (in_flight >= MaxImpacts) return;
                admit(MakeImpact(e), start);
            } else if (e.kind == 1) {
                // Only ever at the first frame of a span: nothing of this span has been rendered (or planned) yet.
                const size_t before = scheduled.size();
                std::erase_if(scheduled, [&](const ScheduledImpact &s) { return s.AtStart.Object == e.object; });
                if (scheduled.size() != before) {
                    std::fill(retire_in_block.begin(), retire_in_block.end(), 0u);
                    retire_cursor = 0, in_flight = 0;
                    for (const auto &s : scheduled) note_retirement(s);
                }
                ResetObjectOnDevice(e.object, true, stream);
            }
        };
        // The span's impacts are admitted on demand, in start order: RenderSpan asks for those starting before a frame when it is
        // about to plan them, so the kernels of the first stretch run while the host is still admitting the next.
        bool carried_in = false;
        const auto admit_before = [&](uint32_t limit) {
            if (!carried_in) {
                carried_in = true;
                for (const auto &im : Impacts) admit(im, 0);
                Impacts.clear();
                if (span_begin == 0)
                    for (const auto &e : ring) apply(e, 0);
            }
            const uint64_t until = uint64_t(span_begin) + std::min(limit, span_frames);
            for (; next_event < n_events && event_frames[next_event] < until; ++next_event) apply(events[next_event], uint32_t(event_frames[next_event]) - span_begin);
        };

        RenderSpan(span_frames, block_frames, scheduled, admit_before, bounds, out_dev + span_begin, stream);
        for (const auto &s : scheduled)
            if (s.Survives) Impacts.push_back(s.AtEnd);
        span_begin = span_end;
    }
    ME_CUDA(cudaEventRecord(EvEnd, stream));
    Stats.kernel_launches = Counter.Launches;
    if (!use_own_stream) ME_CUDA(cudaStreamWaitEvent(caller, EvEnd, 0)); // the caller's stream continues behind the render

    if (!out_is_device) {
        POut.Reserve(total_frames);
        ME_CUDA(cudaMemcpyAsync(POut.Ptr, out_dev, total_frames * sizeof(float), cudaMemcpyDeviceToHost, stream));
        ME_CUDA(cudaStreamSynchronize(stream));
        for (uint64_t i = 0; i < total_frames; ++i) out[i] += POut.Ptr[i];
        Stats.d2h_bytes += total_frames * sizeof(float);
    }
}

const MeRenderStats &Bank::LastStats() {
    if (!StatsResolved && EvEnd) {
        cudaSetDevice(Device);
        if (cudaEventSynchronize(EvEnd) == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, EvBegin, EvEnd) == cudaSuccess) Stats.total_device_ms = ms;
            float kernel[4] = {0.f, 0.f, 0.f, 0.f};
            const bool trace = std::getenv("ME_RENDER_TRACE") != nullptr; // the call's timeline: every bracketed launch against the call's start
            for (uint32_t i = 0; i + 1 < EventsUsed; i += 2) {
                if (cudaEventElapsedTime(&ms, EventPool[i], EventPool[i + 1]) == cudaSuccess) kernel[EventKind[i / 2]] += ms;
                if (trace) {
                    float t0 = 0.f, t1 = 0.f;
                    cudaEventElapsedTime(&t0, EvBegin, EventPool[i]), cudaEventElapsedTime(&t1, EvBegin, EventPool[i + 1]);
                    static const char *names[4] = {"window", "walk", "tensor mix", "force + pulses"};
                    fprintf(stderr, "[me trace] %-14s %8.3f .. %8.3f ms\n", names[EventKind[i / 2]], t0, t1);
                }
            }
            if (trace) fprintf(stderr, "[me trace] call            0.000 .. %8.3f ms (host planning %.3f ms)\n", Stats.total_device_ms, Stats.host_plan_ms);
            Stats.resonator_kernel_ms = kernel[0];
            Stats.walk_kernel_ms = kernel[1];
            Stats.tensor_mix_kernel_ms = kernel[2];
            Stats.pulse_kernels_ms = kernel[3];
        }
        StatsResolved = true;
    }
    return Stats;
}

void Bank::GetModeColumn(MeModeColumn which, float *out) {
    if (!out) Fail(ME_BAD_ARG, "null output");
    const std::vector<float> *host = nullptr;
    switch (which) {
        case ME_COL_COEFF_RE: host = &CoeffRe; break;
        case ME_COL_COEFF_IM: host = &CoeffIm; break;
        case ME_COL_RADIATION_GAIN: host = &RadiationGain; break;
        case ME_COL_RADIATION_AREA: host = &RadiationArea; break;
        case ME_COL_OUT_PHASE_IM: host = &OutPhaseIm; break;
        case ME_COL_OUT_PHASE_RE: host = &OutPhaseRe; break;
        case ME_COL_DEFLECTION_GAIN: host = &DeflectionGain; break;
        case ME_COL_QUAD_COMPLIANCE: host = &QuadCompliance; break;
        case ME_COL_QUAD_DRIVE_SCALE: host = &QuadDriveScale; break;
        case ME_COL_STATE_RE:
        case ME_COL_STATE_IM: break;
        default: Fail(ME_BAD_ARG, "unknown mode column %d", int(which));
    }
    if (host) {
        std::copy(host->begin(), host->end(), out);
        return;
    }
    if (!Installed) {
        std::fill_n(out, ModeTotal(), 0.f);
        return;
    }
    ME_CUDA(cudaSetDevice(Device));
    ME_CUDA(cudaStreamSynchronize(OwnStream));
    std::vector<float> padded(size_t(NChunks) * kLanes);
    const float *src = which == ME_COL_STATE_RE ? DStateRe[Side].Ptr : DStateIm[Side].Ptr;
    ME_CUDA(cudaMemcpy(padded.data(), src, padded.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (uint32_t o = 0; o < ObjectCount(); ++o) std::copy_n(padded.begin() + size_t(ObjFirstChunk[o]) * kLanes, ModeCount[o], out + ModeOffset[o]);
}

void Bank::GetObjectStatus(uint32_t slot, uint32_t *live_mode_count, uint32_t *ringing) {
    CheckSlot(slot);
    RequireInstalled();
    ME_CUDA(cudaSetDevice(Device));
    ME_CUDA(cudaStreamSynchronize(OwnStream));
    const uint32_t chunks = ObjStride[slot] / kLanes;
    std::vector<uint8_t> live(chunks);
    uint8_t ring = 0;
    if (chunks) ME_CUDA(cudaMemcpy(live.data(), DChunkLive[Side].Ptr + ObjFirstChunk[slot], chunks, cudaMemcpyDeviceToHost));
    ME_CUDA(cudaMemcpy(&ring, DObjRinging[Side].Ptr + slot, 1, cudaMemcpyDeviceToHost));
    // The live set is a prefix of the tuned set, a whole number of chunks or the tuned count itself (:139,:146).
    uint32_t last = 0;
    for (uint32_t c = 0; c < chunks; ++c)
        if (live[c]) last = c + 1;
    if (live_mode_count) *live_mode_count = std::min(last * kLanes, TunedModeCount[slot]);
    if (ringing) *ringing = ring;
}

// DealObjects, ModalAudio.cpp:430-461.
void DealObjects(const uint64_t *costs, uint32_t n, uint32_t renderers, uint32_t *owner, uint32_t *local_slot) {
    if (renderers == 0) Fail(ME_BAD_ARG, "no renderer to deal objects to");
    if (n && (!costs || !owner)) Fail(ME_BAD_ARG, "null cost / owner array");
    if (renderers == 1) {
        std::fill_n(owner, n, 0u);
    } else {
        std::vector<uint32_t> order(n);
        std::iota(order.begin(), order.end(), 0u);
        // Heaviest first, and by object for equal weights, so the deal never depends on the sort's own tie-breaking (:452).
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return costs[a] != costs[b] ? costs[a] > costs[b] : a < b; });
        std::vector<uint64_t> load(renderers, 0);
        for (const uint32_t o : order) {
            const auto least = uint32_t(std::min_element(load.begin(), load.end()) - load.begin());
            load[least] += costs[o];
            owner[o] = least;
        }
    }
    if (local_slot) {
        std::vector<uint32_t> next(renderers, 0);
        for (uint32_t o = 0; o < n; ++o) local_slot[o] = next[owner[o]]++;
    }
}

void Bank::GetObjectLayout(uint32_t slot, uint32_t *mode_offset, uint32_t *mode_count, uint32_t *tuned, float *radius) const {
    CheckSlot(slot);
    if (mode_offset) *mode_offset = ModeOffset[slot];
    if (mode_count) *mode_count = ModeCount[slot];
    if (tuned) *tuned = TunedModeCount[slot];
    if (radius) *radius = RadiantRadius[slot];
}

} // namespace me
