// Drop-in replacement translation unit for src/audio/ModalAudio.cpp of khiner/MeshEditor: the SAME entry points with the SAME
// signatures (src/audio/ModalAudio.h:297-315 — AddModalObject, TuneModalObject, SetModalObjectShapes, FindModalObject,
// InstallModalBank, EnqueueModalEvent, RenderModal — plus ModalAudio's constructor and ModalRenderPool, which that file also
// defines), implemented on top of libme_modal.so's C ABI (include/me_modal.h). Build the reference with this file INSTEAD of
// ModalAudio.cpp and link libme_modal.so: the reference's own tests/ModalRenderTest.cpp and tests/ModalBench.h then compile
// unchanged and render on the B200 (oracle/Makefile target `shim_tests`, run by tests/test_reference_shim_gpu.py).
//
// The reference's ModalBank stays what callers read and write in place (OutGain, ListenerGain, RigidInvMass, Entities, the
// UI's columns). Every bank a caller builds is mirrored slot for slot by a device bank; the per-mode columns TuneModalObject
// computes are read back from the library, which evaluates them with the reference's own libm calls (bit-identical).
// It is compiled inside the reference tree (it includes the reference's headers), so it is not part of libme_modal.so.
#include "audio/ModalAudio.h"
#include "audio/ModalModes.h"
#include "audio/SurfaceContact.h"

#include "me_modal.h"

#include <algorithm>
#include <cstddef>
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

static_assert(sizeof(ModalEvent) == sizeof(MeModalEvent), "MeModalEvent mirrors ModalEvent field for field");

namespace {
struct Tables {
    std::mutex Mutex;
    std::unordered_map<const ModalBank *, MeBank *> Building; // banks under construction, keyed by the caller's `next`
    std::unordered_map<const ModalAudio *, MeBank *> Live;    // the published bank of each ModalAudio
};
Tables &T() {
    static Tables t;
    return t;
}
int Device() {
    static const int device = [] {
        const char *env = std::getenv("ME_DEVICE");
        return env ? std::atoi(env) : 0;
    }();
    return device;
}
void Check(MeStatus status) {
    if (status != ME_OK && status != ME_QUEUE_FULL) throw std::runtime_error(me_last_error());
}
MeBank *BuildingBank(const ModalBank &b) {
    auto &t = T();
    std::scoped_lock lock{t.Mutex};
    auto &slot = t.Building[&b];
    if (!slot) Check(me_bank_create(b.SampleRate, Device(), &slot));
    return slot;
}
// The device bank mirroring host bank `b`: a live one when some ModalAudio has published `b`, else the one being built.
MeBank *Mirror(const ModalBank &b) {
    {
        auto &t = T();
        std::scoped_lock lock{t.Mutex};
        for (const auto &[audio, bank] : t.Live)
            if (audio->Live.get() == &b) return bank;
    }
    return BuildingBank(b);
}
std::vector<float> Flatten(const std::vector<std::vector<vec3>> &shapes) { // [point][mode] vec3 -> [point][mode][3]
    std::vector<float> flat;
    for (const auto &point : shapes)
        for (const auto &s : point) flat.insert(flat.end(), {s.x, s.y, s.z});
    return flat;
}
// Host columns of one object <- what the library computed for it (TuneModalObject's results, ModalAudio.cpp:340-393).
void PullColumns(MeBank *bank, ModalBank &b, uint32_t object) {
    const uint32_t total = me_bank_mode_total(bank);
    std::vector<float> column(total);
    const uint32_t k0 = b.ModeOffset[object], count = b.ModeCount[object];
    const std::pair<MeModeColumn, std::vector<float> *> columns[] = {
        {ME_COL_COEFF_RE, &b.CoeffRe}, {ME_COL_COEFF_IM, &b.CoeffIm}, {ME_COL_RADIATION_GAIN, &b.RadiationGain}, {ME_COL_RADIATION_AREA, &b.RadiationArea},
        {ME_COL_OUT_PHASE_IM, &b.OutPhaseIm}, {ME_COL_OUT_PHASE_RE, &b.OutPhaseRe}, {ME_COL_DEFLECTION_GAIN, &b.DeflectionGain},
        {ME_COL_QUAD_COMPLIANCE, &b.QuadCompliance}, {ME_COL_QUAD_DRIVE_SCALE, &b.QuadDriveScale},
    };
    for (const auto &[which, host] : columns) {
        Check(me_bank_get_mode_column(bank, which, column.data()));
        std::copy_n(column.begin() + k0, count, host->begin() + k0);
    }
    uint32_t offset = 0, modes = 0, tuned = 0;
    float radius = 0.f;
    Check(me_bank_get_object_layout(bank, object, &offset, &modes, &tuned, &radius));
    b.TunedModeCount[object] = tuned;
    b.LiveModeCount[object] = tuned;
    b.RadiantRadius[object] = radius;
}
} // namespace

// ---- ModalRenderPool / ModalAudio: the parts of ModalAudio.cpp that are not the bank ------------------------------------------
// The device renders every object of a block in one launch, so the pool holds no threads: its width is kept only because
// callers set and read it (ModalBench.h:57, AudioSystem.cpp:1109).
ModalRenderPool::~ModalRenderPool() {
    // The pool is a member of ModalAudio and the only one with a user-declared destructor: the ModalAudio that owns it is going
    // away, and its device bank with it.
    const auto *owner = reinterpret_cast<const ModalAudio *>(reinterpret_cast<const char *>(this) - offsetof(ModalAudio, RenderPool));
    auto &t = T();
    std::scoped_lock lock{t.Mutex};
    if (const auto it = t.Live.find(owner); it != t.Live.end()) {
        me_bank_free(it->second);
        t.Live.erase(it);
    }
}
void ModalRenderPool::SetSize(uint32_t count) {
    const std::scoped_lock lock{ResizeMutex};
    Active = std::max(1u, count);
}
void ModalRenderPool::SetWorkgroup(void *workgroup) {
    const std::scoped_lock lock{ResizeMutex};
    Workgroup = workgroup;
}

ModalAudio::ModalAudio() : Live{std::make_unique<ModalBank>()}, Published{Live.get()}, Surface{MakeSurfaceAudioState()} {}

// ---- the bank -------------------------------------------------------------------------------------------------------------------
uint32_t AddModalObject(ModalBank &b, entt::entity e, const ModalModes &modes) {
    const auto count = uint32_t(modes.Freqs.size());
    const auto slot = uint32_t(b.Entities.size());
    const auto flat = Flatten(modes.Shapes);
    uint32_t device_slot = 0;
    MeBank *bank = BuildingBank(b);
    Check(me_bank_add_object(bank, count, uint32_t(modes.Shapes.size()), flat.data(), modes.Positions.empty() ? nullptr : &modes.Positions[0].x, modes.Indices.data(), uint32_t(modes.Indices.size()), &device_slot));
    if (device_slot != slot) throw std::logic_error("device bank out of step with the host bank");
    // The host bank's slot, with the sizes and defaults of ModalAudio.cpp:291-312 (what an untuned object stands at).
    b.Entities.push_back(e);
    b.ModeOffset.push_back(uint32_t(b.CoeffRe.size()));
    b.ModeCount.push_back(count);
    b.TunedModeCount.push_back(count);
    b.LiveModeCount.push_back(count);
    b.ShapeOffset.push_back(uint32_t(b.ShapeX.size()));
    b.Ringing.push_back(0);
    b.RigidVel.emplace_back(0.f);
    for (auto *col : {&b.OutGain, &b.RigidInvMass, &b.RadiatorB0, &b.AirB0, &b.AirB1, &b.AirB2, &b.RecoilA1, &b.RecoilA2, &b.RadiatorZ1, &b.RadiatorZ2, &b.AirZ1, &b.AirZ2}) col->push_back(0.f);
    for (auto *col : {&b.ListenerGain, &b.DeflectionScale}) col->push_back(1.f);
    for (auto *col : {&b.CoeffRe, &b.CoeffIm, &b.StateRe, &b.StateIm, &b.RadiationGain, &b.RadiationArea, &b.DeflectionGain, &b.QuadCompliance, &b.QuadDriveScale}) col->resize(col->size() + count, 0.f);
    b.OutPhaseIm.resize(b.OutPhaseIm.size() + count, 1.f);
    b.OutPhaseRe.resize(b.OutPhaseRe.size() + count, 0.f);
    for (size_t i = 0; i < flat.size(); i += 3) b.ShapeX.push_back(flat[i]), b.ShapeY.push_back(flat[i + 1]), b.ShapeZ.push_back(flat[i + 2]);
    b.RadiantRadius.push_back(0.f);
    PullColumns(bank, b, slot); // the radiating areas and the radiant radius (:313-337)
    return slot;
}

void TuneModalObject(ModalBank &b, uint32_t object, std::span<const float> freqs, std::span<const float> t60s, float radius_scale) {
    MeBank *bank = Mirror(b);
    Check(me_bank_tune_object(bank, object, freqs.data(), t60s.data(), uint32_t(std::min(freqs.size(), t60s.size())), radius_scale));
    b.DeflectionScale[object] = 1.f / (radius_scale * radius_scale * radius_scale);
    PullColumns(bank, b, object);
}

bool SetModalObjectShapes(ModalBank &b, uint32_t object, const ModalModes &modes) {
    const auto begin = b.ShapeOffset[object];
    const auto end = object + 1 < b.ShapeOffset.size() ? b.ShapeOffset[object + 1] : uint32_t(b.ShapeX.size());
    const auto count = uint32_t(modes.Freqs.size());
    if (b.ModeCount[object] != count || end - begin != count * modes.Shapes.size()) return false;
    const auto flat = Flatten(modes.Shapes);
    if (me_bank_set_object_shapes(Mirror(b), object, count, uint32_t(modes.Shapes.size()), flat.data()) != ME_OK) return false;
    for (size_t i = 0; i < flat.size(); i += 3) b.ShapeX[begin + i / 3] = flat[i], b.ShapeY[begin + i / 3] = flat[i + 1], b.ShapeZ[begin + i / 3] = flat[i + 2];
    return true;
}

std::optional<uint32_t> FindModalObject(const ModalBank &b, entt::entity e) {
    const auto it = std::ranges::find(b.Entities, e);
    return it != b.Entities.end() ? std::optional{uint32_t(std::ranges::distance(b.Entities.begin(), it))} : std::nullopt;
}

void InstallModalBank(ModalAudio &m, ModalBank &next) {
    MeBank *bank = BuildingBank(next);
    for (uint32_t slot = 0; slot < next.OutGain.size(); ++slot) Check(me_bank_set_gain(bank, slot, next.OutGain[slot], next.ListenerGain[slot]));
    Check(me_bank_install(bank)); // SoA upload to HBM; queued events become stale, as with FlushEvents (:277-289)
    auto &t = T();
    {
        std::scoped_lock lock{t.Mutex};
        t.Building.erase(&next);
        auto &live = t.Live[&m];
        me_bank_free(live);
        live = bank;
    }
    m.Live = std::make_unique<ModalBank>(std::move(next));
    m.FlushEvents.store(true, std::memory_order_relaxed);
    m.Published.store(m.Live.get(), std::memory_order_seq_cst);
    m.ActiveVoices.store(0, std::memory_order_relaxed);
    SurfaceInstallBank(m);
}

void EnqueueModalEvent(ModalAudio &m, const ModalEvent &event) {
    MeBank *bank = nullptr;
    {
        auto &t = T();
        std::scoped_lock lock{t.Mutex};
        if (const auto it = t.Live.find(&m); it != t.Live.end()) bank = it->second;
    }
    // Before the first install the reference queues into its own ring and the adopting render flushes it (:496-498): nothing to mirror.
    if (!bank) return;
    if (me_bank_enqueue(bank, reinterpret_cast<const MeModalEvent *>(&event)) == ME_QUEUE_FULL) ++m.EventsDropped; // :419-422
}

// Adds `frame_count` samples into `out`. The gains the main thread stores in place on the live host bank (SetModalOutGain,
// UpdateListenerGains: atomic_ref stores, AudioSystem.cpp:227-243) are forwarded before every block.
void RenderModal(ModalAudio &m, float *out, uint32_t frame_count) {
    MeBank *bank = nullptr;
    {
        auto &t = T();
        std::scoped_lock lock{t.Mutex};
        if (const auto it = t.Live.find(&m); it != t.Live.end()) bank = it->second;
    }
    if (!bank) return;
    const ModalBank &host = LiveBank(m);
    for (uint32_t slot = 0; slot < host.OutGain.size(); ++slot) me_bank_set_gain(bank, slot, host.OutGain[slot], host.ListenerGain[slot]);
    me_bank_set_click_gain(bank, m.ClickGain.load(std::memory_order_relaxed));
    me_bank_set_max_impacts(bank, m.MaxImpacts.load(std::memory_order_relaxed));
    if (me_bank_render(bank, out, frame_count) != ME_OK) return; // the audio path never throws to the device (AudioDevice.cpp:153-156)
    m.ActiveImpacts.store(me_bank_active_impacts(bank), std::memory_order_relaxed);
}
