// Reference-side binding of the resonator bank: the calls a MeshEditor maintainer swaps in for the six ModalAudio.h entry
// points AudioSystem.cpp uses (src/audio/ModalAudio.h:297-315), implemented on top of libme_modal.so's C ABI
// (include/me_modal.h). The reference's ModalBank stays what AudioSystem.cpp reads and writes in place (OutGain, ListenerGain,
// Entities, the UI's columns); a device bank mirrors it slot for slot, and RenderModal becomes one me_bank_render call.
// Compiled inside the reference tree beside ModalAudio.cpp (it includes the reference's own headers), so it is not built in
// this repository; `g++ -fsyntax-only` against /root/reference is part of tests/test_host_cpu.py where the tree exists.
//
//   AudioSystem.cpp call                          ->  b200:: call
//   AddModalObject(next, e, modes)        :320    ->  b200::AddModalObject(m, next, e, modes)
//   TuneModalObject(b, slot, f, t, scale) :308    ->  b200::TuneModalObject(m, b, slot, f, t, scale)
//   SetModalObjectShapes(b, slot, modes)          ->  b200::SetModalObjectShapes(m, b, slot, modes)
//   InstallModalBank(m, next)             :325    ->  b200::InstallModalBank(m, next)
//   EnqueueModalEvent(m, event)           :448    ->  b200::EnqueueModalEvent(m, event)
//   RenderModal(m, output, frame_count)   :1199   ->  b200::RenderModal(m, output, frame_count)
#include "audio/ModalAudio.h"
#include "audio/ModalModes.h"

#include "me_modal.h"

#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

static_assert(sizeof(ModalEvent) == sizeof(MeModalEvent), "MeModalEvent mirrors ModalEvent field for field");

namespace b200 {
namespace {
// The device banks of one ModalAudio: the one being built beside `next`, and the published one.
struct DeviceBanks {
    MeBank *Building{nullptr}, *Live{nullptr};
    const ModalBank *BuildingFor{nullptr};
    int Device{0};
};
std::mutex TableMutex;
std::unordered_map<const ModalAudio *, DeviceBanks> Table;

DeviceBanks &Of(const ModalAudio &m) {
    std::scoped_lock lock{TableMutex};
    return Table[&m];
}
void Check(MeStatus status) {
    if (status != ME_OK && status != ME_QUEUE_FULL) throw std::runtime_error(me_last_error());
}
// The device bank that mirrors host bank `b`: the published one when `b` is live, else the one under construction.
MeBank *Mirror(ModalAudio &m, const ModalBank &b) {
    auto &d = Of(m);
    if (m.Live.get() == &b) return d.Live;
    if (d.BuildingFor != &b) { // a new `next`: RebuildModalBank starts from an empty bank every time (AudioSystem.cpp:314-326)
        me_bank_free(d.Building);
        d.Building = nullptr;
        Check(me_bank_create(b.SampleRate, d.Device, &d.Building));
        d.BuildingFor = &b;
    }
    return d.Building;
}
std::vector<float> Flatten(const std::vector<std::vector<vec3>> &shapes) { // [point][mode] vec3 -> [point][mode][3]
    std::vector<float> flat;
    for (const auto &point : shapes)
        for (const auto &s : point) flat.insert(flat.end(), {s.x, s.y, s.z});
    return flat;
}
} // namespace

void UseDevice(ModalAudio &m, int cuda_device) { Of(m).Device = cuda_device; }

uint32_t AddModalObject(ModalAudio &m, ModalBank &next, entt::entity e, const ModalModes &modes) {
    const uint32_t slot = ::AddModalObject(next, e, modes); // host columns: what the editor displays and retunes from
    const auto flat = Flatten(modes.Shapes);
    uint32_t device_slot = 0;
    Check(me_bank_add_object(Mirror(m, next), uint32_t(modes.Freqs.size()), uint32_t(modes.Shapes.size()), flat.data(), modes.Positions.empty() ? nullptr : &modes.Positions[0].x,
                             modes.Indices.data(), uint32_t(modes.Indices.size()), &device_slot));
    if (device_slot != slot) throw std::logic_error("device bank out of step with the host bank");
    return slot;
}

void TuneModalObject(ModalAudio &m, ModalBank &b, uint32_t slot, std::span<const float> freqs, std::span<const float> t60s, float radius_scale = 1.f) {
    ::TuneModalObject(b, slot, freqs, t60s, radius_scale);
    Check(me_bank_tune_object(Mirror(m, b), slot, freqs.data(), t60s.data(), uint32_t(std::min(freqs.size(), t60s.size())), radius_scale));
}

bool SetModalObjectShapes(ModalAudio &m, ModalBank &b, uint32_t slot, const ModalModes &modes) {
    if (!::SetModalObjectShapes(b, slot, modes)) return false;
    const auto flat = Flatten(modes.Shapes);
    return me_bank_set_object_shapes(Mirror(m, b), slot, uint32_t(modes.Freqs.size()), uint32_t(modes.Shapes.size()), flat.data()) == ME_OK;
}

void InstallModalBank(ModalAudio &m, ModalBank &next) {
    auto &d = Of(m);
    MeBank *bank = Mirror(m, next);
    for (uint32_t slot = 0; slot < next.OutGain.size(); ++slot) Check(me_bank_set_gain(bank, slot, next.OutGain[slot], next.ListenerGain[slot]));
    Check(me_bank_install(bank)); // SoA upload to HBM; queued events become stale, as with FlushEvents
    ::InstallModalBank(m, next); // the host bank the editor keeps reading
    me_bank_free(d.Live);
    d.Live = bank, d.Building = nullptr, d.BuildingFor = nullptr;
}

void EnqueueModalEvent(ModalAudio &m, const ModalEvent &event) {
    MeBank *bank = Of(m).Live;
    if (!bank) return;
    if (me_bank_enqueue(bank, reinterpret_cast<const MeModalEvent *>(&event)) == ME_QUEUE_FULL) ++m.EventsDropped; // ModalAudio.cpp:419-422
}

// Adds `frame_count` samples into `out`, like the reference. The gains the main thread stores in place on the live host bank
// (SetModalOutGain, UpdateListenerGains: atomic_ref stores, AudioSystem.cpp:227-243) are forwarded before every block.
void RenderModal(ModalAudio &m, float *out, uint32_t frame_count) {
    MeBank *bank = Of(m).Live;
    if (!bank) return;
    const ModalBank &host = LiveBank(m);
    for (uint32_t slot = 0; slot < host.OutGain.size(); ++slot) me_bank_set_gain(bank, slot, host.OutGain[slot], host.ListenerGain[slot]);
    me_bank_set_click_gain(bank, m.ClickGain.load(std::memory_order_relaxed));
    me_bank_set_max_impacts(bank, m.MaxImpacts.load(std::memory_order_relaxed));
    if (me_bank_render(bank, out, frame_count) != ME_OK) return; // the audio path never throws to the device (AudioDevice.cpp:153-156)
    m.ActiveImpacts.store(me_bank_active_impacts(bank), std::memory_order_relaxed);
}

// The whole offline timeline in one call (the loop at AudioSystem.cpp:1155-1159 / tests/ModalBench.h:76-80): events[i] takes
// effect at frame event_frames[i], a multiple of block_frames.
void RenderModalOffline(ModalAudio &m, std::span<const ModalEvent> events, std::span<const uint64_t> event_frames, uint64_t total_frames, uint32_t block_frames, float *out) {
    MeBank *bank = Of(m).Live;
    if (!bank) return;
    Check(me_bank_render_offline(bank, reinterpret_cast<const MeModalEvent *>(events.data()), event_frames.data(), uint32_t(events.size()), total_frames, block_frames, out));
}

void Release(ModalAudio &m) {
    std::scoped_lock lock{TableMutex};
    if (const auto it = Table.find(&m); it != Table.end()) {
        me_bank_free(it->second.Building), me_bank_free(it->second.Live);
        Table.erase(it);
    }
}
} // namespace b200
