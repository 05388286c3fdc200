// The C ABI (include/me_modal.h): thin extern "C" shims over me::Bank (and, later, the solver).
#include "bank.h"
#include "tensor_mix.cuh"
#include "common.h"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

namespace me {
namespace {
thread_local std::string g_last_error;
}
void SetLastError(const char *fmt, ...) {
    char buffer[512];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buffer, sizeof buffer, fmt, args);
    va_end(args);
    g_last_error = buffer;
}
const char *LastError() { return g_last_error.c_str(); }

namespace {
// One pool stream per device, created on first use; the pool keeps freed memory (release threshold = max).
cudaStream_t PoolStream() {
    static thread_local int cached_device = -1;
    static thread_local cudaStream_t cached_stream = nullptr;
    static cudaStream_t streams[64]{};
    static std::mutex mutex;
    int device = 0;
    ME_CUDA(cudaGetDevice(&device));
    if (device == cached_device) return cached_stream;
    std::lock_guard<std::mutex> lock(mutex);
    if (device < 0 || device >= 64) Fail(ME_BAD_ARG, "device ordinal %d out of range", device);
    if (!streams[device]) {
        cudaMemPool_t pool;
        ME_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t threshold = UINT64_MAX;
        ME_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
        ME_CUDA(cudaStreamCreateWithFlags(&streams[device], cudaStreamNonBlocking));
    }
    cached_device = device;
    cached_stream = streams[device];
    return cached_stream;
}
} // namespace

// Device blocks released by the library are kept, by size, for the next request of (about) that size. A modal solve takes
// and returns ~20 GB in a few dozen buffers, the same sizes every time the same mesh is solved again (the edit loop, a bench
// step); handing those straight back to cudaFreeAsync / cudaMallocAsync left the stream-ordered pool re-carving its arena
// on every solve, and one request in five or so stalled for 0.3 - 1.1 s (measured: profiles/r02_solve.md). With the cache a
// steady-state solve makes no allocator call at all. At most ME_POOL_CACHE_GB (default 64) are held per device; beyond that
// the largest idle blocks go back to the driver.
namespace {
struct BlockCache {
    std::mutex Mutex;
    std::unordered_map<void *, size_t> Sizes;   // every live block handed out by PoolAllocate
    std::multimap<size_t, void *> Idle;         // released blocks, by size
    size_t IdleBytes{0};
};
BlockCache &CacheOf(int device) {
    static BlockCache caches[64];
    if (device < 0 || device >= 64) Fail(ME_BAD_ARG, "device ordinal %d out of range", device);
    return caches[device];
}
size_t CacheLimit() {
    static const size_t limit = [] {
        const char *env = std::getenv("ME_POOL_CACHE_GB");
        return size_t(env ? std::max(0.0, std::atof(env)) : 64.0) << 30;
    }();
    return limit;
}
} // namespace

void *PoolAllocate(size_t bytes) {
    bytes = std::max<size_t>((bytes + 255) & ~size_t(255), 256);
    int device = 0;
    ME_CUDA(cudaGetDevice(&device));
    BlockCache &cache = CacheOf(device);
    {
        std::lock_guard<std::mutex> lock(cache.Mutex);
        // the smallest idle block that fits, if it is not wastefully larger (an eighth, or 1 MB for the small ones)
        const auto it = cache.Idle.lower_bound(bytes);
        if (it != cache.Idle.end() && it->first <= bytes + std::max<size_t>(bytes / 8, size_t(1) << 20)) {
            void *ptr = it->second;
            cache.IdleBytes -= it->first;
            cache.Sizes[ptr] = it->first;
            cache.Idle.erase(it);
            return ptr;
        }
    }
    void *ptr = nullptr;
    const cudaStream_t s = PoolStream();
    cudaError_t err = cudaMallocAsync(&ptr, bytes, s);
    if (err == cudaErrorMemoryAllocation) {
        // make room: everything idle goes back to the driver, then once more
        cudaGetLastError();
        std::lock_guard<std::mutex> lock(cache.Mutex);
        for (auto &[size, idle] : cache.Idle) cudaFreeAsync(idle, s);
        cache.Idle.clear();
        cache.IdleBytes = 0;
        cudaStreamSynchronize(s);
        err = cudaMallocAsync(&ptr, bytes, s);
    }
    if (err != cudaSuccess) Fail(err == cudaErrorMemoryAllocation ? ME_OUT_OF_MEMORY : ME_CUDA_ERROR, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(err));
    ME_CUDA(cudaStreamSynchronize(s)); // usable from any stream from here on
    std::lock_guard<std::mutex> lock(cache.Mutex);
    cache.Sizes[ptr] = bytes;
    return ptr;
}
void PoolFree(void *ptr) {
    // Called from destructors: never throws. The device-wide wait is what cudaFree did implicitly: nothing in flight uses the
    // block any more, so whoever gets it next may use it on any stream.
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return;
    cudaPointerAttributes attr{};
    const int owner = (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeDevice) ? attr.device : device;
    if (owner != device) cudaSetDevice(owner);
    cudaDeviceSynchronize();
    try {
        BlockCache &cache = CacheOf(owner);
        std::vector<void *> evicted;
        bool keep = false;
        {
            std::lock_guard<std::mutex> lock(cache.Mutex);
            const auto known = cache.Sizes.find(ptr);
            const size_t size = known != cache.Sizes.end() ? known->second : 0;
            if (known != cache.Sizes.end()) cache.Sizes.erase(known);
            if (size && size <= CacheLimit()) {
                while (cache.IdleBytes + size > CacheLimit() && !cache.Idle.empty()) { // the largest idle blocks make room
                    const auto last = std::prev(cache.Idle.end());
                    cache.IdleBytes -= last->first;
                    evicted.push_back(last->second);
                    cache.Idle.erase(last);
                }
                cache.Idle.emplace(size, ptr);
                cache.IdleBytes += size;
                keep = true;
            }
        }
        const cudaStream_t s = PoolStream();
        for (void *e : evicted) cudaFreeAsync(e, s);
        if (!keep) cudaFreeAsync(ptr, s);
    } catch (...) {
    }
    if (owner != device) cudaSetDevice(device);
}
} // namespace me

struct MeBank {
    me::Bank Impl;
    MeBank(float sample_rate, int device) : Impl(sample_rate, device) {}
};

using me::Fail;
using me::Guard;
namespace me {
// tuning.cpp
void RetuneModes(const float *freqs, const float *t60s, uint32_t n, const MeRetune &rt, float *out_freqs, float *out_t60s);
float ModalOutGain(const MeRetune &rt);
} // namespace me
using me::ModalOutGain;
using me::RetuneModes;

extern "C" {

const char *me_last_error(void) { return me::LastError(); }
const char *me_build_info(void) { return "libme_modal 0.1 (C ABI 1), CUDA sm_100a, no CPU fallback"; }
int me_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

MeStatus me_bank_create(float sample_rate, int device, MeBank **out) {
    return Guard([&] {
        if (!out) Fail(ME_BAD_ARG, "null out pointer");
        *out = nullptr;
        if (!(sample_rate > 0)) Fail(ME_BAD_ARG, "sample_rate must be positive");
        *out = new MeBank(sample_rate, device);
    });
}
void me_bank_free(MeBank *b) { delete b; }

#define ME_BANK_CALL(b, body)                          \
    Guard([&] {                                        \
        if (!(b)) Fail(ME_BAD_ARG, "null bank handle"); \
        body;                                          \
    })

MeStatus me_bank_add_object(MeBank *b, uint32_t n_modes, uint32_t n_points, const float *shapes_xyz, const float *positions_xyz, const uint32_t *indices, uint32_t n_indices, uint32_t *slot_out) {
    return ME_BANK_CALL(b, {
        const auto slot = b->Impl.AddObject(n_modes, n_points, shapes_xyz, positions_xyz, indices, n_indices);
        if (slot_out) *slot_out = slot;
    });
}
MeStatus me_bank_tune_object(MeBank *b, uint32_t slot, const float *freqs, const float *t60s, uint32_t n, float radius_scale) {
    return ME_BANK_CALL(b, b->Impl.TuneObject(slot, freqs, t60s, n, radius_scale));
}
MeStatus me_bank_set_object_shapes(MeBank *b, uint32_t slot, uint32_t n_modes, uint32_t n_points, const float *shapes_xyz) {
    return ME_BANK_CALL(b, {
        if (!shapes_xyz) Fail(ME_BAD_ARG, "null shapes");
        b->Impl.SetObjectShapes(slot, n_modes, n_points, shapes_xyz);
    });
}
MeStatus me_bank_set_gain(MeBank *b, uint32_t slot, float out_gain, float listener_gain) { return ME_BANK_CALL(b, b->Impl.SetGain(slot, out_gain, listener_gain)); }
MeStatus me_bank_retune_object(MeBank *b, uint32_t slot, const float *freqs, const float *t60s, uint32_t n, const MeRetune *rt) {
    return ME_BANK_CALL(b, {
        if (!rt || (n && (!freqs || !t60s))) Fail(ME_BAD_ARG, "null argument");
        if (!(rt->scale > 0)) Fail(ME_BAD_ARG, "scale must be positive");
        if (!n) return; // RetuneModalObject returns before touching the slot when the model has no modes
        std::vector<float> tuned_freqs(n);
        std::vector<float> tuned_t60s(n);
        RetuneModes(freqs, t60s, n, *rt, tuned_freqs.data(), tuned_t60s.data());
        b->Impl.TuneObject(slot, tuned_freqs.data(), tuned_t60s.data(), n, rt->scale);
        b->Impl.SetOutGain(slot, ModalOutGain(*rt));
    });
}
MeStatus me_bank_set_click_gain(MeBank *b, float g) { return ME_BANK_CALL(b, b->Impl.SetClickGain(g)); }
MeStatus me_bank_set_max_impacts(MeBank *b, uint32_t n) { return ME_BANK_CALL(b, b->Impl.SetMaxImpacts(n)); }
MeStatus me_bank_set_time_segments(MeBank *b, uint32_t n) { return ME_BANK_CALL(b, b->Impl.SetTimeSegments(n)); }
MeStatus me_bank_set_render_path(MeBank *b, uint32_t path) {
    return ME_BANK_CALL(b, {
        if (path > 2) Fail(ME_BAD_ARG, "render path must be 0, 1 or 2");
        b->Impl.SetRenderPath(path);
    });
}
MeStatus me_bank_install(MeBank *b) { return ME_BANK_CALL(b, b->Impl.Install()); }
MeStatus me_deal_objects(const uint64_t *costs, uint32_t n_objects, uint32_t n_renderers, uint32_t *owner, uint32_t *local_slot) {
    return Guard([&] { me::DealObjects(costs, n_objects, n_renderers, owner, local_slot); });
}

MeStatus me_bank_enqueue(MeBank *b, const MeModalEvent *e) {
    MeStatus queued = ME_OK;
    const MeStatus s = ME_BANK_CALL(b, {
        if (!e) Fail(ME_BAD_ARG, "null event");
        queued = b->Impl.Enqueue(*e);
    });
    return s != ME_OK ? s : queued;
}

MeStatus me_bank_render(MeBank *b, float *out, uint32_t frames) {
    return ME_BANK_CALL(b, b->Impl.RenderTimeline(nullptr, nullptr, 0, frames, frames ? frames : 1, out, false, nullptr, true));
}
MeStatus me_bank_render_offline(MeBank *b, const MeModalEvent *events, const uint64_t *event_frames, uint32_t n_events, uint64_t total_frames, uint32_t block_frames, float *out) {
    return ME_BANK_CALL(b, b->Impl.RenderTimeline(events, event_frames, n_events, total_frames, block_frames, out, false, nullptr, true));
}
MeStatus me_bank_render_offline_device(MeBank *b, const MeModalEvent *events, const uint64_t *event_frames, uint32_t n_events, uint64_t total_frames, uint32_t block_frames, float *out_device, void *cuda_stream) {
    return ME_BANK_CALL(b, b->Impl.RenderTimeline(events, event_frames, n_events, total_frames, block_frames, out_device, true, static_cast<cudaStream_t>(cuda_stream), cuda_stream == nullptr));
}

uint32_t me_bank_object_count(const MeBank *b) { return b ? b->Impl.ObjectCount() : 0; }
uint32_t me_bank_mode_total(const MeBank *b) { return b ? b->Impl.ModeTotal() : 0; }
uint32_t me_bank_active_impacts(const MeBank *b) { return b ? b->Impl.ActiveImpacts() : 0; }
uint64_t me_bank_events_dropped(const MeBank *b) { return b ? b->Impl.EventsDroppedCount() : 0; }
MeStatus me_bank_get_mode_column(MeBank *b, MeModeColumn which, float *out) { return ME_BANK_CALL(b, b->Impl.GetModeColumn(which, out)); }
MeStatus me_bank_get_object_layout(const MeBank *b, uint32_t slot, uint32_t *mode_offset, uint32_t *mode_count, uint32_t *tuned_mode_count, float *radiant_radius) {
    return ME_BANK_CALL(b, b->Impl.GetObjectLayout(slot, mode_offset, mode_count, tuned_mode_count, radiant_radius));
}
MeStatus me_bank_get_object_status(MeBank *b, uint32_t slot, uint32_t *live_mode_count, uint32_t *ringing) {
    return ME_BANK_CALL(b, b->Impl.GetObjectStatus(slot, live_mode_count, ringing));
}
MeStatus me_bank_last_render_stats(const MeBank *b, MeRenderStats *out) {
    return ME_BANK_CALL(b, {
        if (!out) Fail(ME_BAD_ARG, "null out");
        *out = const_cast<MeBank *>(b)->Impl.LastStats();
    });
}

MeStatus me_measure_fp32_fma_rate(int device, int packed, int iters, double *fma_per_second) {
    return Guard([&] {
        if (!fma_per_second || iters <= 0) Fail(ME_BAD_ARG, "bad arguments");
        ME_CUDA(cudaSetDevice(device));
        *fma_per_second = me::MeasureFmaRate(packed, iters);
    });
}

MeStatus me_debug_tensor_mix(int device, const float *powers, const float *states, uint32_t groups, uint32_t groups_per_row, uint32_t tiles, uint32_t blocks_per_tile, uint32_t frames, uint32_t repeats,
                             float *out, float *milliseconds) {
    return Guard([&] {
        if (!powers || !states || !out || groups == 0 || tiles == 0 || repeats == 0 || groups_per_row == 0 || groups % groups_per_row) Fail(ME_BAD_ARG, "bad arguments");
        if (frames > uint64_t(tiles) * blocks_per_tile * me::kTmBlock) Fail(ME_BAD_ARG, "frames exceed the tiles");
        ME_CUDA(cudaSetDevice(device));
        const size_t np = size_t(groups) * me::kTmStagesPerGroup * me::TmPowerStageFloats();
        const size_t ns = size_t(tiles) * groups * me::TmStateTileFloats(blocks_per_tile);
        me::DeviceBuffer<float> dp, ds, dplanes, dscale, dout;
        dp.Upload(powers, np, nullptr), ds.Upload(states, ns, nullptr), dout.Reserve(size_t(groups / groups_per_row) * frames);
        dscale.Reserve(size_t(tiles) * groups * me::TmScaleTileFloats(blocks_per_tile));
        dplanes.Reserve(ns);
        me::LaunchStateScaleKernel(ds.Ptr, tiles * groups, blocks_per_tile, dscale.Ptr, nullptr); // (in the bank the walk kernel writes both as it goes)
        me::LaunchStateSplitKernel(ds.Ptr, tiles * groups, blocks_per_tile, dscale.Ptr, dplanes.Ptr, nullptr);
        const me::TensorMixPlan plan{.Groups = groups, .StagesPerRow = groups_per_row * me::kTmStagesPerGroup, .Tiles = tiles, .BlocksPerTile = blocks_per_tile, .Frames = frames, .Powers = dp.Ptr, .States = dplanes.Ptr, .Scales = dscale.Ptr, .Partial = dout.Ptr};
        cudaEvent_t a, b;
        ME_CUDA(cudaEventCreate(&a));
        ME_CUDA(cudaEventCreate(&b));
        for (uint32_t r = 0; r < repeats; ++r) {
            if (r + 1 == repeats) ME_CUDA(cudaEventRecord(a, nullptr));
            me::LaunchTensorMixKernel(plan, nullptr);
        }
        ME_CUDA(cudaEventRecord(b, nullptr));
        ME_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        ME_CUDA(cudaEventElapsedTime(&ms, a, b));
        cudaEventDestroy(a), cudaEventDestroy(b);
        if (milliseconds) *milliseconds = ms;
        ME_CUDA(cudaMemcpy(out, dout.Ptr, size_t(groups / groups_per_row) * frames * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

} // extern "C"
