/* me_modal.h — C ABI of the B200-native modal analysis / synthesis library (libme_modal.so).
 *
 * Drop-in boundary for MeshEditor's linear modal hot path. The reference has no FFI layer:
 * the boundary is two C++ free-function APIs, src/audio/ModalAudio.h:297-315 (synthesis) and
 * src/audio/mesh2modes.h:77-88 (analysis). Every entry point below names the reference function it replaces.
 * A header-only C++ shim (include/me_modal_shim.hpp) re-creates the reference's own signatures on top of this ABI.
 *
 * Conventions
 *   - plain pointers and sizes only; inputs are borrowed for the duration of the call;
 *   - every function returns an MeStatus; no exception crosses the ABI; me_last_error() gives the text
 *     of the calling thread's last failure;
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point returns ME_CUDA_ERROR;
 *   - handles are opaque and freed with the matching *_free.
 */
#ifndef ME_MODAL_H
#define ME_MODAL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum MeStatus {
    ME_OK = 0,
    ME_BAD_ARG = 1,
    ME_CUDA_ERROR = 2,
    ME_CANCELLED = 3,       /* JobMonitor::Cancelled(): the reference returns an empty result (mesh2modes.cpp:462,616) */
    ME_NOT_CONVERGED = 4,   /* Spectra CompInfo != Successful: empty result (mesh2modes.cpp:490) */
    ME_NO_MODES = 5,        /* no eigenfrequency at or above MinModeFreq: empty ModalModes (mesh2modes.cpp:548) */
    ME_FACTOR_FAILED = 6,   /* the reference throws std::runtime_error (CholeskyShiftInvert.cpp:44) */
    ME_QUEUE_FULL = 7,      /* EnqueueModalEvent drops and counts (ModalAudio.cpp:419-422) */
    ME_OUT_OF_MEMORY = 8
} MeStatus;

const char *me_last_error(void);
/* Library/ABI version and the compute capability it was built for ("sm_100a"). */
const char *me_build_info(void);
/* Number of CUDA devices visible; 0 when none (compute entry points then fail with ME_CUDA_ERROR). */
int me_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * Synthesis: the modal resonator bank (reference: src/audio/ModalAudio.{h,cpp}).
 * ---------------------------------------------------------------------------------------------- */

/* Layout-identical to the reference's ModalEvent (ModalAudio.h:28-37); 48 bytes. */
typedef struct MeModalEvent {
    uint32_t kind;        /* 0 = Impact, 1 = Silence (ModalEventKind, ModalAudio.h:22-25) */
    uint32_t object;      /* object slot in the bank */
    uint32_t ex_pos;      /* excitation position index */
    float jx, jy, jz;     /* node-local impulse vector */
    float pulse_step;     /* per-sample phase increment of the raised-cosine contact pulse */
    float pulse_gamma;    /* contact pulse amplitude */
    float accel_amp;      /* scales the unit-sum pulse to the click filter's input force, N */
    float click_b0, click_a1, click_a2; /* the click's coupled recoil filter */
} MeModalEvent;

typedef struct MeBank MeBank; /* ModalAudio + ModalBank (ModalAudio.h:101-167, 245-289) */

/* ModalAudio{} with ModalBank::SampleRate. `device` is the CUDA ordinal the bank lives on. */
MeStatus me_bank_create(float sample_rate, int device, MeBank **out);
void me_bank_free(MeBank *);

/* AddModalObject (ModalAudio.cpp:291-338). shapes_xyz: [point][mode][3] mass-normalised mode shapes;
 * positions_xyz: [point][3]; indices: triangles over the points (may be NULL/0: no radiating surface).
 * The slot is appended to the bank being built; it becomes audible after me_bank_install. */
MeStatus me_bank_add_object(MeBank *, uint32_t n_modes, uint32_t n_points, const float *shapes_xyz,
                            const float *positions_xyz, const uint32_t *indices, uint32_t n_indices,
                            uint32_t *slot_out);
/* TuneModalObject (ModalAudio.cpp:340-393). Valid before and after install (live retune). */
MeStatus me_bank_tune_object(MeBank *, uint32_t slot, const float *freqs, const float *t60s, uint32_t n,
                             float radius_scale);
/* SetModalObjectShapes (ModalAudio.cpp:395-410). ME_BAD_ARG when the mode/shape layout differs (reference: false). */
MeStatus me_bank_set_object_shapes(MeBank *, uint32_t slot, uint32_t n_modes, uint32_t n_points,
                                   const float *shapes_xyz);
/* ModalBank::OutGain / ListenerGain (ModalAudio.h:129-131), written in place like the reference's atomic_ref stores. */
MeStatus me_bank_set_gain(MeBank *, uint32_t slot, float out_gain, float listener_gain);
/* ModalAudio::ClickGain / MaxImpacts (ModalAudio.h:262-263). */
MeStatus me_bank_set_click_gain(MeBank *, float click_gain);
MeStatus me_bank_set_max_impacts(MeBank *, uint32_t max_impacts);

/* InstallModalBank (ModalAudio.cpp:277-289): publishes the built bank to the device (SoA upload to HBM) and
 * flags queued events as stale; the next render drops them, exactly like the reference's FlushEvents. */
MeStatus me_bank_install(MeBank *);
/* EnqueueModalEvent (ModalAudio.cpp:417-425). ME_QUEUE_FULL when the 256-entry ring is full (event dropped, counted). */
MeStatus me_bank_enqueue(MeBank *, const MeModalEvent *);
/* RenderModal (ModalAudio.cpp:486-590): ADDS `frames` mono samples into the HOST buffer `out`.
 * Drains the event ring first; impacts start at the first frame of this block. */
MeStatus me_bank_render(MeBank *, float *out_accumulate, uint32_t frames);

/* Offline render of a whole timeline in one call: equivalent to the loop
 *     for (b = 0; b*block_frames < total_frames; ++b) { enqueue events with event_frames == b*block_frames;
 *                                                       RenderModal(out + b*block_frames, block_frames); }
 * (tests/ModalBench.h:76-80, src/audio/AudioSystem.cpp:1155-1159). event_frames must be ascending multiples of
 * block_frames. `out` (HOST, total_frames floats) is ADDED into. The bank's state, impacts and event ring
 * carry over from and to the streaming calls. */
MeStatus me_bank_render_offline(MeBank *, const MeModalEvent *events, const uint64_t *event_frames,
                                uint32_t n_events, uint64_t total_frames, uint32_t block_frames,
                                float *out_accumulate);
/* Same, with the result left in DEVICE memory (`out_device`, total_frames floats, overwritten, on the bank's
 * device) and the work enqueued on `cuda_stream` (a cudaStream_t; NULL = the bank's own stream) without a final
 * synchronise. Used by the multi-GPU mixdown (NCCL all-reduce of the per-rank mixes) and by the throughput bench. */
MeStatus me_bank_render_offline_device(MeBank *, const MeModalEvent *events, const uint64_t *event_frames,
                                       uint32_t n_events, uint64_t total_frames, uint32_t block_frames,
                                       float *out_device, void *cuda_stream);

/* Bank introspection (main-thread reads of LiveBank(), ModalAudio.h:292). */
typedef enum MeModeColumn {
    ME_COL_COEFF_RE = 0, ME_COL_COEFF_IM = 1, ME_COL_STATE_RE = 2, ME_COL_STATE_IM = 3,
    ME_COL_RADIATION_GAIN = 4, ME_COL_RADIATION_AREA = 5, ME_COL_OUT_PHASE_IM = 6, ME_COL_OUT_PHASE_RE = 7,
    ME_COL_DEFLECTION_GAIN = 8, ME_COL_QUAD_COMPLIANCE = 9, ME_COL_QUAD_DRIVE_SCALE = 10
} MeModeColumn;
uint32_t me_bank_object_count(const MeBank *);
uint32_t me_bank_mode_total(const MeBank *);       /* sum of ModeCount over objects */
uint32_t me_bank_active_impacts(const MeBank *);   /* ModalAudio::ActiveImpacts */
uint64_t me_bank_events_dropped(const MeBank *);   /* ModalAudio::EventsDropped */
/* Copies one per-mode column (objects concatenated, ModeOffset order) into out[me_bank_mode_total()].
 * State columns are read back from the device. */
MeStatus me_bank_get_mode_column(MeBank *, MeModeColumn which, float *out);
/* ModeOffset / ModeCount / TunedModeCount of one object. */
MeStatus me_bank_get_object_layout(const MeBank *, uint32_t slot, uint32_t *mode_offset, uint32_t *mode_count,
                                   uint32_t *tuned_mode_count, float *radiant_radius);

/* ModalBank::LiveModeCount / Ringing of one object (ModalAudio.h:126-127,133), read back from the device:
 * the audible prefix the culling of ModalAudio.cpp:139-146 left, in modes, and whether the object holds state. */
MeStatus me_bank_get_object_status(MeBank *, uint32_t slot, uint32_t *live_mode_count, uint32_t *ringing);

/* Kernel statistics of the last render call, for the bench (gpu_launches, kernel time from CUDA events). */
typedef struct MeRenderStats {
    uint32_t kernel_launches;       /* launches of this library's kernels in the last render call */
    float resonator_kernel_ms;      /* device time of the resonator kernel(s) (CUDA events on the render stream) */
    float total_device_ms;          /* first launch to last launch of the call */
    uint64_t mode_samples;          /* sum over objects of TunedModeCount x frames */
    uint64_t h2d_bytes, d2h_bytes;  /* host<->device bytes moved by the call */
    uint32_t time_segments;         /* segments of the block-parallel scan along time (1 = sequential in time) */
    uint32_t scan_fallbacks;        /* windows re-rendered sequentially because culling fell inside them */
} MeRenderStats;
MeStatus me_bank_last_render_stats(const MeBank *, MeRenderStats *out);
/* Scheduling knobs of the offline renderer: time_segments 0 = automatic. */
MeStatus me_bank_set_time_segments(MeBank *, uint32_t time_segments);

/* FP32 FMA issue-rate micro-benchmark (the ceiling the resonator is bound by).
 * Returns FP32 lane-operations per second on `device`, measured with CUDA events over `iters` launches.
 * packed: 0 scalar FFMA; 1 Blackwell packed FFMA2 (fma.rn.f32x2); 2 / 3 FFMA2 and FFMA interleaved 1:1 / 1:2
 * (instructions), probing whether both FMA pipes run concurrently; 4 scalar FADD. */
MeStatus me_measure_fp32_fma_rate(int device, int packed, int iters, double *fma_per_second);

#ifdef __cplusplus
}
#endif
#endif /* ME_MODAL_H */
