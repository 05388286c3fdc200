/* me_modal.h — C ABI of the B200-native modal analysis / synthesis library (libme_modal.so).
 *
 * Drop-in boundary for MeshEditor's linear modal hot path. The reference has no FFI layer:
 * the boundary is two C++ free-function APIs, src/audio/ModalAudio.h:297-315 (synthesis) and
 * src/audio/mesh2modes.h:77-88 (analysis). Every entry point below names the reference function it replaces.
 * integration/mesh2modes_b200.cpp re-creates the reference's own analysis signatures on top of this ABI (a replacement
 * translation unit for src/audio/mesh2modes.cpp); INTEGRATION.md shows the synthesis-side calls.
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
/* Tuning front-end: RetuneModalObject (src/audio/AudioSystem.cpp:263-311) after its scene lookups - what the reference computes
 * between a stored modal model and TuneModalObject. Host-only float arithmetic, evaluated as the reference does. */
typedef struct MeRetune {
    float scale;        /* UniformScaleRatio (ContactScene.h:97-101): me_uniform_scale_ratio; > 0 */
    float fundamental;  /* ModalTuning::FundamentalFreq; <= 0 keeps the model's own first frequency */
    float t60_scale;    /* ModalTuning::T60Scale; 1 without a tuning */
    int has_alpha;      /* the object carries an AcousticMaterial ... */
    double alpha;       /* ... and this is its Rayleigh alpha: d' = alpha/2 + (d - alpha/2) / scale^2 */
    float modal_level;  /* ModalControls::ModalLevel */
    float gain;         /* ModalGain::Value; 1 without */
} MeRetune;
/* freqs * (fundamental / freqs[0]) / scale; T60 through the damping law above, times t60_scale; T60 <= 0 stays 0 (muted). */
MeStatus me_retune_modes(const float *freqs, const float *t60s, uint32_t n, const MeRetune *, float *out_freqs, float *out_t60s);
/* ModalOutGain (AudioSystem.cpp:221-224): modal_level * gain * scale^-2. */
float me_modal_out_gain(const MeRetune *);
/* UniformScaleRatio (ContactScene.h:97-101) over MeanScale (TransformMath.h:17-20): mean |world| / mean |baked| clamped to
 * 0.001..1000; 1 without a world transform (NULL) or with a zero baked scale. */
float me_uniform_scale_ratio(const float world_scale[3], const float baked_scale[3]);
/* UpdateListenerGains (AudioSystem.cpp:232-243): 1/r from the object's node, held at the 1 m mix level inside 1 m. */
float me_listener_gain(float distance);
/* MonitorFrames (AudioSystem.cpp:1177-1189), the stage ProcessAudio applies after RenderModal on the device path: pressure in Pa
 * to device units at 20 Pa full scale, in place, under a limiter whose gain follows the running peak envelope (instant attack,
 * 100 ms release). *envelope is MonitorLimiter::Envelope (AudioTypes.h:20-22), carried across calls; start it at 0. */
MeStatus me_monitor_frames(float *frames, uint64_t n, float sample_rate, float *envelope);
/* RetuneModalObject itself: me_retune_modes -> TuneModalObject(slot, ..., scale) -> OutGain[slot] = me_modal_out_gain
 * (the listener gain is left as it is). No-op for n == 0. Valid before and after install. */
MeStatus me_bank_retune_object(MeBank *, uint32_t slot, const float *freqs, const float *t60s, uint32_t n, const MeRetune *);
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
    uint32_t tensor_windows;        /* launch windows rendered in the tensor-core form (tcgen05 mix kernel) */
    uint32_t partial_rows;          /* partial mix rows the last window left for the mix kernel */
    float walk_kernel_ms;           /* tensor-core form: device time of the state walk kernel(s) */
    float tensor_mix_kernel_ms;     /* tensor-core form: device time of the tcgen05 mix kernel(s) */
    float host_plan_ms;             /* host time spent planning the spans (impact schedule, uploads) before their first launch */
    float pulse_kernels_ms;         /* device time of the force + pulse kernels */
} MeRenderStats;
MeStatus me_bank_last_render_stats(const MeBank *, MeRenderStats *out);
/* Scheduling knobs of the offline renderer: time_segments 0 = automatic. */
MeStatus me_bank_set_time_segments(MeBank *, uint32_t time_segments);
/* Which kernels render the free-running bank: 0 = automatic (tensor-core form for long offline spans, sample loop
 * otherwise), 1 = always the FP32 sample loop, 2 = the tensor-core form wherever the span allows it (block_frames a
 * multiple of 256 that divides 32768). Both forms implement RenderObjectFast (ModalAudio.cpp:86-147) to the same
 * 1e-5-of-peak bar. */
MeStatus me_bank_set_render_path(MeBank *, uint32_t path);

/* DealObjects (ModalAudio.cpp:430-461), the reference's split of the ringing objects between its renderers, used here to
 * split a bank's objects between GPUs (one renderer = one device): objects taken heaviest first, ties
 * by object index, each onto the renderer carrying least so far (the first such renderer); with one renderer everything
 * stays in bank order. costs[i] = modes x (1 + voices) of object i (`order.emplace_back(modes * (1 + voices), o)`, :444).
 * Writes owner[i] in [0, n_renderers) and, when `local_slot` is not NULL, the object's slot inside its owner's bank (each
 * renderer takes its objects in bank order, :459). Pure host function of its inputs: every rank computes the same deal. */
MeStatus me_deal_objects(const uint64_t *costs, uint32_t n_objects, uint32_t n_renderers, uint32_t *owner, uint32_t *local_slot);

/* ------------------------------------------------------------------------------------------------
 * Analysis: tet mesh + material -> modal model (reference: src/audio/mesh2modes.{h,cpp}).
 * ---------------------------------------------------------------------------------------------- */

/* AcousticMaterialProperties (src/audio/AcousticMaterialProperties.h:6-16). */
typedef struct MeMaterial {
    double density, young_modulus, poisson_ratio, alpha, beta;
} MeMaterial;

/* modal::SolverConfig (src/audio/mesh2modes.h:17-26) plus the two knobs this library adds. */
typedef struct MeSolverConfig {
    float min_mode_freq, max_mode_freq;   /* 20, 16000 */
    uint32_t num_modes, num_fem_modes;    /* 30, 45 */
    double tolerance, warm_tolerance;     /* 1e-8, 1e-4 */
    uint32_t max_restarts;                /* 100 */
    int32_t has_fundamental_freq;         /* std::optional<float> FundamentalFreq */
    float fundamental_freq;
    uint32_t element_order;               /* 2 = the reference's 10-node tets (parity mode); 1 = 4-node tets */
    int32_t device;                       /* CUDA ordinal the solve runs on */
} MeSolverConfig;
void me_solver_config_default(MeSolverConfig *);

/* JobMonitor (src/Job.h:13-19): progress written by the solve, cancelled polled between stages and restarts. */
typedef struct MeJobMonitor {
    volatile float progress;
    volatile int32_t cancelled;
} MeJobMonitor;

/* modal::SolveProfile (src/audio/mesh2modes.h:30-50): the same fields (seconds, counters), then device-side detail. */
typedef struct MeSolveProfile {
    double mass_props, quad_mesh, assemble, sample_excite, factorize, iterate, op_solve, extract;
    uint32_t dofs, stiffness_nonzeros, op_applications, restarts;
    double analyse;                 /* host symbolic analysis (ordering + structure), part of `factorize` */
    double factor_flops;            /* flops of the numeric factorisation, from the symbolic analysis */
    uint64_t factor_nonzeros;       /* scalars in the supernodal factor */
    float assemble_kernel_ms, factor_device_ms;
    uint32_t supernodes, levels, kernel_launches, tets_kept;
} MeSolveProfile;

/* MassProperties (src/audio/ContactModel.h:16-23). Orientation is a unit quaternion (w, x, y, z) of the principal axes. */
typedef struct MeMassProperties {
    double mass;
    float center_of_mass[3];
    float inertia_diagonal[3];
    float inertia_orientation[4];
} MeMassProperties;

typedef struct MeModalResult MeModalResult; /* modal::ModalResult (mesh2modes.h:52-62) */

/* modal::mesh2modes (src/audio/mesh2modes.h:77, mesh2modes.cpp:605-658).
 *   points_xyz [n_points][3] doubles, tets [n_tets][4] positively oriented (TetMesh, src/mesh/TetMesh.h:10-13);
 *   excite_xyz [n_excite][3] floats (SI), each sampled at its nearest tet point; baked_scale[3];
 *   seed_basis: a prior solve's eigenvector basis (seed_rows x seed_cols floats, column-major; SolveReuse::SeedBasis,
 *   mesh2modes.h:28-36) or NULL. When seed_rows equals this mesh's DOF count and seed_cols >= min(NumFemModes, n-1) the
 *   warm path runs: subspace iteration over NumFemModes + 15 columns to config->warm_tolerance
 *   (SubspaceIterate, mesh2modes.cpp:339-428, selected at :459-472); any other seed falls back to the cold solve;
 *   keep_basis: fill the result's basis (SolveReuse::KeepBasis).
 * Status mirrors the reference's failure modes. A cancel seen right after assembly leaves *out EMPTY (mesh2modes.cpp:616
 * returns `{}`). ME_NOT_CONVERGED, and a cancel seen later (inside ComputeModes, :462,479,490), leave *out with empty modes,
 * eigen summary and basis but WITH the mass properties, the profile and the excitation remap, as :655-657 still build
 * them. ME_NO_MODES keeps everything but the (empty) modes. ME_FACTOR_FAILED is the reference's std::runtime_error.
 * Calls may run concurrently from several host threads, on one device or on several (the reference's generation jobs are one
 * worker thread each, AudioSystem.cpp:830-862): every call works on a CUDA stream of its own, so one call's host stages overlap
 * the others' kernels; the results are those of the same calls made one after the other. */
MeStatus me_modal_solve(const double *points_xyz, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const MeMaterial *material,
                        const float *excite_xyz, uint32_t n_excite, const float baked_scale[3], const MeSolverConfig *config,
                        const float *seed_basis, uint32_t seed_rows, uint32_t seed_cols, int keep_basis, MeJobMonitor *monitor,
                        MeModalResult **out);
void me_modal_result_free(MeModalResult *);

/* ModalModes (src/audio/ModalModes.h:7-20). Arrays are owned by the result. */
uint32_t me_modal_result_mode_count(const MeModalResult *);        /* Freqs.size() */
uint32_t me_modal_result_point_count(const MeModalResult *);       /* Positions.size() == Shapes.size() */
const float *me_modal_result_freqs(const MeModalResult *);
const float *me_modal_result_t60s(const MeModalResult *);
const float *me_modal_result_shapes(const MeModalResult *);        /* [point][mode][3] */
const float *me_modal_result_positions(const MeModalResult *);     /* [point][3], node-local */
float me_modal_result_original_fundamental(const MeModalResult *); /* OriginalFundamentalFreq */
/* ModalResult::SamplePointOfExcitation: [n_excite] indices into positions. */
const uint32_t *me_modal_result_sample_point_of_excitation(const MeModalResult *, uint32_t *count);
/* ModalEigenSummary (src/audio/ModalEigenSummary.h:12-23): raw eigenpairs at the sample points. */
uint32_t me_modal_result_eigenpair_count(const MeModalResult *);
const double *me_modal_result_eigenvalues(const MeModalResult *);
const float *me_modal_result_summary_shapes(const MeModalResult *); /* [point][eigenpair][3] */
MeStatus me_modal_result_mass_properties(const MeModalResult *, MeMassProperties *out);
MeStatus me_modal_result_profile(const MeModalResult *, MeSolveProfile *out);
/* ModalResult::Basis (n x eigenpairs floats, column-major) when keep_basis was set, else NULL. */
const float *me_modal_result_basis(const MeModalResult *, uint32_t *rows, uint32_t *cols);

/* modal::PostprocessModes (mesh2modes.h:82, mesh2modes.cpp:515-588), host-only: eigenvalues[n_eigen] ascending,
 * shapes [n_points][n_eigen][3], positions [n_points][3]. Returns a result holding only the ModalModes part. */
MeStatus me_postprocess_modes(const double *eigenvalues, uint32_t n_eigen, const float *shapes, uint32_t n_points, float shape_scale,
                              const MeMaterial *material, const MeSolverConfig *config, const float *positions, MeModalResult **out);
/* modal::RescaleModes (mesh2modes.h:88, mesh2modes.cpp:590-603), host-only: re-derives the modes of `solved` (solved with
 * `solved_material`) under `material`. ME_BAD_ARG when the edit is not exactly scalable (the reference returns nullopt).
 * Like the reference, the eigen summary is left untouched: *out carries the SOLVED eigenvalues / summary shapes / mass
 * properties / excitation remap next to the rescaled modes, so it can be rescaled again (always against `solved_material`)
 * or archived with me_modal_file_serialize and the original solved material. */
MeStatus me_rescale_modes(const MeModalResult *solved, const MeMaterial *solved_material, const MeMaterial *material,
                          const MeSolverConfig *config, MeModalResult **out);
/* What the edit loop hands to me_rescale_modes (src/audio/AudioSystem.cpp:595-616). EffectiveModalMaterial: `props` unless the
 * object is an authoritative dynamic rigid body (body_mass > 0: its authored mass), in which case density = solved density *
 * body_mass / solve_mass and Young's modulus scales with it (E / rho kept). RescaledModes' rule for SolverConfig::
 * FundamentalFreq: a fundamental pinned at solve time (first frequency != OriginalFundamentalFreq > 0) stays pinned; returns
 * 1 and writes it, else 0. */
MeStatus me_effective_modal_material(const MeMaterial *props, const MeMaterial *solved, double solve_mass, double body_mass, MeMaterial *out);
int me_pinned_fundamental(const float *freqs, uint32_t n, float original_fundamental, float *fundamental);

/* --- Stages of the solve, exported for the parity tests and the roofline bench (inner operator concept of
 *     src/audio/CholeskyShiftInvert.h:18-23 and lib/spectra/.../SparseSymMatProd.h:83). ------------------------------ */
typedef struct MeFemSystem MeFemSystem;
typedef struct MeFemInfo {
    uint32_t tets_kept, node_count, dofs, nodes_per_element;
    uint64_t nnz_stiffness, nnz_mass;   /* stored lower-triangular scalars, Eigen layout */
    uint64_t node_blocks_lower, node_blocks_full;
    float assemble_kernel_ms;
    uint32_t kernel_launches;
} MeFemInfo;
/* FilterDegenerate + BuildQuadMesh + AssembleQuadratic on the device (mesh2modes.cpp:42-60, 246-264, 273-327). */
MeStatus me_fem_assemble(const double *points_xyz, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const MeMaterial *material,
                         uint32_t element_order, int device, MeFemSystem **out);
void me_fem_free(MeFemSystem *);
MeStatus me_fem_info(const MeFemSystem *, MeFemInfo *out);
MeStatus me_fem_get_element_nodes(MeFemSystem *, uint32_t *out /* [tets_kept][nodes_per_element] */);
/* The lower-triangular CSC the reference's Eigen matrices hold: which = 0 stiffness, 1 mass. */
MeStatus me_fem_get_csc(MeFemSystem *, int which, uint64_t *colptr /* [dofs+1] */, uint32_t *rowidx, double *values);
/* First-fit element colouring in element order (defined by oracle/modal.py greedy_colouring). */
MeStatus me_fem_colour_elements(MeFemSystem *, uint32_t *colours /* [tets_kept] */, uint32_t *n_colours);
/* y = A x (which = 0: K, 1: M), host vectors; `repeats` products are timed on the device (ms per product). */
MeStatus me_fem_spmv(MeFemSystem *, int which, const double *x, double *y, uint32_t repeats, float *ms_per_product);

typedef struct MeFactor MeFactor;
typedef struct MeFactorInfo {
    double analyse_seconds, factor_flops;
    float factor_device_ms, last_solve_device_ms;
    uint64_t factor_nonzeros;
    uint32_t supernodes, levels, dofs, kernel_launches;
} MeFactorInfo;
/* CholeskyShiftInvert::set_shift (CholeskyShiftInvert.cpp:26-46): analyse and factor K - sigma*M on the device. */
MeStatus me_factor_create(MeFemSystem *, double sigma, MeFactor **out);
void me_factor_free(MeFactor *);
/* perform_op / solve_panel (CholeskyShiftInvert.cpp:48-62): x = (K - sigma M)^-1 b, `width` column-major right-hand sides (host). */
MeStatus me_factor_solve(MeFactor *, const double *b, double *x, uint32_t width);
MeStatus me_factor_info(MeFactor *, MeFactorInfo *out);

/* Host-only: the symbolic analysis behind me_factor_create (geometric nested dissection + supernodes) on a node graph
 * given as full symmetric CSR (rowptr[node_count+1], col) with node coordinates xyz[node_count][3]. perm_out (may be NULL)
 * receives the elimination order (perm[new] = old). `violations` counts structural inconsistencies found by the
 * self-check (entries of the permuted matrix not covered by the supernodal structure, child structures not contained
 * in their parent's): 0 for a valid analysis. Needs no CUDA device. */
typedef struct MeSymbolicInfo {
    uint32_t supernodes, levels, max_panel_columns, max_panel_rows;
    uint64_t factor_nonzeros, update_tiles, panel_tiles, violations;
    double factor_flops, ordering_seconds, structure_seconds;
} MeSymbolicInfo;
MeStatus me_symbolic_analyse(uint32_t node_count, const uint32_t *rowptr, const uint32_t *col, const float *xyz, uint32_t *perm_out, MeSymbolicInfo *out);

/* FP64 issue-rate micro-benchmark (flop/s): mode 0 DFMA, 1 DMMA (mma.sync.m8n8k4.f64). */
MeStatus me_measure_fp64_rate(int device, int mode, int iters, double *flops_per_second);

/* FP32 FMA issue-rate micro-benchmark (the ceiling the resonator is bound by).
 * Returns FP32 lane-operations per second on `device`, measured with CUDA events over `iters` launches.
 * packed: 0 scalar FFMA; 1 Blackwell packed FFMA2 (fma.rn.f32x2); 2 / 3 FFMA2 and FFMA interleaved 1:1 / 1:2
 * (instructions), probing whether both FMA pipes run concurrently; 4 scalar FADD. */
MeStatus me_measure_fp32_fma_rate(int device, int packed, int iters, double *fma_per_second);

/* ------------------------------------------------------------------------------------------------
 * Strike front-end: contact dynamics -> the ModalEvent a strike enqueues.
 * Reference: src/audio/ContactModel.{h,cpp}, RecoilClickFilter (src/audio/ModalAudio.h:92-99) and the arithmetic of
 * TriggerModalStrike (src/audio/AudioSystem.cpp:400-465). Host-only, needs no CUDA device.
 * ---------------------------------------------------------------------------------------------- */

/* Striker (ContactModel.h:36-40): a capsule mallet; defaults steel, 0.01 m tip radius, 0.19 m long. */
typedef struct MeStriker {
    MeMaterial material;
    float tip_radius, length;
} MeStriker;
/* Impactor (ContactModel.h:44-48): the striking side of a Hertz contact. inv_mass 0 = immovable. */
typedef struct MeImpactor {
    MeMaterial material;
    double curvature; /* 1/m */
    double inv_mass;  /* 1/kg */
} MeImpactor;
/* ContactDynamics (ContactModel.h:28-32): mass, inverse inertia about the centre of mass (glm::mat3, column-major) and
 * the contact arm of every excitable vertex (borrowed for the call). */
typedef struct MeContactDynamics {
    double mass;
    float inverse_inertia[9];
    const float *contact_arm_xyz;
    uint32_t arm_count;
} MeContactDynamics;

double me_striker_mass(const MeStriker *);                              /* StrikerMass, ContactModel.cpp:10-13 */
MeStatus me_striker_impactor(const MeStriker *, MeImpactor *out);       /* StrikerImpactor, :15 */
MeStatus me_inverse_inertia_tensor(const MeMassProperties *, float out[9]); /* InverseInertiaTensor, :17-24 */
/* ReducedContactMass (:27-38); 0 when the index is out of range or the object has no mass. */
double me_reduced_contact_mass(const MeContactDynamics *, uint32_t excitable_index, const float impact_direction[3], const MeImpactor *);
/* EstimateContactTime (:76-114): seconds, clamped to [2e-5, 5e-2]; MinContactTime on degenerate input. */
double me_estimate_contact_time(const MeContactDynamics *, uint32_t excitable_index, const float impact_direction[3], double contact_speed, const MeMaterial *object_material,
                                double object_curvature, double nominal_area, const MeImpactor *, double scale_ratio, double combined_roughness);
/* Hertz contact constants (ContactModel.cpp:40-66). Arguments by constant:
 *   INV_EFFECTIVE_MODULUS(a, b); COMBINED_CURVATURE(x = k1, y = k2); STIFFNESS(x = 1/E*, y = 1/R*);
 *   PATCH_RADIUS(x = N, y = 1/E*, z = 1/R*); STATIC_PENETRATION(x = N, y = k); SATURATION_PENETRATION(x = 1/R*, y = A0);
 *   PUNCH_STIFFNESS(x = 1/E*, y = A0). */
typedef enum MeContactConstant {
    ME_CONTACT_INV_EFFECTIVE_MODULUS = 0, ME_CONTACT_COMBINED_CURVATURE = 1, ME_CONTACT_STIFFNESS = 2, ME_CONTACT_PATCH_RADIUS = 3,
    ME_CONTACT_STATIC_PENETRATION = 4, ME_CONTACT_SATURATION_PENETRATION = 5, ME_CONTACT_PUNCH_STIFFNESS = 6
} MeContactConstant;
double me_contact_constant(MeContactConstant which, const MeMaterial *a, const MeMaterial *b, double x, double y, double z);
/* RecoilClickFilter (ModalAudio.h:92-99): {B0, A1, A2}; zeros when radius or mass is not positive. */
void me_recoil_click_filter(double radius, double volume, double mass, double sample_rate, float b0_a1_a2[3]);

/* One strike as TriggerModalStrike sees it once the scene lookups are done (AudioSystem.cpp:400-465). */
typedef struct MeStrike {
    uint32_t object;              /* bank slot (FindModalObject) */
    uint32_t excitable_index;     /* excitation position; also where a mallet's contact time is evaluated */
    float force, contact_speed;
    float direction[3];           /* node-local; normalised here when is_collision (physics->Direction), else taken as is */
    int32_t is_collision;         /* PhysicsStrike present: force is the true contact impulse */
    uint32_t resultant_index;     /* collision: sample point nearest the manifold's load-weighted centre */
    const MeContactDynamics *dynamics; /* NULL (or elastic NULL): 1e-4 s contact, no click */
    const MeMaterial *elastic;    /* material of the struck surface */
    MeImpactor impactor;          /* StrikerImpactor(...) for a mallet, the colliding body's for a collision */
    double curvature;             /* struck surface's contribution to 1/R* where the strike lands */
    double nominal_area;          /* collision only: area the two faces share, m^2 */
    double scale_ratio;           /* UniformScaleRatio */
    double roughness;             /* combined rms asperity height, m */
    double displaced_volume;      /* DisplacedVolume (:247-254); 0 = none, the click corner then uses radiant_radius */
    float radiant_radius;         /* bank.RadiantRadius[slot] */
    float sample_rate;            /* bank.SampleRate */
} MeStrike;
MeStatus me_make_strike_event(const MeStrike *, MeModalEvent *out);
/* UpdateContactDynamics (src/audio/ContactDynamics.cpp:19-46) past its registry lookups. `resolved`: the solve's mass properties, or
 * the authoritative rigid body's (then mass_scale = 1); mass_scale: ModalDensityRatio (:13-17), material density / solved density.
 * Outputs fill an MeContactDynamics: mass * mass_scale, InverseInertiaTensor / mass_scale, arms = (position - centre of mass) *
 * MeanScale(baked_scale) for every sample point ([n_positions][3]). */
MeStatus me_contact_dynamics(const MeMassProperties *resolved, double mass_scale, const float *positions_xyz, uint32_t n_positions, const float baked_scale[3], double *mass,
                             float inverse_inertia[9], float *arms_xyz);
/* TiltAlongNormal (AudioSystem.cpp:359-371): the strike direction of a manual hit - the excited vertex's unit normal tilted toward
 * the surface by a joystick position in the unit disk (centre: along the normal; rim: 90 degrees, in the tangent plane). */
void me_tilt_along_normal(const float normal[3], const float joystick[2], float out[3]);
/* SphereEquivalentCurvature (AudioSystem.cpp:379-380): mean curvature, 1/m, of the solid sphere with this mass and density (the
 * colliding body's MeImpactor::curvature); 0 for an immovable body. */
double me_sphere_equivalent_curvature(double density, double inv_mass);

/* ------------------------------------------------------------------------------------------------
 * Model interchange: the data formats either side of a modal solve. Host-only.
 * ---------------------------------------------------------------------------------------------- */

/* What ModalModelData (src/audio/ModalModelFile.h:13-20) holds beyond a solve's result: ModalModes::Vertices / Indices /
 * BakedScale, the display TetMeshData, and the rest of ModalEigenSummary. Arrays are borrowed for a serialize call and owned
 * by the MeModalFile handle after a parse. */
typedef struct MeModalFileExtras {
    const uint32_t *vertices;          uint32_t n_vertices;         /* ModalModes::Vertices */
    const uint32_t *indices;           uint32_t n_indices;          /* ModalModes::Indices */
    float baked_scale[3];                                           /* ModalModes::BakedScale */
    const float *tet_positions_xyz;    uint32_t n_tet_positions;    /* TetMeshData::Positions */
    const uint32_t *tet_edge_indices;  uint32_t n_tet_edge_indices; /* TetMeshData::EdgeIndices */
    MeMaterial solved_material;                                     /* ModalEigenSummary::SolvedMaterial ... */
    float solved_min_mode_freq, solved_max_mode_freq;
    uint32_t solved_num_modes;
    uint64_t tet_inputs_hash;
    const uint32_t *solved_vertices;   uint32_t n_solved_vertices;
} MeModalFileExtras;
typedef struct MeModalFile MeModalFile;

/* The bytes SaveModalModelFile writes (ModalModelFile.cpp:15-22: zpp::bits archive of ModalModelData{Modes, Mass, Tets,
 * Summary}), byte for byte. *bytes is malloc'ed: release with me_bytes_free. */
MeStatus me_modal_file_serialize(const MeModalResult *, const MeModalFileExtras *, uint8_t **bytes, uint64_t *size);
/* LoadModalModelFile (ModalModelFile.cpp:52-58): ME_BAD_ARG on truncated or inconsistent data. The result answers the
 * me_modal_result_* accessors (modes, mass properties, eigenvalues, summary shapes); the rest through me_modal_file_extras. */
MeStatus me_modal_file_parse(const uint8_t *bytes, uint64_t size, MeModalResult **result, MeModalFile **file);
MeStatus me_modal_file_extras(const MeModalFile *, MeModalFileExtras *out);
void me_modal_file_free(MeModalFile *);
void me_bytes_free(void *);
/* The JSON MeshEditorModalSolve prints (tests/ModalSolveTool.cpp:84-123), which glTF_PhysicalAudio embeds as a
 * KHR_audio_rigid_bodies modal model: frequencies, decayRates (= ln 1000 / T60), positions, mode-major shapes, the mesh
 * triangles relabelled onto the sample points (merged corners drop the triangle), mass, centerOfMass, inertiaDiagonal.
 * Numbers are the shortest text that round-trips (std::format "{}"). *json is malloc'ed: release with me_bytes_free. */
MeStatus me_modal_solve_json(const MeModalResult *, const uint32_t *triangle_indices, uint32_t n_triangle_indices, char **json);

/* ------------------------------------------------------------------------------------------------
 * Rendered audio out: the file WriteWav produces (src/audio/AudioSystem.cpp:1244-1250 over src/audio/WavWriter.h). Host-only.
 * ---------------------------------------------------------------------------------------------- */

/* Mono 32-bit IEEE-float RIFF/WAVE ("fmt " format 3, "fact", "data"). normalize_max > 0 scales the frames by normalize_max /
 * max(frames), WriteWav's rule. *bytes is malloc'ed: release with me_bytes_free. */
MeStatus me_wav_encode(const float *frames, uint64_t n, uint32_t sample_rate, float normalize_max, uint8_t **bytes, uint64_t *size);
/* Reads such a file back (also mono 16-bit PCM, as recorded impacts come); unknown chunks are skipped. *frames is malloc'ed. */
MeStatus me_wav_decode(const uint8_t *bytes, uint64_t size, float **frames, uint64_t *n, uint32_t *sample_rate);

/* ------------------------------------------------------------------------------------------------
 * Impact spectrum analysis (src/audio/AudioSystem.cpp:492-560): the fundamental of a recorded impact, which LaunchModalSolve
 * passes to the solve as MeSolverConfig::fundamental_freq (:821-829). Host-only.
 * ---------------------------------------------------------------------------------------------- */

/* ComputeFft (:553-558): frames 30 .. sample_rate/16 under a Blackman-Harris window, transformed; complex_re_im receives
 * n_real/2 + 1 (re, im) pairs in float. With complex_re_im NULL only *n_real is written (size query). ME_BAD_ARG when the
 * recording is shorter than sample_rate/16 frames. */
MeStatus me_impact_spectrum(const float *frames, uint64_t n_frames, uint32_t sample_rate, float *complex_re_im, uint64_t *n_real);
/* EstimateFundamentalFrequency (:522-550): dB spectrum, noise threshold = median of the upper half + 15 dB, first local maximum
 * from max(50 Hz, bin 15) that stands >= 10 dB above the mean of its +-15 bins; *hz = bin * sample_rate / n_real in whole hertz.
 * Returns 1 and writes *hz, or 0 (the reference's nullopt). */
int me_estimate_fundamental_from_spectrum(const float *complex_re_im, uint64_t n_real, uint32_t sample_rate, float *hz);
/* Both steps: what the reference computes from a recorded impact before a solve. */
int me_estimate_fundamental(const float *frames, uint64_t n_frames, uint32_t sample_rate, float *hz);

/* ------------------------------------------------------------------------------------------------
 * Generation-job glue (first slice): what the reference's modal generation job does either side of
 * mesh2modes apart from simplifying and tetrahedralizing the surface (src/audio/AudioSystem.cpp:838-862). Host-only.
 * Every *out array is malloc'ed (one element at least, so never NULL on ME_OK): release with me_bytes_free.
 * ---------------------------------------------------------------------------------------------- */

/* DesiredSolveVertices (AudioSystem.cpp:667-671) without copied sound vertices: clamp(requested, 1, num_vertices) excitation
 * vertices, i * num_vertices / count. */
MeStatus me_desired_solve_vertices(uint32_t requested, uint32_t num_vertices, uint32_t **out, uint32_t *n_out);
/* SampleSurfaceTriangles (AudioSystem.cpp:701-746): the mesh's triangulation collapsed onto the excitation vertices. Every
 * mesh vertex takes the excitation vertex (its index in `excitation_vertices`) it reaches in the fewest edges; a triangle whose
 * corners took three different ones contributes a triangle; one triangle per distinct point set, ordered by the sorted set, in
 * the winding first seen (UniqueSampleTriangles, :675-695). Empty with fewer than 3 excitation vertices or no triangle. */
MeStatus me_sample_surface_triangles(const uint32_t *triangle_indices, uint32_t n_triangle_indices, uint32_t vertex_count, const uint32_t *excitation_vertices, uint32_t n_excitation_vertices,
                                     uint32_t **out, uint32_t *n_out);
/* CompactExcitationVertices (AudioSystem.cpp:750-757): the first excitation vertex of every sample point, in sample point
 * order, from me_modal_result_sample_point_of_excitation -> ModalModes::Vertices. */
MeStatus me_compact_excitation_vertices(const uint32_t *vertices, uint32_t n_vertices, const uint32_t *sample_point_of, uint32_t n_sample_point_of, uint32_t **out, uint32_t *n_out);
/* RelabelSampleTriangles (AudioSystem.cpp:761-769): the sample surface after the solve merged positions -> ModalModes::Indices.
 * Empty when sample_point_of is empty; ME_BAD_ARG for a corner without a sample point. */
MeStatus me_relabel_sample_triangles(const uint32_t *triangles, uint32_t n_triangle_indices, const uint32_t *sample_point_of, uint32_t n_sample_point_of, uint32_t **out, uint32_t *n_out);
/* BuildTetMeshData (src/mesh/Tets.cpp:268-293): the TetMeshData a `.modal` file stores beside the modes (MeModalFileExtras
 * tet_positions_xyz / tet_edge_indices): points divided by the node scale, as floats [n_points][3], and the distinct tet edges
 * as (low, high) corner pairs in ascending order. */
MeStatus me_build_tet_mesh_data(const double *points_xyz, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const float scale[3], float **positions_xyz, uint32_t **edge_indices,
                                uint32_t *n_edge_indices);

/* Unit-test entry of the tensor-core mix (tensor_mix.cuh): out[row][frame] = sum over the 4096 reduction elements of
 * each of the row's groups_per_row consecutive groups of power * state, operands given as host images of the stage layout documented in tensor_mix.cuh
 * (powers: groups*256 stages of two FP16 images of 256 x 16, stage layout; states: [tiles][groups][blocks_per_tile][4096] FP32 - the entry scales and
 * splits them into the FP16 hi / lo rows the walk kernel writes, with the same helper kernels).
 * blocks_per_tile is 128; frames <= tiles*blocks_per_tile*256. `milliseconds` (may be NULL) receives the
 * kernel time of the last of `repeats` launches. */
MeStatus me_debug_tensor_mix(int device, const float *powers, const float *states, uint32_t groups, uint32_t groups_per_row, uint32_t tiles, uint32_t blocks_per_tile, uint32_t frames,
                             uint32_t repeats,
                             float *out, float *milliseconds);

#ifdef __cplusplus
}
#endif
#endif /* ME_MODAL_H */
