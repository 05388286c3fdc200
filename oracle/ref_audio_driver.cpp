// TEST INFRASTRUCTURE ONLY (oracle). Never linked into the product library.
//
// extern "C" wrapper around the UNMODIFIED reference resonator bank
// (/root/reference/src/audio/ModalAudio.{h,cpp}). The reference sources are compiled where they
// lie by oracle/Makefile; only this driver lives in the repo. Output: oracle/_ref/libme_ref_audio.so.
//
// It mirrors the harness the reference's own tests use (tests/ModalBench.h:47-81 `ModalScene`):
// AddModalObject + TuneModalObject + OutGain, InstallModalBank, one discard block, then
// EnqueueModalEvent / RenderModal.
#include "action/SerializeGlm.h"
#include "audio/ContactModel.h"
#include "audio/ModalModelFile.h"
#include "audio/ModalAudio.h"
#include "audio/ModalModes.h"

#include <entt/entity/entity.hpp>

#include <cstring>
#include <memory>
#include <vector>

namespace {
struct RefScene {
    ModalAudio Audio;
    ModalBank Next;
    bool Installed{false};
};

ModalModes MakeModes(uint32_t n_modes, uint32_t n_points, const float *freqs, const float *t60s, const float *shapes, const float *positions, const uint32_t *indices, uint32_t n_indices) {
    ModalModes modes;
    modes.Freqs.assign(freqs, freqs + n_modes);
    modes.T60s.assign(t60s, t60s + n_modes);
    modes.Shapes.resize(n_points);
    modes.Positions.resize(n_points);
    for (uint32_t p = 0; p < n_points; ++p) {
        modes.Shapes[p].resize(n_modes);
        for (uint32_t k = 0; k < n_modes; ++k) {
            const float *s = shapes + (size_t(p) * n_modes + k) * 3;
            modes.Shapes[p][k] = vec3{s[0], s[1], s[2]};
        }
        if (positions) modes.Positions[p] = vec3{positions[3 * p], positions[3 * p + 1], positions[3 * p + 2]};
    }
    if (indices) modes.Indices.assign(indices, indices + n_indices);
    return modes;
}
ModalBank &BankOf(RefScene &s) { return s.Installed ? LiveBank(s.Audio) : s.Next; }
} // namespace

extern "C" {
// Layout-identical to the reference's ModalEvent (ModalAudio.h:28-37), 48 bytes.
struct RefEvent {
    uint32_t Kind, Object, ExPos;
    float Jx, Jy, Jz, PulseStep, PulseGamma, AccelAmp, ClickB0, ClickA1, ClickA2;
};
static_assert(sizeof(RefEvent) == sizeof(ModalEvent));

void *ref_scene_create(float sample_rate, uint32_t renderers) {
    auto *s = new RefScene;
    s->Audio.RenderPool.SetSize(renderers);
    s->Next.SampleRate = sample_rate;
    return s;
}
void ref_scene_free(void *h) { delete static_cast<RefScene *>(h); }

// shapes: [point][mode][3]; positions: [point][3]; indices: triangles over the points.
uint32_t ref_scene_add_object(void *h, uint32_t n_modes, uint32_t n_points, const float *freqs, const float *t60s, const float *shapes, const float *positions, const uint32_t *indices, uint32_t n_indices, float out_gain, float radius_scale) {
    auto &s = *static_cast<RefScene *>(h);
    const auto modes = MakeModes(n_modes, n_points, freqs, t60s, shapes, positions, indices, n_indices);
    auto &b = s.Next;
    const auto slot = AddModalObject(b, entt::entity{uint32_t(b.Entities.size())}, modes);
    TuneModalObject(b, slot, modes.Freqs, modes.T60s, radius_scale);
    b.OutGain[slot] = out_gain;
    return slot;
}
void ref_scene_retune(void *h, uint32_t slot, const float *freqs, const float *t60s, uint32_t n, float radius_scale) {
    auto &s = *static_cast<RefScene *>(h);
    TuneModalObject(BankOf(s), slot, {freqs, n}, {t60s, n}, radius_scale);
}
void ref_scene_set_gain(void *h, uint32_t slot, float out_gain, float listener_gain) {
    auto &b = BankOf(*static_cast<RefScene *>(h));
    b.OutGain[slot] = out_gain;
    b.ListenerGain[slot] = listener_gain;
}
void ref_scene_set_click_gain(void *h, float g) { static_cast<RefScene *>(h)->Audio.ClickGain.store(g); }
void ref_scene_set_max_impacts(void *h, uint32_t n) { static_cast<RefScene *>(h)->Audio.MaxImpacts.store(n); }

// InstallModalBank, then one discard block as ModalScene does (tests/ModalBench.h:64-69).
void ref_scene_install(void *h, uint32_t discard_frames) {
    auto &s = *static_cast<RefScene *>(h);
    InstallModalBank(s.Audio, s.Next);
    s.Installed = true;
    if (discard_frames) {
        std::vector<float> discard(discard_frames, 0.f);
        RenderModal(s.Audio, discard.data(), discard_frames);
    }
}
void ref_scene_enqueue(void *h, const RefEvent *e) {
    ModalEvent ev;
    std::memcpy(&ev, e, sizeof ev);
    EnqueueModalEvent(static_cast<RefScene *>(h)->Audio, ev);
}
// RenderModal adds into `out`.
void ref_scene_render(void *h, float *out, uint32_t frames) { RenderModal(static_cast<RefScene *>(h)->Audio, out, frames); }

uint32_t ref_scene_mode_total(void *h) { return uint32_t(BankOf(*static_cast<RefScene *>(h)).CoeffRe.size()); }
uint32_t ref_scene_object_count(void *h) { return uint32_t(BankOf(*static_cast<RefScene *>(h)).Entities.size()); }
uint32_t ref_scene_active_impacts(void *h) { return uint32_t(BankOf(*static_cast<RefScene *>(h)).Impacts.size()); }
uint64_t ref_scene_events_dropped(void *h) { return static_cast<RefScene *>(h)->Audio.EventsDropped; }

// Per-mode columns, concatenated over objects. which: 0 CoeffRe 1 CoeffIm 2 StateRe 3 StateIm 4 RadiationGain
// 5 RadiationArea 6 OutPhaseIm 7 OutPhaseRe 8 DeflectionGain 9 QuadCompliance 10 QuadDriveScale
void ref_scene_get_mode_column(void *h, uint32_t which, float *out) {
    auto &b = BankOf(*static_cast<RefScene *>(h));
    const std::vector<float> *cols[]{&b.CoeffRe, &b.CoeffIm, &b.StateRe, &b.StateIm, &b.RadiationGain, &b.RadiationArea, &b.OutPhaseIm, &b.OutPhaseRe, &b.DeflectionGain, &b.QuadCompliance, &b.QuadDriveScale};
    if (which < std::size(cols)) std::memcpy(out, cols[which]->data(), cols[which]->size() * sizeof(float));
}
// Per-object columns. which: 0 ModeOffset 1 ModeCount 2 TunedModeCount 3 LiveModeCount 4 Ringing
void ref_scene_get_object_column_u32(void *h, uint32_t which, uint32_t *out) {
    auto &b = BankOf(*static_cast<RefScene *>(h));
    const auto n = b.Entities.size();
    for (size_t o = 0; o < n; ++o) {
        switch (which) {
            case 0: out[o] = b.ModeOffset[o]; break;
            case 1: out[o] = b.ModeCount[o]; break;
            case 2: out[o] = b.TunedModeCount[o]; break;
            case 3: out[o] = b.LiveModeCount[o]; break;
            case 4: out[o] = b.Ringing[o]; break;
            default: out[o] = 0;
        }
    }
}
// which: 0 RadiantRadius 1 DeflectionScale 2 OutGain 3 ListenerGain
void ref_scene_get_object_column_f32(void *h, uint32_t which, float *out) {
    auto &b = BankOf(*static_cast<RefScene *>(h));
    const std::vector<float> *cols[]{&b.RadiantRadius, &b.DeflectionScale, &b.OutGain, &b.ListenerGain};
    if (which < std::size(cols)) std::memcpy(out, cols[which]->data(), cols[which]->size() * sizeof(float));
}

// The recoil click filter the reference's own test builds its events with (ModalAudio.h:92-99).
void ref_click_filter(double radius, double volume, double mass, double sample_rate, float *b0_a1_a2) {
    const auto f = RecoilClickFilter(radius, volume, mass, sample_rate);
    b0_a1_a2[0] = f.B0;
    b0_a1_a2[1] = f.A1;
    b0_a1_a2[2] = f.A2;
}

// ---- Strike front-end (src/audio/ContactModel.{h,cpp}): the unmodified reference functions behind plain arguments. ----
// material: {Density, YoungModulus, PoissonRatio, Alpha, Beta}; mass_props: {Mass} + com[3], inertia_diagonal[3], quat wxyz.
static AcousticMaterialProperties RefMaterial(const double *m) { return {m[0], m[1], m[2], m[3], m[4]}; }
static Striker RefStriker(const double *material, float tip_radius, float length) {
    Striker s;
    s.Material.Properties = RefMaterial(material);
    s.TipRadius = tip_radius;
    s.Length = length;
    return s;
}
static ContactDynamics RefDynamics(double mass, const float *inverse_inertia, const float *arms, uint32_t n_arms) {
    ContactDynamics d;
    d.Mass = mass;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) d.InverseInertia[c][r] = inverse_inertia[3 * c + r];
    for (uint32_t i = 0; i < n_arms; ++i) d.ContactArm.push_back(vec3{arms[3 * i], arms[3 * i + 1], arms[3 * i + 2]});
    return d;
}
double ref_striker_mass(const double *material, float tip_radius, float length) { return StrikerMass(RefStriker(material, tip_radius, length)); }
void ref_striker_impactor(const double *material, float tip_radius, float length, double *curvature_invmass) {
    const auto imp = StrikerImpactor(RefStriker(material, tip_radius, length));
    curvature_invmass[0] = imp.Curvature, curvature_invmass[1] = imp.InvMass;
}
// UpdateContactDynamics (src/audio/ContactDynamics.cpp:19-46) past its registry lookups: `resolved` and `mass_scale` are what the
// lookups leave, `modes` the model; the statements between are the reference's own (cut out at build time, oracle/Makefile).
#include "mean_scale.inc"
uint32_t ref_contact_dynamics(double mass, const float *com, const float *inertia_diagonal, const float *quat_wxyz, double mass_scale, const float *positions, uint32_t n,
                              const float *baked, double *out_mass, float *out_inverse_inertia9, float *out_arms) {
    MassProperties resolved;
    resolved.Mass = mass;
    resolved.CenterOfMass = vec3{com[0], com[1], com[2]};
    resolved.InertiaDiagonal = vec3{inertia_diagonal[0], inertia_diagonal[1], inertia_diagonal[2]};
    resolved.InertiaOrientation = glm::quat{quat_wxyz[0], quat_wxyz[1], quat_wxyz[2], quat_wxyz[3]};
    ModalModes model;
    for (uint32_t i = 0; i < n; ++i) model.Positions.emplace_back(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
    model.BakedScale = vec3{baked[0], baked[1], baked[2]};
    const ModalModes *modes = &model;
#include "contact_dynamics.inc"
    *out_mass = cd.Mass;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) out_inverse_inertia9[3 * c + r] = cd.InverseInertia[c][r];
    for (uint32_t i = 0; i < n; ++i) out_arms[3 * i] = cd.ContactArm[i].x, out_arms[3 * i + 1] = cd.ContactArm[i].y, out_arms[3 * i + 2] = cd.ContactArm[i].z;
    return uint32_t(cd.ContactArm.size());
}
void ref_inverse_inertia(double mass, const float *inertia_diagonal, const float *quat_wxyz, float *out9) {
    MassProperties mp;
    mp.Mass = mass;
    mp.InertiaDiagonal = vec3{inertia_diagonal[0], inertia_diagonal[1], inertia_diagonal[2]};
    mp.InertiaOrientation = glm::quat{quat_wxyz[0], quat_wxyz[1], quat_wxyz[2], quat_wxyz[3]};
    const glm::mat3 m = InverseInertiaTensor(mp);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) out9[3 * c + r] = m[c][r];
}
double ref_reduced_contact_mass(double mass, const float *inverse_inertia, const float *arms, uint32_t n_arms, uint32_t i, const float *dir, const double *imp_material, double imp_curvature, double imp_inv_mass) {
    const Impactor imp{RefMaterial(imp_material), imp_curvature, imp_inv_mass};
    return ReducedContactMass(RefDynamics(mass, inverse_inertia, arms, n_arms), i, vec3{dir[0], dir[1], dir[2]}, imp);
}
double ref_estimate_contact_time(double mass, const float *inverse_inertia, const float *arms, uint32_t n_arms, uint32_t i, const float *dir, double contact_speed, const double *object_material,
                                 double object_curvature, double nominal_area, const double *imp_material, double imp_curvature, double imp_inv_mass, double scale_ratio, double combined_roughness) {
    const Impactor imp{RefMaterial(imp_material), imp_curvature, imp_inv_mass};
    return EstimateContactTime(RefDynamics(mass, inverse_inertia, arms, n_arms), i, vec3{dir[0], dir[1], dir[2]}, contact_speed, RefMaterial(object_material), object_curvature, nominal_area, imp, scale_ratio,
                               combined_roughness);
}
// Hertz constants (ContactModel.cpp:40-66): which = 0 InvEffectiveModulus(a,b as materials), else scalar helpers.
double ref_inv_effective_modulus(const double *a, const double *b) { return InvEffectiveModulus(RefMaterial(a), RefMaterial(b)); }
double ref_contact_scalar(int which, double x, double y, double z) {
    switch (which) {
        case 0: return CombinedCurvature(x, y);
        case 1: return ContactStiffness(x, y);
        case 2: return ContactPatchRadius(x, y, z);
        case 3: return StaticPenetration(x, y);
        case 4: return SaturationPenetration(x, y);
        default: return PunchStiffness(x, y);
    }
}

// ---- Model interchange: the reference's .modal bytes (src/audio/ModalModelFile.cpp:15-22 Serialize, :52-58 load) ----
// Flat arguments -> ModalModelData -> zpp::bits (the reference's own struct definitions and ADL hooks). Returns the byte
// count (0 on failure); `out` may be NULL to query the size.
struct RefModalFlat {
    uint32_t n_modes, n_points, n_vertices, n_indices, n_tet_positions, n_tet_edges, n_eigen, n_solved_vertices;
    const float *freqs, *t60s, *shapes /*[point][mode][3]*/, *positions;
    const uint32_t *vertices, *indices;
    float original_fundamental, baked_scale[3];
    double mass;
    float com[3], inertia[3], quat_wxyz[4];
    const float *tet_positions;
    const uint32_t *tet_edges;
    const double *eigenvalues;
    const float *summary_shapes /*[point][eigenpair][3]*/;
    double material[5];
    float min_freq, max_freq;
    uint32_t num_modes;
    uint64_t tet_hash;
    const uint32_t *solved_vertices;
};
static ModalModelData RefModalData(const RefModalFlat &f) {
    ModalModelData d;
    d.Modes.Freqs.assign(f.freqs, f.freqs + f.n_modes);
    d.Modes.T60s.assign(f.t60s, f.t60s + f.n_modes);
    d.Modes.Shapes.resize(f.n_points);
    d.Summary.Shapes.resize(f.n_points);
    for (uint32_t p = 0; p < f.n_points; ++p) {
        for (uint32_t k = 0; k < f.n_modes; ++k) d.Modes.Shapes[p].push_back(vec3{f.shapes[(size_t(p) * f.n_modes + k) * 3], f.shapes[(size_t(p) * f.n_modes + k) * 3 + 1], f.shapes[(size_t(p) * f.n_modes + k) * 3 + 2]});
        for (uint32_t k = 0; k < f.n_eigen; ++k)
            d.Summary.Shapes[p].push_back(vec3{f.summary_shapes[(size_t(p) * f.n_eigen + k) * 3], f.summary_shapes[(size_t(p) * f.n_eigen + k) * 3 + 1], f.summary_shapes[(size_t(p) * f.n_eigen + k) * 3 + 2]});
        d.Modes.Positions.push_back(vec3{f.positions[3 * p], f.positions[3 * p + 1], f.positions[3 * p + 2]});
    }
    d.Modes.Vertices.assign(f.vertices, f.vertices + f.n_vertices);
    d.Modes.Indices.assign(f.indices, f.indices + f.n_indices);
    d.Modes.OriginalFundamentalFreq = f.original_fundamental;
    d.Modes.BakedScale = vec3{f.baked_scale[0], f.baked_scale[1], f.baked_scale[2]};
    d.Mass.Mass = f.mass;
    d.Mass.CenterOfMass = vec3{f.com[0], f.com[1], f.com[2]};
    d.Mass.InertiaDiagonal = vec3{f.inertia[0], f.inertia[1], f.inertia[2]};
    d.Mass.InertiaOrientation = glm::quat{f.quat_wxyz[0], f.quat_wxyz[1], f.quat_wxyz[2], f.quat_wxyz[3]};
    for (uint32_t i = 0; i < f.n_tet_positions; ++i) d.Tets.Positions.push_back(vec3{f.tet_positions[3 * i], f.tet_positions[3 * i + 1], f.tet_positions[3 * i + 2]});
    d.Tets.EdgeIndices.assign(f.tet_edges, f.tet_edges + f.n_tet_edges);
    d.Summary.Eigenvalues.assign(f.eigenvalues, f.eigenvalues + f.n_eigen);
    d.Summary.SolvedMaterial = RefMaterial(f.material);
    d.Summary.SolvedMinModeFreq = f.min_freq, d.Summary.SolvedMaxModeFreq = f.max_freq;
    d.Summary.SolvedNumModes = f.num_modes;
    d.Summary.TetInputsHash = size_t(f.tet_hash);
    d.Summary.SolvedVertices.assign(f.solved_vertices, f.solved_vertices + f.n_solved_vertices);
    return d;
}
uint64_t ref_modal_file_serialize(const RefModalFlat *flat, uint8_t *out, uint64_t capacity) {
    ModalModelData d = RefModalData(*flat);
    std::vector<std::byte> bytes;
    zpp::bits::out archive{bytes};
    if (zpp::bits::failure(archive(d))) return 0;
    if (out && capacity >= bytes.size()) std::memcpy(out, bytes.data(), bytes.size());
    return bytes.size();
}
// Parses bytes with the reference's loader arithmetic and reports whether they decode to exactly `flat`'s content.
int ref_modal_file_roundtrip_equal(const RefModalFlat *flat, const uint8_t *bytes, uint64_t size) {
    std::vector<std::byte> in(reinterpret_cast<const std::byte *>(bytes), reinterpret_cast<const std::byte *>(bytes) + size);
    ModalModelData parsed;
    if (zpp::bits::failure(zpp::bits::in{in}(parsed))) return -1;
    return parsed == RefModalData(*flat) ? 1 : 0;
}
}
