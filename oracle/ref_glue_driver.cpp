// TEST INFRASTRUCTURE ONLY (oracle). Never linked into the product library.
//
// extern "C" wrapper around the UNMODIFIED sample-surface helpers of the reference's modal generation job
// (UniqueSampleTriangles, SampleSurfaceTriangles, CompactExcitationVertices, RelabelSampleTriangles:
// /root/reference/src/audio/AudioSystem.cpp:673-769). AudioSystem.cpp as a whole needs the editor's dependencies, so
// oracle/Makefile cuts exactly those definitions out of it at build time (glue_slice.inc, written to oracle/_ref/ and removed
// after compiling) and this file includes them: the code that runs is the reference's, and none of it is stored in the repo.
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <span>
#include <vector>

#include "glue_slice.inc"

// RetuneModalObject's arithmetic (AudioSystem.cpp:271,299-308): the statements themselves are cut out of the reference like
// the functions above; this wrapper supplies the names they read (what the function looked up in the scene before them).
#include <numbers>
#include <optional>
namespace {
constexpr float Ln1000 = 3 * std::numbers::ln10_v<float>; // src/audio/ModalAudio.h:46
struct {
    std::vector<float> Freqs, T60s;
} modes;
} // namespace
extern "C" void ref_retune(const float *in_freqs, const float *in_t60s, uint32_t n, float scale, float fundamental, float t60_scale, int has_alpha, double alpha_value, float *out_freqs,
                           float *out_t60s) {
    modes.Freqs.assign(in_freqs, in_freqs + n), modes.T60s.assign(in_t60s, in_t60s + n);
    const size_t mode_count = n;
    std::optional<double> alpha;
    if (has_alpha) alpha = alpha_value;
#include "retune_ratio.inc"
#include "retune_loop.inc"
    std::memcpy(out_freqs, freqs.data(), n * sizeof(float)), std::memcpy(out_t60s, t60s.data(), n * sizeof(float));
}

namespace {
std::vector<uint32_t> g_out;
uint32_t Keep(std::vector<uint32_t> v) {
    g_out = std::move(v);
    return uint32_t(g_out.size());
}
} // namespace

extern "C" {
uint32_t ref_sample_surface_triangles(const uint32_t *triangle_indices, uint32_t n, uint32_t vertex_count, const uint32_t *excitation_vertices, uint32_t n_ex) {
    return Keep(SampleSurfaceTriangles({triangle_indices, n}, vertex_count, {excitation_vertices, n_ex}));
}
uint32_t ref_compact_excitation_vertices(const uint32_t *vertices, uint32_t n, const uint32_t *sample_point_of, uint32_t n_sp) {
    return Keep(CompactExcitationVertices({vertices, n}, {sample_point_of, n_sp}));
}
uint32_t ref_relabel_sample_triangles(const uint32_t *triangles, uint32_t n, const uint32_t *sample_point_of, uint32_t n_sp) {
    return Keep(RelabelSampleTriangles({triangles, n}, {sample_point_of, n_sp}));
}
void ref_glue_copy(uint32_t *out) { std::memcpy(out, g_out.data(), g_out.size() * sizeof(uint32_t)); }
}

// MonitorFrames (AudioSystem.cpp:1177-1189): its constant and loop cut out of the reference; the release coefficient is the
// function's first statement with the device sample rate passed in (the reference reads it from the registry).
#include <cmath>
#include "monitor_const.inc"
namespace {
struct MonitorLimiter {
    float Envelope{0};
}; // src/audio/AudioTypes.h:20-22
} // namespace
extern "C" float ref_monitor_frames(float *data, uint64_t n, float sample_rate, float envelope_in) {
    MonitorLimiter limiter{envelope_in};
    std::span<float> frames{data, n};
    const float release = std::exp(-1.f / (0.1f * float(sample_rate)));
#include "monitor_loop.inc"
    return limiter.Envelope;
}

// EffectiveModalMaterial (AudioSystem.cpp:595-601): its arithmetic cut out of the reference; the guard is the caller's
// (oracle/generation.py: an authoritative dynamic body with positive masses and densities).
namespace {
struct Properties {
    double Density, YoungModulus;
};
struct Motion {
    std::optional<float> Mass;
};
constexpr float DefaultMass{1}; // src/physics/PhysicsTypes.h:132
} // namespace
extern "C" void ref_effective_material(double *density, double *young, double solved_density, double solve_mass, int has_mass, float body_mass) {
    Properties props{*density, *young};
    const struct {
        Properties SolvedMaterial;
    } summary{{solved_density, 0.0}};
    Motion body;
    if (has_mass) body.Mass = body_mass;
    const Motion *motion = &body;
#include "effective_material.inc"
    *density = props.Density, *young = props.YoungModulus;
}

// EstimateFundamentalFrequency (AudioSystem.cpp:522-550), whole. FFTData (src/audio/FFTData.h) wraps an FFTW plan, which this
// image does not have; the function reads only the spectrum and the real length, which this stand-in carries (the spectrum is
// the caller's: numpy's transform of the windowed segment).
namespace {
struct FFTData {
    const float (*Complex)[2];
    size_t NumReal;
};
using std::ranges::nth_element; // AudioSystem.cpp:65
} // namespace
#if defined(__GLIBCXX__) && _GLIBCXX_RELEASE < 14
namespace std {
inline float log10f(float x) { return ::log10f(x); } // C++23's std::log10f, which this libstdc++ does not declare yet
} // namespace std
#endif
#include "fundamental.inc"
extern "C" int ref_estimate_fundamental(const float *complex_re_im, uint64_t n_real, uint32_t sample_rate, float *hz) {
    const FFTData fft{reinterpret_cast<const float (*)[2]>(complex_re_im), size_t(n_real)};
    const auto found = EstimateFundamentalFrequency(fft, sample_rate);
    if (found) *hz = *found;
    return found.has_value();
}

// TiltAlongNormal and SphereEquivalentCurvature (AudioSystem.cpp:359-380), whole, over the reference's glm vector types.
#include "numeric/vec2.h"
#include "numeric/vec3.h"
#include <glm/geometric.hpp>
#include "strike_direction.inc"
extern "C" void ref_tilt_along_normal(const float *n, const float *joy, float *out) {
    const vec3 d = TiltAlongNormal({n[0], n[1], n[2]}, {joy[0], joy[1]});
    out[0] = d.x, out[1] = d.y, out[2] = d.z;
}
extern "C" double ref_sphere_equivalent_curvature(double density, double inv_mass) { return SphereEquivalentCurvature(density, inv_mass); }

// DesiredSolveVertices (AudioSystem.cpp:667-671) past its copied-vertices branch.
#include <ranges>
namespace {
using std::ranges::iota_view; // AudioSystem.cpp:65-66
using std::views::transform;
#if defined(__GLIBCXX__) && _GLIBCXX_RELEASE < 14
// C++23's std::ranges::to, which this libstdc++ does not have yet: the one use below collects a view into a vector.
template<typename Container>
struct Collect {};
template<typename Container>
Collect<Container> to() { return {}; }
template<std::ranges::input_range Range, typename Container>
Container operator|(Range &&range, Collect<Container>) {
    Container out;
    for (auto &&v : range) out.push_back(v);
    return out;
}
#else
using std::ranges::to;
#endif
std::vector<uint32_t> EvenVertices(uint32_t requested, uint32_t num_vertices) {
    const struct {
        uint32_t NumVertices;
    } settings{requested};
#include "solve_vertices.inc"
}
} // namespace
extern "C" uint32_t ref_desired_solve_vertices(uint32_t requested, uint32_t num_vertices) { return Keep(EvenVertices(requested, num_vertices)); }
