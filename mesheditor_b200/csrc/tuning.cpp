// Tuning front-end: what the reference computes between a stored modal model and TuneModalObject, after its scene lookups
// (RetuneModalObject, src/audio/AudioSystem.cpp:263-311): the frequency ratio of a fundamental target and a size change, the
// Rayleigh damping law under uniform scaling, the T60 scale; ModalOutGain (:221-224); UniformScaleRatio / MeanScale
// (src/audio/ContactScene.h:97-101, src/TransformMath.h:17-20); the listener attenuation of UpdateListenerGains (:232-243);
// the material a model's modes re-derive at and the pinned fundamental of the edit loop (EffectiveModalMaterial :595-601,
// RescaledModes :612-616), which feed me_rescale_modes;
// and the monitor stage after the mix (MonitorFrames, :1177-1189): pressure to device units under a peak-envelope limiter.
// Host code in float, operation for operation as the reference evaluates it, so that the tuned columns are bit-identical.
#include "common.h"

#include <algorithm>
#include <cmath>
#include <numbers>

namespace me {

void RetuneModes(const float *freqs, const float *t60s, uint32_t n, const MeRetune &rt, float *out_freqs, float *out_t60s) {
    if (!n) return;
    constexpr float ln1000 = 3 * std::numbers::ln10_v<float>; // ModalAudio.h:46
    const float scale = rt.scale;
    // A fundamental target moves every mode by target / first mode; size moves them by 1 / scale.
    const float ratio = (rt.fundamental > 0 && freqs[0] > 0 ? rt.fundamental / freqs[0] : 1.f) / scale;
    const float half_alpha = float(rt.alpha / 2);
    for (uint32_t k = 0; k < n; ++k) {
        out_freqs[k] = freqs[k] * ratio;
        if (t60s[k] <= 0) { // the undamped sentinel stays 0 and mutes the mode
            out_t60s[k] = 0;
            continue;
        }
        // d = (alpha + beta w^2) / 2 with w -> w / scale: the alpha half stays, the rest shrinks by scale^2.
        float rate = ln1000 / t60s[k];
        if (rt.has_alpha) rate = half_alpha + (rate - half_alpha) / (scale * scale);
        out_t60s[k] = rt.t60_scale * ln1000 / std::max(rate, 1e-9f);
    }
}

float ModalOutGain(const MeRetune &rt) { return rt.modal_level * rt.gain * std::pow(rt.scale, -2.f); }

} // namespace me

using namespace me;

extern "C" {

MeStatus me_retune_modes(const float *freqs, const float *t60s, uint32_t n, const MeRetune *rt, float *out_freqs, float *out_t60s) {
    return Guard([&] {
        if (!rt || (n && (!freqs || !t60s || !out_freqs || !out_t60s))) Fail(ME_BAD_ARG, "null argument");
        if (!(rt->scale > 0)) Fail(ME_BAD_ARG, "scale must be positive (me_uniform_scale_ratio clamps it to 0.001..1000)");
        RetuneModes(freqs, t60s, n, *rt, out_freqs, out_t60s);
    });
}

float me_modal_out_gain(const MeRetune *rt) { return rt ? ModalOutGain(*rt) : 0.f; }

float me_uniform_scale_ratio(const float *world_scale, const float *baked_scale) {
    const auto mean = [](const float *s) { return (std::fabs(s[0]) + std::fabs(s[1]) + std::fabs(s[2])) / 3; };
    if (!world_scale || !baked_scale) return 1.f;
    const float baked = mean(baked_scale);
    return baked > 0 ? std::clamp(mean(world_scale) / baked, 0.001f, 1000.f) : 1.f;
}

MeStatus me_monitor_frames(float *frames, uint64_t n, float sample_rate, float *envelope) {
    return Guard([&] {
        if ((n && !frames) || !envelope) Fail(ME_BAD_ARG, "null argument");
        if (!(sample_rate > 0)) Fail(ME_BAD_ARG, "sample_rate must be positive");
        constexpr float full_scale_pressure = 20.f; // Pa: 120 dB SPL (AudioSystem.cpp:1177)
        const float release = std::exp(-1.f / (0.1f * sample_rate)); // 100 ms
        float peak = *envelope;
        for (uint64_t i = 0; i < n; ++i) {
            const float x = frames[i] / full_scale_pressure;
            peak = std::max(std::abs(x), peak * release); // instant attack
            frames[i] = peak > 1.f ? x / peak : x;
        }
        *envelope = peak;
    });
}

MeStatus me_effective_modal_material(const MeMaterial *props, const MeMaterial *solved, double solve_mass, double body_mass, MeMaterial *out) {
    return Guard([&] {
        if (!props || !solved || !out) Fail(ME_BAD_ARG, "null argument");
        *out = *props;
        // The body's one mass: the modes derive at the density that makes the solve's mass meet it, and the stiffness follows so
        // that E / rho - and with it every frequency - stays the named material's.
        if (!(body_mass > 0) || solve_mass <= 0 || solved->density <= 0 || props->density <= 0) return;
        const double density = solved->density * body_mass / solve_mass;
        out->young_modulus *= density / props->density;
        out->density = density;
    });
}

int me_pinned_fundamental(const float *freqs, uint32_t n, float original_fundamental, float *fundamental) {
    const bool pinned = n && freqs && original_fundamental > 0 && freqs[0] != original_fundamental;
    if (pinned && fundamental) *fundamental = freqs[0];
    return pinned;
}

float me_listener_gain(float distance) {
    constexpr float listener_distance = 1.f; // ModalAudio.h:43
    return listener_distance / std::max(distance, listener_distance);
}

} // extern "C"
