// Strike front-end: from a struck object's contact dynamics to the ModalEvent the bank consumes.
// Host code, like the reference's (scalar FP64 per strike; nothing here is a hot loop). Reference: src/audio/ContactModel.cpp
// (Hertz contact constants :40-66, ReducedContactMass :27-38, EstimateContactTime :76-114), RecoilClickFilter
// (src/audio/ModalAudio.h:50-99) and the arithmetic of TriggerModalStrike (src/audio/AudioSystem.cpp:400-465) with the
// scene lookups replaced by plain arguments. Float and double roles follow the reference (glm::vec3 / glm::mat3 are
// float, everything scalar is double), so results agree to rounding.
#include "common.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numbers>

namespace me {
namespace {

constexpr double kMinContactTime = 2e-5, kMaxContactTime = 5e-2; // ContactModel.h:80
constexpr float kAirDensity = 1.204f, kSpeedOfSound = 343.f, kListenerDistance = 1.f; // ModalAudio.h:41-43

struct Float3 {
    float x, y, z;
};
Float3 Normalize(const float *v) {
    const float inv = 1.f / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    return {v[0] * inv, v[1] * inv, v[2] * inv};
}
Float3 Cross(Float3 a, Float3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }

double InvEffectiveModulus(const MeMaterial &a, const MeMaterial &b) {
    return (1 - a.poisson_ratio * a.poisson_ratio) / a.young_modulus + (1 - b.poisson_ratio * b.poisson_ratio) / b.young_modulus;
}
double CombinedCurvature(double a, double b) { return std::max(a + b, 1e-6); }
double ContactStiffness(double inv_modulus, double curvature) { return 4.0 / 3.0 / inv_modulus / std::sqrt(curvature); }
double SaturationPenetration(double curvature, double area) { return area > 0 ? area * curvature / std::numbers::pi : std::numeric_limits<double>::infinity(); }
double PunchStiffness(double inv_modulus, double area) {
    if (area <= 0) return std::numeric_limits<double>::infinity();
    return 2 * std::sqrt(area / std::numbers::pi) / inv_modulus;
}

// Work against the contact pressed to `depth`: Hertz's k x^(3/2) below saturation, the filled patch's constant stiffness above.
double ContactWork(double depth, double hertz, double saturation, double punch) {
    if (depth <= 0) return 0;
    const auto below = [hertz](double x) { return 0.4 * hertz * x * x * std::sqrt(x); };
    if (depth <= saturation) return below(depth);
    const double over = depth - saturation;
    const double force_at_saturation = hertz * saturation * std::sqrt(saturation);
    return below(saturation) + force_at_saturation * over + 0.5 * punch * over * over;
}

double ReducedMass(const MeContactDynamics &d, uint32_t i, const float *direction, const MeImpactor &impactor) {
    if (i >= d.arm_count || d.mass <= 0) return 0;
    const Float3 n = Normalize(direction);
    const Float3 arm{d.contact_arm_xyz[3 * i], d.contact_arm_xyz[3 * i + 1], d.contact_arm_xyz[3 * i + 2]};
    const Float3 c = Cross(arm, n);
    const float *m = d.inverse_inertia; // column-major
    const Float3 mc{m[0] * c.x + m[3] * c.y + m[6] * c.z, m[1] * c.x + m[4] * c.y + m[7] * c.z, m[2] * c.x + m[5] * c.y + m[8] * c.z};
    const float lever = c.x * mc.x + c.y * mc.y + c.z * mc.z;
    return 1.0 / (1.0 / d.mass + lever + impactor.inv_mass);
}

struct Poles {
    double A0;
    float A1, A2;
};
Poles RecoilDenominator(double wc, double kk, double beta) {
    const double a0 = kk * kk + beta * wc * kk + beta * wc * wc;
    return {a0, float((2 * beta * wc * wc - 2 * kk * kk) / a0), float((kk * kk - beta * wc * kk + beta * wc * wc) / a0)};
}

} // namespace
} // namespace me

using namespace me;

extern "C" {

double me_striker_mass(const MeStriker *s) {
    if (!s) return 0;
    const double r = s->tip_radius, l = s->length;
    return s->material.density * std::numbers::pi * (r * r * l + 4.0 / 3.0 * r * r * r);
}

MeStatus me_striker_impactor(const MeStriker *s, MeImpactor *out) {
    return Guard([&] {
        if (!s || !out) Fail(ME_BAD_ARG, "null striker / impactor");
        *out = {.material = s->material, .curvature = 1.0 / s->tip_radius, .inv_mass = 1.0 / me_striker_mass(s)};
    });
}

MeStatus me_inverse_inertia_tensor(const MeMassProperties *mp, float out[9]) {
    return Guard([&] {
        if (!mp || !out) Fail(ME_BAD_ARG, "null mass properties / output");
        const float w = mp->inertia_orientation[0], x = mp->inertia_orientation[1], y = mp->inertia_orientation[2], z = mp->inertia_orientation[3];
        const float xx = x * x, yy = y * y, zz = z * z, xz = x * z, xy = x * y, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
        // Rotation of the principal axes, column-major r[column][row].
        const float r[3][3] = {{1.f - 2.f * (yy + zz), 2.f * (xy + wz), 2.f * (xz - wy)}, {2.f * (xy - wz), 1.f - 2.f * (xx + zz), 2.f * (yz + wx)}, {2.f * (xz + wy), 2.f * (yz - wx), 1.f - 2.f * (xx + yy)}};
        float inv[3];
        for (int i = 0; i < 3; ++i) inv[i] = mp->inertia_diagonal[i] > 0 ? 1.f / mp->inertia_diagonal[i] : 0.f;
        // (r * diag(inv)) * r^T, each product accumulated over the inner index in order.
        float scaled[3][3];
        for (int c = 0; c < 3; ++c)
            for (int row = 0; row < 3; ++row) scaled[c][row] = r[0][row] * (c == 0 ? inv[0] : 0.f) + r[1][row] * (c == 1 ? inv[1] : 0.f) + r[2][row] * (c == 2 ? inv[2] : 0.f);
        for (int c = 0; c < 3; ++c)
            for (int row = 0; row < 3; ++row) out[3 * c + row] = scaled[0][row] * r[0][c] + scaled[1][row] * r[1][c] + scaled[2][row] * r[2][c];
    });
}

double me_reduced_contact_mass(const MeContactDynamics *d, uint32_t i, const float direction[3], const MeImpactor *impactor) {
    if (!d || !direction || !impactor || (d->arm_count && !d->contact_arm_xyz)) return 0;
    return ReducedMass(*d, i, direction, *impactor);
}

double me_estimate_contact_time(const MeContactDynamics *d, uint32_t i, const float direction[3], double contact_speed, const MeMaterial *object_material, double object_curvature, double nominal_area,
                                const MeImpactor *impactor, double scale_ratio, double combined_roughness) {
    if (!d || !direction || !object_material || !impactor || (d->arm_count && !d->contact_arm_xyz)) return kMinContactTime;
    if (i >= d->arm_count || d->mass <= 0) return kMinContactTime;
    const double effective_mass = ReducedMass(*d, i, direction, *impactor);
    const double inv_modulus = InvEffectiveModulus(*object_material, impactor->material);
    if (effective_mass <= 0 || inv_modulus <= 0) return kMinContactTime;

    const double curvature = CombinedCurvature(object_curvature, impactor->curvature);
    const double speed = std::max(std::abs(contact_speed), 1e-6);
    const double hertz = ContactStiffness(inv_modulus, curvature);
    const double saturation = SaturationPenetration(curvature, nominal_area);
    const double punch = PunchStiffness(inv_modulus, nominal_area);
    const double energy = 0.5 * effective_mass * speed * speed;

    // Deepest penetration: where the approach energy has all gone into the contact.
    const double work_at_saturation = std::isfinite(saturation) ? ContactWork(saturation, hertz, saturation, punch) : std::numeric_limits<double>::infinity();
    double deepest;
    if (energy <= work_at_saturation) {
        deepest = std::pow(energy / (0.4 * hertz), 0.4);
    } else {
        const double force_at_saturation = hertz * saturation * std::sqrt(saturation);
        deepest = saturation + (std::sqrt(force_at_saturation * force_at_saturation + 2 * punch * (energy - work_at_saturation)) - force_at_saturation) / punch;
    }
    // Twice the approach time, by the midpoint rule in s with x = deepest * (1 - s^2) (removes the turning-point singularity).
    constexpr int steps = 64;
    double sum = 0;
    for (int n = 0; n < steps; ++n) {
        const double s = (double(n) + 0.5) / steps;
        const double left = 1 - ContactWork(deepest * (1 - s * s), hertz, saturation, punch) / energy;
        if (left > 0) sum += 2 * s / std::sqrt(left);
    }
    const double bulk_time = 2 * deepest / speed * sum / steps * scale_ratio;
    // Asperity cushion of a rough interface, in series with the bulk: contact times add in quadrature.
    const double u0 = 0.4 * combined_roughness;
    const double bed_time = std::numbers::sqrt2 * std::numbers::pi * u0 / speed;
    return std::clamp(std::sqrt(bulk_time * bulk_time + bed_time * bed_time), kMinContactTime, kMaxContactTime);
}

double me_contact_constant(MeContactConstant which, const MeMaterial *a, const MeMaterial *b, double x, double y, double z) {
    switch (which) {
        case ME_CONTACT_INV_EFFECTIVE_MODULUS: return a && b ? InvEffectiveModulus(*a, *b) : 0;
        case ME_CONTACT_COMBINED_CURVATURE: return CombinedCurvature(x, y);
        case ME_CONTACT_STIFFNESS: return ContactStiffness(x, y);
        case ME_CONTACT_PATCH_RADIUS: return std::cbrt(0.75 * std::max(x, 0.0) * y / z);
        case ME_CONTACT_STATIC_PENETRATION: return y > 0 ? std::pow(std::max(x, 0.0) / y, 2.0 / 3.0) : 0.0;
        case ME_CONTACT_SATURATION_PENETRATION: return SaturationPenetration(x, y);
        case ME_CONTACT_PUNCH_STIFFNESS: return PunchStiffness(x, y);
    }
    return 0;
}

void me_recoil_click_filter(double radius, double volume, double mass, double sample_rate, float b0_a1_a2[3]) {
    if (!b0_a1_a2) return;
    b0_a1_a2[0] = b0_a1_a2[1] = b0_a1_a2[2] = 0.f;
    if (radius <= 0 || mass <= 0) return;
    const double wc = kSpeedOfSound / radius;
    const double kk = 2 * sample_rate;
    const Poles poles = RecoilDenominator(wc, kk, 2 + kAirDensity * volume / mass);
    const double g = kAirDensity * kSpeedOfSound * radius / (kListenerDistance * mass);
    b0_a1_a2[0] = float(g * kk / poles.A0), b0_a1_a2[1] = poles.A1, b0_a1_a2[2] = poles.A2;
}

MeStatus me_make_strike_event(const MeStrike *s, MeModalEvent *out) {
    return Guard([&] {
        if (!s || !out) Fail(ME_BAD_ARG, "null strike / event");
        if (!(s->sample_rate > 0)) Fail(ME_BAD_ARG, "sample rate must be positive");
        float dir[3] = {s->direction[0], s->direction[1], s->direction[2]};
        if (s->is_collision) { // a collision's direction arrives unnormalised (AudioSystem.cpp:408)
            const Float3 n = Normalize(dir);
            dir[0] = n.x, dir[1] = n.y, dir[2] = n.z;
        }
        // A short default contact with no click applies when the material or contact dynamics are missing (:412-415).
        double tau = 1e-4;
        float click_amp = 0.f, click[3] = {0.f, 0.f, 0.f};
        if (s->dynamics && s->elastic) {
            const MeContactDynamics &cd = *s->dynamics;
            tau = me_estimate_contact_time(&cd, s->is_collision ? s->resultant_index : s->excitable_index, dir, s->contact_speed, s->elastic, s->curvature, s->is_collision ? s->nominal_area : 0.0, &s->impactor,
                                           s->scale_ratio, s->roughness);
            const double volume = s->displaced_volume;
            const double radius = volume > 0 ? std::cbrt(3.0 * volume / (4.0 * std::numbers::pi)) : double(s->radiant_radius * s->scale_ratio);
            me_recoil_click_filter(radius, volume, cd.mass, s->sample_rate, click);
            // A collision's force is the true contact impulse; a mallet's is nominal, from the reduced mass and approach speed.
            const double impulse = s->is_collision ? double(s->force) : me_reduced_contact_mass(&cd, s->excitable_index, dir, &s->impactor) * std::abs(double(s->contact_speed));
            click_amp = float(impulse * s->sample_rate);
        }
        const float step = float(1.0 / (tau * s->sample_rate));
        *out = MeModalEvent{};
        out->kind = 0;
        out->object = s->object;
        out->ex_pos = s->excitable_index;
        out->jx = dir[0] * s->force, out->jy = dir[1] * s->force, out->jz = dir[2] * s->force;
        out->pulse_step = step;
        out->pulse_gamma = 2 * step;
        out->accel_amp = click_amp;
        out->click_b0 = click[0], out->click_a1 = click[1], out->click_a2 = click[2];
    });
}

// UpdateContactDynamics (src/audio/ContactDynamics.cpp:19-46) past its registry lookups: the ContactDynamics a strike's contact time
// is estimated with, from the mass properties the lookups resolved (the solve's, or an authoritative rigid body's) and the model.
MeStatus me_contact_dynamics(const MeMassProperties *resolved, double mass_scale, const float *positions_xyz, uint32_t n_positions, const float baked_scale[3], double *mass,
                             float inverse_inertia[9], float *arms_xyz) {
    return Guard([&] {
        if (!resolved || !baked_scale || !mass || !inverse_inertia || (n_positions && (!positions_xyz || !arms_xyz))) Fail(ME_BAD_ARG, "null argument");
        const float size = std::max((std::fabs(baked_scale[0]) + std::fabs(baked_scale[1]) + std::fabs(baked_scale[2])) / 3, 1e-6f);
        *mass = resolved->mass * mass_scale;
        if (const MeStatus status = me_inverse_inertia_tensor(resolved, inverse_inertia); status != ME_OK) throw Failure{status};
        const float lighten = float(1 / mass_scale); // a denser material is as much heavier as it is harder to turn
        for (int k = 0; k < 9; ++k) inverse_inertia[k] *= lighten;
        // Arms from the centre of mass to each sample point, node-local lengths brought to metres at the baked size.
        for (uint32_t p = 0; p < n_positions; ++p)
            for (int k = 0; k < 3; ++k) arms_xyz[3 * p + k] = (positions_xyz[3 * p + k] - resolved->center_of_mass[k]) * size;
    });
}

// Strike direction and the colliding body's curvature, as TriggerModalStrike's callers prepare them (AudioSystem.cpp:359-379).
void me_tilt_along_normal(const float normal[3], const float joystick[2], float out[3]) {
    const float nx = normal[0], ny = normal[1], nz = normal[2], jx = joystick[0], jy = joystick[1];
    const float radius = std::sqrt(jx * jx + jy * jy);
    if (radius < 1e-6f) { // the centre of the pad: straight along the normal
        out[0] = nx, out[1] = ny, out[2] = nz;
        return;
    }
    // Branchless orthonormal tangent frame of the normal (Duff et al. 2017), the one the reference builds.
    const float sign = nz >= 0 ? 1.f : -1.f;
    const float a = -1.f / (sign + nz);
    const float b = nx * ny * a;
    const float t[3] = {1.f + sign * nx * nx * a, sign * b, -sign * nx};
    const float bt[3] = {b, sign + ny * ny * a, -ny};
    const float theta = std::min(radius, 1.f) * 1.57079633f; // the rim lies in the tangent plane
    const float c = std::cos(theta), s = std::sin(theta);
    for (int k = 0; k < 3; ++k) out[k] = c * normal[k] + s * (jx * t[k] + jy * bt[k]) / radius;
}

double me_sphere_equivalent_curvature(double density, double inv_mass) { return std::cbrt(4.0 * std::numbers::pi / 3.0 * density * inv_mass); }

} // extern "C"
