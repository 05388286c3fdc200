/* TEST INFRASTRUCTURE ONLY (oracle). Never linked into, imported by, or called from the product path.
 *
 * Plain-C, single-thread restatement of the reference modal resonator bank:
 *   AddModalObject      /root/reference/src/audio/ModalAudio.cpp:291-338
 *   TuneModalObject     ModalAudio.cpp:340-393
 *   EnqueueModalEvent   ModalAudio.cpp:417-425   (SPSC ring of 256, drop + count when full)
 *   DrainEvents / ActivateImpact / SilenceObject   ModalAudio.cpp:28-82
 *   RenderModal         ModalAudio.cpp:486-590   (force/click stage :506-538, retire :557-561)
 *   RenderObjectFast    ModalAudio.cpp:86-147    (8-lane chunks, audibility culling)
 *   InstallModalBank    ModalAudio.cpp:277-289   (FlushEvents: the next render drops queued events)
 * Constants: ModalAudio.h:41-46,169,263,275; SilentEnergy ModalAudio.cpp:20.
 *
 * Pinned: tests/test_oracle_resonator.py checks this file sample-for-sample (bit-exact) against the
 * unmodified reference built by oracle/Makefile into oracle/_ref/libme_ref_audio.so, and replays the
 * reference's own ModalRenderTest properties (tests/ModalRenderTest.cpp:21-68).
 * Build with -ffp-contract=off: the reference is x86-64 -O2 without FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OR_LANES 8u
#define OR_EVENT_CAPACITY 256u
static const float kSilentEnergy = 1e-12f;
static const float kAirDensity = 1.204f, kSpeedOfSound = 343.f, kListenerDistance = 1.f;
static const float kPi = 3.14159265358979323846f;
#define OR_LN1000 (3.f * 2.302585092994046f)

typedef struct {
    uint32_t Kind, Object, ExPos;
    float Jx, Jy, Jz, PulseStep, PulseGamma, AccelAmp, ClickB0, ClickA1, ClickA2;
} OrEvent;

typedef struct {
    uint32_t Object, ExPos, SamplesLeft;
    float Jx, Jy, Jz, PhaseRe, PhaseIm, RotRe, RotIm, Gamma, AccelAmp, ClickB0, ClickA1, ClickA2, ClickZ1, ClickZ2;
} OrImpact;

typedef struct {
    float *p;
    size_t n, cap;
} FVec;

static void fv_resize(FVec *v, size_t n, float fill) {
    if (n > v->cap) {
        size_t cap = v->cap ? v->cap : 64;
        while (cap < n) cap *= 2;
        v->p = (float *)realloc(v->p, cap * sizeof(float));
        v->cap = cap;
    }
    for (size_t i = v->n; i < n; ++i) v->p[i] = fill;
    v->n = n;
}

typedef struct OrBank {
    float SampleRate;
    /* per mode */
    FVec CoeffRe, CoeffIm, StateRe, StateIm, RadiationGain, RadiationArea, DeflectionGain, OutPhaseIm, OutPhaseRe, QuadCompliance, QuadDriveScale;
    FVec ShapeX, ShapeY, ShapeZ;
    /* per object */
    uint32_t NObjects, ObjCap;
    uint32_t *ModeOffset, *ModeCount, *ShapeOffset, *TunedModeCount, *LiveModeCount;
    uint8_t *Ringing;
    float *OutGain, *ListenerGain, *RadiantRadius, *DeflectionScale;
    /* impacts */
    OrImpact *Impacts;
    uint32_t NImpacts, ImpactCap;
    /* audio-side state (ModalAudio) */
    float ClickGain;
    uint32_t MaxImpacts;
    OrEvent Events[OR_EVENT_CAPACITY];
    uint32_t EventWrite, EventRead;
    int FlushEvents;
    uint64_t EventsDropped;
    float *ForceScratch;
    size_t ForceCap;
    float *Gains;
    size_t GainsCap;
    uint32_t *ObjImpacts;
    size_t ObjImpactsCap;
    int Cull; /* 1 = reference behaviour; 0 = render every tuned mode (diagnostic only) */
    /* FP64 arbiter (or_bank_render_exact): the same bank, the same float parameters, states and sums in double. */
    double *ExactRe, *ExactIm;
    size_t ExactN;
} OrBank;

OrBank *or_bank_create(float sample_rate) {
    OrBank *b = (OrBank *)calloc(1, sizeof(OrBank));
    b->SampleRate = sample_rate;
    b->ClickGain = 1.f;
    b->MaxImpacts = 1024;
    b->Cull = 1;
    return b;
}

void or_bank_free(OrBank *b) {
    if (!b) return;
    FVec *vs[] = {&b->CoeffRe, &b->CoeffIm, &b->StateRe, &b->StateIm, &b->RadiationGain, &b->RadiationArea, &b->DeflectionGain, &b->OutPhaseIm, &b->OutPhaseRe, &b->QuadCompliance, &b->QuadDriveScale, &b->ShapeX, &b->ShapeY, &b->ShapeZ};
    for (size_t i = 0; i < sizeof vs / sizeof *vs; ++i) free(vs[i]->p);
    free(b->ModeOffset), free(b->ModeCount), free(b->ShapeOffset), free(b->TunedModeCount), free(b->LiveModeCount);
    free(b->Ringing), free(b->OutGain), free(b->ListenerGain), free(b->RadiantRadius), free(b->DeflectionScale);
    free(b->Impacts), free(b->ForceScratch), free(b->Gains), free(b->ObjImpacts);
    free(b->ExactRe), free(b->ExactIm);
    free(b);
}

void or_bank_set_cull(OrBank *b, int cull) { b->Cull = cull; }
void or_bank_set_click_gain(OrBank *b, float g) { b->ClickGain = g; }
void or_bank_set_max_impacts(OrBank *b, uint32_t n) { b->MaxImpacts = n; }

static void grow_objects(OrBank *b) {
    if (b->NObjects < b->ObjCap) return;
    const uint32_t cap = b->ObjCap ? b->ObjCap * 2 : 16;
#define GROW(field, type) b->field = (type *)realloc(b->field, cap * sizeof(type))
    GROW(ModeOffset, uint32_t), GROW(ModeCount, uint32_t), GROW(ShapeOffset, uint32_t), GROW(TunedModeCount, uint32_t), GROW(LiveModeCount, uint32_t);
    GROW(Ringing, uint8_t), GROW(OutGain, float), GROW(ListenerGain, float), GROW(RadiantRadius, float), GROW(DeflectionScale, float);
#undef GROW
    b->ObjCap = cap;
}

/* shapes: [point][mode][3], positions: [point][3], indices: triangles over the points. ModalAudio.cpp:291-338 */
uint32_t or_bank_add_object(OrBank *b, uint32_t count, uint32_t n_points, const float *shapes, const float *positions, const uint32_t *indices, uint32_t n_indices) {
    grow_objects(b);
    const uint32_t slot = b->NObjects++;
    b->ModeOffset[slot] = (uint32_t)b->CoeffRe.n;
    b->ModeCount[slot] = b->TunedModeCount[slot] = b->LiveModeCount[slot] = count;
    b->ShapeOffset[slot] = (uint32_t)b->ShapeX.n;
    b->Ringing[slot] = 0;
    b->OutGain[slot] = 0.f;
    b->ListenerGain[slot] = 1.f;
    b->DeflectionScale[slot] = 1.f;
    const size_t k0 = b->CoeffRe.n;
    FVec *zero_cols[] = {&b->CoeffRe, &b->CoeffIm, &b->StateRe, &b->StateIm, &b->RadiationGain, &b->DeflectionGain, &b->QuadCompliance, &b->QuadDriveScale};
    for (size_t i = 0; i < sizeof zero_cols / sizeof *zero_cols; ++i) fv_resize(zero_cols[i], k0 + count, 0.f);
    fv_resize(&b->OutPhaseIm, k0 + count, 1.f);
    fv_resize(&b->OutPhaseRe, k0 + count, 0.f);
    const size_t s0 = b->ShapeX.n, ns = (size_t)n_points * count;
    fv_resize(&b->ShapeX, s0 + ns, 0.f), fv_resize(&b->ShapeY, s0 + ns, 0.f), fv_resize(&b->ShapeZ, s0 + ns, 0.f);
    for (size_t i = 0; i < ns; ++i) {
        b->ShapeX.p[s0 + i] = shapes[3 * i], b->ShapeY.p[s0 + i] = shapes[3 * i + 1], b->ShapeZ.p[s0 + i] = shapes[3 * i + 2];
    }
    fv_resize(&b->RadiationArea, k0 + count, 0.f);
    float total_area = 0.f;
    for (size_t t = 0; t + 2 < n_indices; t += 3) {
        const uint32_t i = indices[t], j = indices[t + 1], l = indices[t + 2];
        const float *pi = positions + 3 * i, *pj = positions + 3 * j, *pl = positions + 3 * l;
        const float ux = pj[0] - pi[0], uy = pj[1] - pi[1], uz = pj[2] - pi[2];
        const float vx = pl[0] - pi[0], vy = pl[1] - pi[1], vz = pl[2] - pi[2];
        /* glm::cross */
        const float cx = uy * vz - vy * uz, cy = uz * vx - vz * ux, cz = ux * vy - vx * uy;
        const float doubled = sqrtf(cx * cx + cy * cy + cz * cz); /* glm::length = sqrt(dot) */
        if (doubled <= 0.f) continue;
        const float nx = cx / doubled, ny = cy / doubled, nz = cz / doubled;
        const float area = doubled / 2;
        total_area += area;
        for (uint32_t k = 0; k < count; ++k) {
            const float *si = shapes + ((size_t)i * count + k) * 3, *sj = shapes + ((size_t)j * count + k) * 3, *sl = shapes + ((size_t)l * count + k) * 3;
            const float mx = (si[0] + sj[0] + sl[0]) / 3.f, my = (si[1] + sj[1] + sl[1]) / 3.f, mz = (si[2] + sj[2] + sl[2]) / 3.f;
            const float normal = mx * nx + my * ny + mz * nz;
            b->RadiationArea.p[k0 + k] += area * normal * normal;
        }
    }
    b->RadiantRadius[slot] = sqrtf(total_area / (4 * kPi));
    return slot;
}

/* ModalAudio.cpp:340-393 */
void or_bank_tune_object(OrBank *b, uint32_t object, const float *freqs, const float *t60s, uint32_t n, float radius_scale) {
    const uint32_t k0 = b->ModeOffset[object];
    const uint32_t count = b->ModeCount[object] < n ? b->ModeCount[object] : n;
    const float sr = b->SampleRate;
    const float radius = b->RadiantRadius[object] * radius_scale;
    b->DeflectionScale[object] = 1.f / (radius_scale * radius_scale * radius_scale);
    for (uint32_t k = 0; k < count; ++k) {
        const float freq = freqs[k], t60 = t60s[k];
        const size_t m = (size_t)k0 + k;
        if (!isfinite(freq) || !isfinite(t60) || freq <= 0.f || freq >= sr / 2 - 1 || t60 <= 0.f) {
            b->CoeffRe.p[m] = b->CoeffIm.p[m] = b->RadiationGain.p[m] = b->DeflectionGain.p[m] = 0.f;
            b->OutPhaseIm.p[m] = 1.f, b->OutPhaseRe.p[m] = 0.f;
            b->QuadCompliance.p[m] = b->QuadDriveScale.p[m] = 0.f;
            continue;
        }
        const float omega = 2 * kPi * freq / sr;
        const float omega_si = 2 * kPi * freq;
        const float ka = omega_si * radius / kSpeedOfSound;
        const float sigma = ka * ka / (1 + ka * ka);
        const float area = b->RadiationArea.p[m] / radius_scale;
        const float radiation_rate = kAirDensity * kSpeedOfSound * sigma * area * 0.5f;
        const float decay = expf(-(OR_LN1000 / t60 + radiation_rate) / sr);
        b->CoeffRe.p[m] = decay * cosf(omega);
        b->CoeffIm.p[m] = decay * sinf(omega);
        const float gain = kAirDensity * kSpeedOfSound * sqrtf(sigma * b->RadiationArea.p[m] / (4 * kPi)) / kListenerDistance;
        b->RadiationGain.p[m] = gain;
        const float spread = sigma * kPi * (2.f * fmodf(0.6180339887f * (float)(k + 1), 1.0f) - 1.f);
        b->OutPhaseIm.p[m] = cosf(spread);
        b->OutPhaseRe.p[m] = sinf(spread);
        b->DeflectionGain.p[m] = gain > 0.f ? 1.f / (gain * omega_si) : 0.f;
        const float dt = 1.f / sr;
        const float central = dt * (1 + decay * decay + 2 * decay * cosf(omega)) / 4;
        b->QuadCompliance.p[m] = central;
        b->QuadDriveScale.p[m] = central * omega_si / (decay * sinf(omega));
    }
    uint32_t live = b->ModeCount[object];
    while (live > 0 && b->CoeffRe.p[k0 + live - 1] == 0.f && b->CoeffIm.p[k0 + live - 1] == 0.f) --live;
    b->TunedModeCount[object] = live;
    b->LiveModeCount[object] = live;
}

void or_bank_set_gain(OrBank *b, uint32_t slot, float out_gain, float listener_gain) {
    b->OutGain[slot] = out_gain;
    b->ListenerGain[slot] = listener_gain;
}

/* InstallModalBank: the adopting callback drops whatever was queued against the old layout. ModalAudio.cpp:277-289,496-498 */
void or_bank_install(OrBank *b) { b->FlushEvents = 1; }

/* ModalAudio.cpp:417-425. Returns 1 when queued, 0 when dropped. */
int or_bank_enqueue(OrBank *b, const OrEvent *e) {
    if (b->EventWrite - b->EventRead >= OR_EVENT_CAPACITY) {
        ++b->EventsDropped;
        return 0;
    }
    b->Events[b->EventWrite % OR_EVENT_CAPACITY] = *e;
    ++b->EventWrite;
    return 1;
}

static void remove_impact(OrBank *b, uint32_t i) { b->Impacts[i] = b->Impacts[--b->NImpacts]; }

static void silence_object(OrBank *b, uint32_t o) {
    const uint32_t k0 = b->ModeOffset[o], count = b->ModeCount[o];
    memset(b->StateRe.p + k0, 0, count * sizeof(float));
    memset(b->StateIm.p + k0, 0, count * sizeof(float));
    b->Ringing[o] = 0;
    b->LiveModeCount[o] = b->TunedModeCount[o];
    for (uint32_t i = b->NImpacts; i-- > 0;) {
        if (b->Impacts[i].Object == o) remove_impact(b, i);
    }
}

static void activate_impact(OrBank *b, const OrEvent *e) {
    if (b->NImpacts >= b->MaxImpacts) return;
    if (b->NImpacts == b->ImpactCap) {
        b->ImpactCap = b->ImpactCap ? b->ImpactCap * 2 : 64;
        b->Impacts = (OrImpact *)realloc(b->Impacts, b->ImpactCap * sizeof(OrImpact));
    }
    const float theta = 2 * kPi * e->PulseStep;
    OrImpact im = {e->Object, e->ExPos, (uint32_t)ceilf(1.f / e->PulseStep), e->Jx, e->Jy, e->Jz, 1.f, 0.f, cosf(theta), sinf(theta), e->PulseGamma, e->AccelAmp, e->ClickB0, e->ClickA1, e->ClickA2, 0.f, 0.f};
    b->Impacts[b->NImpacts++] = im;
    b->Ringing[e->Object] = 1;
}

static void drain_events(OrBank *b) {
    uint32_t read = b->EventRead;
    for (; read != b->EventWrite; ++read) {
        const OrEvent *e = &b->Events[read % OR_EVENT_CAPACITY];
        if (e->Object >= b->NObjects) continue;
        if (e->Kind == 0) {
            if (e->PulseStep > 0) activate_impact(b, e);
        } else if (e->Kind == 1) {
            silence_object(b, e->Object);
        }
    }
    b->EventRead = read;
}

/* ModalAudio.cpp:86-147 */
static void render_object_fast(OrBank *b, uint32_t o, const uint32_t *impacts, uint32_t n_imp, float *out, uint32_t frame_count) {
    const uint32_t k0 = b->ModeOffset[o], stride = b->ModeCount[o];
    const uint32_t count = (n_imp == 0 && b->Cull) ? b->LiveModeCount[o] : b->TunedModeCount[o];
    const uint32_t shape0 = b->ShapeOffset[o];
    const float out_gain = b->OutGain[o];
    const float mix_gain = out_gain * b->ListenerGain[o];
    if ((size_t)n_imp * OR_LANES > b->GainsCap) {
        b->GainsCap = (size_t)n_imp * OR_LANES;
        b->Gains = (float *)realloc(b->Gains, b->GainsCap * sizeof(float));
    }
    float energy = 0.f;
    uint32_t live = 0;
    for (uint32_t k = 0; k < count; k += OR_LANES) {
        const uint32_t width = OR_LANES < count - k ? OR_LANES : count - k;
        float z_re[OR_LANES] = {0}, z_im[OR_LANES] = {0}, c_re[OR_LANES] = {0}, c_im[OR_LANES] = {0}, p_re[OR_LANES] = {0}, p_im[OR_LANES] = {0};
        for (uint32_t l = 0; l < width; ++l) {
            z_re[l] = b->StateRe.p[k0 + k + l], z_im[l] = b->StateIm.p[k0 + k + l];
            c_re[l] = b->CoeffRe.p[k0 + k + l], c_im[l] = b->CoeffIm.p[k0 + k + l];
            p_im[l] = b->OutPhaseIm.p[k0 + k + l], p_re[l] = b->OutPhaseRe.p[k0 + k + l];
        }
        for (uint32_t t = 0; t < n_imp; ++t) {
            float *gain = b->Gains + (size_t)t * OR_LANES;
            const OrImpact *im = &b->Impacts[impacts[t]];
            const uint32_t base = shape0 + im->ExPos * stride + k; /* ImpactGainRow, ModalAudio.h:182-188 */
            for (uint32_t i = 0; i < width; ++i) {
                gain[i] = b->RadiationGain.p[k0 + k + i] * (b->ShapeX.p[base + i] * im->Jx + b->ShapeY.p[base + i] * im->Jy + b->ShapeZ.p[base + i] * im->Jz);
            }
            for (uint32_t i = width; i < OR_LANES; ++i) gain[i] = 0.f;
        }
        for (uint32_t s = 0; s < frame_count; ++s) {
            float excite[OR_LANES] = {0};
            for (uint32_t t = 0; t < n_imp; ++t) {
                const float force = b->ForceScratch[(size_t)impacts[t] * frame_count + s];
                if (force == 0.f) continue;
                const float *gain = b->Gains + (size_t)t * OR_LANES;
                for (uint32_t l = 0; l < OR_LANES; ++l) excite[l] += force * gain[l];
            }
            float acc = 0.f;
            for (uint32_t l = 0; l < OR_LANES; ++l) {
                const float re = z_re[l] * c_re[l] - z_im[l] * c_im[l] + excite[l];
                z_im[l] = z_re[l] * c_im[l] + z_im[l] * c_re[l];
                z_re[l] = re;
                acc += p_im[l] * z_im[l] + p_re[l] * re;
            }
            out[s] += acc * mix_gain;
        }
        float chunk = 0.f;
        for (uint32_t l = 0; l < width; ++l) {
            b->StateRe.p[k0 + k + l] = z_re[l], b->StateIm.p[k0 + k + l] = z_im[l];
            chunk += z_re[l] * z_re[l] + z_im[l] * z_im[l];
        }
        energy += chunk;
        if (chunk * out_gain * out_gain >= kSilentEnergy) live = k + width;
    }
    if (!b->Cull) {
        b->Ringing[o] = 1;
        return;
    }
    if (n_imp == 0 && energy * out_gain * out_gain < kSilentEnergy) {
        silence_object(b, o);
        return;
    }
    b->Ringing[o] = 1;
    b->LiveModeCount[o] = n_imp == 0 ? live : b->TunedModeCount[o];
}

/* ModalAudio.cpp:486-590, single renderer (DealObjects with count == 1 is bank order, :446-449). Adds into out. */
void or_bank_render(OrBank *b, float *out, uint32_t frame_count) {
    if (frame_count == 0) return;
    if (b->FlushEvents) {
        b->FlushEvents = 0;
        b->EventRead = b->EventWrite;
    }
    drain_events(b);
    const uint32_t impact_count = b->NImpacts;
    if ((size_t)impact_count * frame_count > b->ForceCap) {
        b->ForceCap = (size_t)impact_count * frame_count;
        b->ForceScratch = (float *)realloc(b->ForceScratch, b->ForceCap * sizeof(float));
    }
    for (uint32_t i = 0; i < impact_count; ++i) {
        OrImpact *im = &b->Impacts[i];
        float phase_re = im->PhaseRe, phase_im = im->PhaseIm;
        const float rot_re = im->RotRe, rot_im = im->RotIm, gamma = im->Gamma, amp = im->AccelAmp;
        const float b0 = im->ClickB0, a1 = im->ClickA1, a2 = im->ClickA2;
        const float impact_click_gain = b->ClickGain * b->ListenerGain[im->Object];
        float z1 = im->ClickZ1, z2 = im->ClickZ2;
        uint32_t left = im->SamplesLeft;
        float *force = b->ForceScratch + (size_t)i * frame_count;
        for (uint32_t s = 0; s < frame_count; ++s) {
            float cur = 0.f;
            if (left > 0) {
                const float re = phase_re * rot_re - phase_im * rot_im;
                phase_im = phase_re * rot_im + phase_im * rot_re;
                phase_re = re;
                cur = gamma * 0.5f * (1.f - phase_re);
                --left;
            }
            force[s] = cur;
            const float u = amp * cur;
            const float y = b0 * u + z1;
            z1 = -a1 * y + z2;
            z2 = -b0 * u - a2 * y;
            out[s] += y * impact_click_gain;
        }
        im->PhaseRe = phase_re, im->PhaseIm = phase_im, im->SamplesLeft = left, im->ClickZ1 = z1, im->ClickZ2 = z2;
    }
    /* One renderer: its Out buffer starts at zero and is added to `out` afterwards (:543,:553-555). */
    float *mix = (float *)calloc(frame_count, sizeof(float));
    if (impact_count > b->ObjImpactsCap) {
        b->ObjImpactsCap = impact_count;
        b->ObjImpacts = (uint32_t *)realloc(b->ObjImpacts, impact_count * sizeof(uint32_t));
    }
    for (uint32_t o = 0; o < b->NObjects; ++o) {
        if (!b->Ringing[o]) continue;
        uint32_t n_imp = 0;
        for (uint32_t i = 0; i < impact_count; ++i) {
            if (b->Impacts[i].Object == o) b->ObjImpacts[n_imp++] = i;
        }
        render_object_fast(b, o, b->ObjImpacts, n_imp, mix, frame_count);
    }
    for (uint32_t s = 0; s < frame_count; ++s) out[s] += mix[s];
    free(mix);
    for (uint32_t i = b->NImpacts; i-- > 0;) {
        const OrImpact *im = &b->Impacts[i];
        if (im->SamplesLeft == 0 && fabsf(im->ClickZ1) + fabsf(im->ClickZ2) < 1e-12f) remove_impact(b, i);
    }
}

/* ---- FP64 arbiter -------------------------------------------------------------------------------------------------
 * The recurrence of RenderObjectFast (ModalAudio.cpp:115-131) evaluated in double over the SAME float32 parameters
 * (coefficients, gains, output rotation, the float force samples of the stage above) and summed into a double mix.
 * It answers one question only: when two float32 renderings of a long, undamped timeline differ by ~1e-5 of peak,
 * which of them moved away from the recurrence's exact value. Not the reference's arithmetic, so never the parity oracle.
 * Audibility culling follows the same rules on the double energies. */
static void exact_states(OrBank *b) {
    if (b->ExactN == b->CoeffRe.n) return;
    b->ExactRe = (double *)realloc(b->ExactRe, b->CoeffRe.n * sizeof(double));
    b->ExactIm = (double *)realloc(b->ExactIm, b->CoeffRe.n * sizeof(double));
    for (size_t i = b->ExactN; i < b->CoeffRe.n; ++i) b->ExactRe[i] = b->ExactIm[i] = 0.0;
    b->ExactN = b->CoeffRe.n;
}

static void silence_object_exact(OrBank *b, uint32_t o) {
    const uint32_t k0 = b->ModeOffset[o], count = b->ModeCount[o];
    for (uint32_t k = 0; k < count; ++k) b->ExactRe[k0 + k] = b->ExactIm[k0 + k] = 0.0;
    silence_object(b, o);
}

static void render_object_exact(OrBank *b, uint32_t o, const uint32_t *impacts, uint32_t n_imp, double *out, uint32_t frame_count) {
    const uint32_t k0 = b->ModeOffset[o], stride = b->ModeCount[o];
    const uint32_t count = (n_imp == 0 && b->Cull) ? b->LiveModeCount[o] : b->TunedModeCount[o];
    const uint32_t shape0 = b->ShapeOffset[o];
    const double out_gain = b->OutGain[o];
    const double mix_gain = (double)(b->OutGain[o] * b->ListenerGain[o]);
    double energy = 0.0;
    uint32_t live = 0;
    float *gains = (float *)malloc((size_t)(n_imp ? n_imp : 1) * sizeof(float));
    for (uint32_t k = 0; k < count; ++k) {
        const size_t m = (size_t)k0 + k;
        for (uint32_t t = 0; t < n_imp; ++t) {
            const OrImpact *im = &b->Impacts[impacts[t]];
            const uint32_t base = shape0 + im->ExPos * stride + k;
            gains[t] = b->RadiationGain.p[m] * (b->ShapeX.p[base] * im->Jx + b->ShapeY.p[base] * im->Jy + b->ShapeZ.p[base] * im->Jz);
        }
        double zr = b->ExactRe[m], zi = b->ExactIm[m];
        const double cr = b->CoeffRe.p[m], ci = b->CoeffIm.p[m], pi = b->OutPhaseIm.p[m], pr = b->OutPhaseRe.p[m];
        for (uint32_t s = 0; s < frame_count; ++s) {
            double e = 0.0;
            for (uint32_t t = 0; t < n_imp; ++t) e += (double)b->ForceScratch[(size_t)impacts[t] * frame_count + s] * (double)gains[t];
            const double re = zr * cr - zi * ci + e;
            zi = zr * ci + zi * cr;
            zr = re;
            out[s] += (pi * zi + pr * zr) * mix_gain;
        }
        b->ExactRe[m] = zr, b->ExactIm[m] = zi;
    }
    free(gains);
    for (uint32_t k = 0; k < count; k += OR_LANES) {
        const uint32_t width = OR_LANES < count - k ? OR_LANES : count - k;
        double chunk = 0.0;
        for (uint32_t l = 0; l < width; ++l) chunk += b->ExactRe[k0 + k + l] * b->ExactRe[k0 + k + l] + b->ExactIm[k0 + k + l] * b->ExactIm[k0 + k + l];
        energy += chunk;
        if (chunk * out_gain * out_gain >= (double)kSilentEnergy) live = k + width;
    }
    if (!b->Cull) {
        b->Ringing[o] = 1;
        return;
    }
    if (n_imp == 0 && energy * out_gain * out_gain < (double)kSilentEnergy) {
        silence_object_exact(b, o);
        return;
    }
    b->Ringing[o] = 1;
    b->LiveModeCount[o] = n_imp == 0 ? live : b->TunedModeCount[o];
}

/* Same block structure as or_bank_render; the force / click stage stays the reference's float recurrence (it IS the input). */
void or_bank_render_exact(OrBank *b, double *out, uint32_t frame_count) {
    if (frame_count == 0) return;
    exact_states(b);
    if (b->FlushEvents) {
        b->FlushEvents = 0;
        b->EventRead = b->EventWrite;
    }
    /* drain_events with the double states silenced alongside */
    for (uint32_t read = b->EventRead; read != b->EventWrite; ++read) {
        const OrEvent *e = &b->Events[read % OR_EVENT_CAPACITY];
        if (e->Object < b->NObjects && e->Kind == 1) silence_object_exact(b, e->Object);
    }
    drain_events(b);
    const uint32_t impact_count = b->NImpacts;
    if ((size_t)impact_count * frame_count > b->ForceCap) {
        b->ForceCap = (size_t)impact_count * frame_count;
        b->ForceScratch = (float *)realloc(b->ForceScratch, b->ForceCap * sizeof(float));
    }
    for (uint32_t i = 0; i < impact_count; ++i) {
        OrImpact *im = &b->Impacts[i];
        float phase_re = im->PhaseRe, phase_im = im->PhaseIm;
        const float rot_re = im->RotRe, rot_im = im->RotIm, gamma = im->Gamma, amp = im->AccelAmp;
        const float b0 = im->ClickB0, a1 = im->ClickA1, a2 = im->ClickA2;
        const float impact_click_gain = b->ClickGain * b->ListenerGain[im->Object];
        float z1 = im->ClickZ1, z2 = im->ClickZ2;
        uint32_t left = im->SamplesLeft;
        float *force = b->ForceScratch + (size_t)i * frame_count;
        for (uint32_t s = 0; s < frame_count; ++s) {
            float cur = 0.f;
            if (left > 0) {
                const float re = phase_re * rot_re - phase_im * rot_im;
                phase_im = phase_re * rot_im + phase_im * rot_re;
                phase_re = re;
                cur = gamma * 0.5f * (1.f - phase_re);
                --left;
            }
            force[s] = cur;
            const float u = amp * cur;
            const float y = b0 * u + z1;
            z1 = -a1 * y + z2;
            z2 = -b0 * u - a2 * y;
            out[s] += (double)(y * impact_click_gain);
        }
        im->PhaseRe = phase_re, im->PhaseIm = phase_im, im->SamplesLeft = left, im->ClickZ1 = z1, im->ClickZ2 = z2;
    }
    if (impact_count > b->ObjImpactsCap) {
        b->ObjImpactsCap = impact_count;
        b->ObjImpacts = (uint32_t *)realloc(b->ObjImpacts, impact_count * sizeof(uint32_t));
    }
    for (uint32_t o = 0; o < b->NObjects; ++o) {
        if (!b->Ringing[o]) continue;
        uint32_t n_imp = 0;
        for (uint32_t i = 0; i < impact_count; ++i) {
            if (b->Impacts[i].Object == o) b->ObjImpacts[n_imp++] = i;
        }
        render_object_exact(b, o, b->ObjImpacts, n_imp, out, frame_count);
    }
    for (uint32_t i = b->NImpacts; i-- > 0;) {
        const OrImpact *im = &b->Impacts[i];
        if (im->SamplesLeft == 0 && fabsf(im->ClickZ1) + fabsf(im->ClickZ2) < 1e-12f) remove_impact(b, i);
    }
}

uint32_t or_bank_mode_total(const OrBank *b) { return (uint32_t)b->CoeffRe.n; }
uint32_t or_bank_object_count(const OrBank *b) { return b->NObjects; }
uint32_t or_bank_active_impacts(const OrBank *b) { return b->NImpacts; }
uint64_t or_bank_events_dropped(const OrBank *b) { return b->EventsDropped; }

/* Same column numbering as oracle/ref_audio_driver.cpp. */
void or_bank_get_mode_column(const OrBank *b, uint32_t which, float *out) {
    const FVec *cols[] = {&b->CoeffRe, &b->CoeffIm, &b->StateRe, &b->StateIm, &b->RadiationGain, &b->RadiationArea, &b->OutPhaseIm, &b->OutPhaseRe, &b->DeflectionGain, &b->QuadCompliance, &b->QuadDriveScale};
    if (which < sizeof cols / sizeof *cols) memcpy(out, cols[which]->p, cols[which]->n * sizeof(float));
}
void or_bank_get_object_column_u32(const OrBank *b, uint32_t which, uint32_t *out) {
    for (uint32_t o = 0; o < b->NObjects; ++o) {
        out[o] = which == 0 ? b->ModeOffset[o] : which == 1 ? b->ModeCount[o] : which == 2 ? b->TunedModeCount[o] : which == 3 ? b->LiveModeCount[o] : which == 4 ? b->Ringing[o] : 0;
    }
}
void or_bank_get_object_column_f32(const OrBank *b, uint32_t which, float *out) {
    const float *cols[] = {b->RadiantRadius, b->DeflectionScale, b->OutGain, b->ListenerGain};
    if (which < 4) memcpy(out, cols[which], b->NObjects * sizeof(float));
}
