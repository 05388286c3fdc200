// Impact spectrum analysis (src/audio/AudioSystem.cpp:492-560): the fundamental of a recorded impact, which the reference hands
// to the solve as SolverConfig::FundamentalFreq (:821-829) so that the model's first mode lands on the recording's. A
// Blackman-Harris-windowed segment shortly after the impact (frames 30 .. sample_rate/16), its spectrum in dB, and the first
// prominent peak above 50 Hz. Host code: one transform of ~3k samples per solve. The reference transforms with FFTW (single
// precision); here the segment is transformed by a direct double-precision DFT and rounded to float, then judged the same way.
#include "common.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <numbers>
#include <vector>

namespace me {
namespace {

// dB spectrum -> the first local maximum at or above the noise threshold (median of the upper half + 15 dB) that stands 10 dB
// or more above the mean of its +-15-bin neighbourhood. 0 when there is none.
bool FirstProminentPeak(const std::vector<float> &db, uint64_t n_real, uint32_t sample_rate, float *hz) {
    const size_t bins = db.size();
    constexpr size_t kReach = 15;
    if (bins <= 2 * kReach) return false;
    std::vector<float> upper(db.begin() + bins / 2, db.end());
    std::nth_element(upper.begin(), upper.begin() + upper.size() / 2, upper.end());
    const float threshold = upper[upper.size() / 2] + 15.f;
    const size_t lowest = size_t(50 * n_real / sample_rate);
    for (size_t i = std::max(lowest, kReach); i < bins - kReach; ++i) {
        if (db[i] <= db[i - 1] || db[i] <= db[i + 1] || db[i] < threshold) continue;
        float sum = 0;
        for (size_t j = i - kReach; j <= i + kReach; ++j) sum += db[j];
        if (db[i] - sum / float(2 * kReach + 1) >= 10.f) {
            *hz = float(i * sample_rate / n_real); // whole hertz: the reference divides integers
            return true;
        }
    }
    return false;
}

std::vector<float> DecibelSpectrum(const float *re_im, size_t bins) {
    std::vector<float> db(bins);
    for (size_t i = 0; i < bins; ++i) {
        const float power = re_im[2 * i] * re_im[2 * i] + re_im[2 * i + 1] * re_im[2 * i + 1];
        db[i] = 10.f * std::log10(std::max(power, 1e-20f));
    }
    return db;
}

} // namespace
} // namespace me

using namespace me;

extern "C" {

int me_estimate_fundamental_from_spectrum(const float *complex_re_im, uint64_t n_real, uint32_t sample_rate, float *hz) {
    if (!complex_re_im || !hz || !sample_rate || n_real < 2) return 0;
    return FirstProminentPeak(DecibelSpectrum(complex_re_im, size_t(n_real / 2 + 1)), n_real, sample_rate, hz);
}

MeStatus me_impact_spectrum(const float *frames, uint64_t n_frames, uint32_t sample_rate, float *complex_re_im, uint64_t *n_real) {
    return Guard([&] {
        if (!frames || !n_real) Fail(ME_BAD_ARG, "null argument");
        constexpr uint32_t first = 30; // FftStartFrame
        const uint32_t last = sample_rate / 16;
        if (last <= first + 1 || n_frames < last) Fail(ME_BAD_ARG, "the segment is frames 30 .. sample_rate/16: %u frames needed", last);
        const uint32_t n = last - first;
        *n_real = n;
        if (!complex_re_im) return; // size query
        // Blackman-Harris: sum_j c_j cos(2 pi i j / n), in float like the reference's window.
        constexpr float coeff[4] = {0.35875f, -0.48829f, 0.14128f, -0.01168f};
        std::vector<double> x(n);
        for (uint32_t i = 0; i < n; ++i) {
            float w = 0.f;
            for (uint32_t j = 0; j < 4; ++j) w += coeff[j] * float(std::cos(std::numbers::pi * double(float(2 * i * j) / float(n))));
            x[i] = double(w * frames[first + i]);
        }
        // X_k = sum_i x_i e^{-2 pi i k i / n}, k = 0 .. n/2, twiddles from one table indexed by (k i) mod n.
        std::vector<std::complex<double>> twiddle(n);
        for (uint32_t t = 0; t < n; ++t) twiddle[t] = std::polar(1.0, -2.0 * std::numbers::pi * double(t) / double(n));
        for (uint32_t k = 0; k <= n / 2; ++k) {
            std::complex<double> acc = 0;
            uint32_t phase = 0;
            for (uint32_t i = 0; i < n; ++i) {
                acc += x[i] * twiddle[phase];
                phase += k;
                if (phase >= n) phase -= n;
            }
            complex_re_im[2 * k] = float(acc.real()), complex_re_im[2 * k + 1] = float(acc.imag());
        }
    });
}

int me_estimate_fundamental(const float *frames, uint64_t n_frames, uint32_t sample_rate, float *hz) {
    uint64_t n = 0;
    if (!hz || me_impact_spectrum(frames, n_frames, sample_rate, nullptr, &n) != ME_OK) return 0;
    std::vector<float> spectrum(2 * (n / 2 + 1));
    if (me_impact_spectrum(frames, n_frames, sample_rate, spectrum.data(), &n) != ME_OK) return 0;
    return me_estimate_fundamental_from_spectrum(spectrum.data(), n, sample_rate, hz);
}

} // extern "C"
