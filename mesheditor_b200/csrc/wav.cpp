// Rendered audio out: the file WriteWav produces (src/audio/AudioSystem.cpp:1244-1250 over src/audio/WavWriter.h) - mono
// 32-bit IEEE-float RIFF/WAVE, optionally scaled so that the largest sample lands on `normalize_max`. The reference writes it
// through CoreAudio (macOS only); the container is the standard one, so this is a plain encoder / decoder of that layout:
// "fmt " (format 3, 1 channel, 32 bits), "fact" (frame count, required for non-PCM formats), "data". Host code.
#include "common.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace me {
namespace {
void Put32(std::vector<uint8_t> &b, uint32_t v) {
    for (int i = 0; i < 4; ++i) b.push_back(uint8_t(v >> (8 * i)));
}
void Put16(std::vector<uint8_t> &b, uint16_t v) { b.push_back(uint8_t(v)), b.push_back(uint8_t(v >> 8)); }
void Tag(std::vector<uint8_t> &b, const char *t) { b.insert(b.end(), t, t + 4); }
uint32_t Get32(const uint8_t *p) { return uint32_t(p[0]) | uint32_t(p[1]) << 8 | uint32_t(p[2]) << 16 | uint32_t(p[3]) << 24; }
uint16_t Get16(const uint8_t *p) { return uint16_t(p[0] | p[1] << 8); }
} // namespace
} // namespace me

using namespace me;

extern "C" {

MeStatus me_wav_encode(const float *frames, uint64_t n, uint32_t sample_rate, float normalize_max, uint8_t **bytes, uint64_t *size) {
    return Guard([&] {
        if ((n && !frames) || !bytes || !size) Fail(ME_BAD_ARG, "null argument");
        if (!sample_rate) Fail(ME_BAD_ARG, "sample_rate must be positive");
        if (n > (0xFFFFFFFFull - 64) / 4) Fail(ME_BAD_ARG, "more frames than a RIFF file holds");
        // WriteWav's normalisation: by the largest sample (not the largest magnitude), as the reference does.
        const float scale = normalize_max > 0 && n ? normalize_max / *std::max_element(frames, frames + n) : 1.f;
        std::vector<uint8_t> b;
        b.reserve(56 + 4 * n);
        Tag(b, "RIFF"), Put32(b, uint32_t(48 + 4 * n)), Tag(b, "WAVE");
        Tag(b, "fmt "), Put32(b, 16), Put16(b, 3), Put16(b, 1), Put32(b, sample_rate), Put32(b, sample_rate * 4), Put16(b, 4), Put16(b, 32);
        Tag(b, "fact"), Put32(b, 4), Put32(b, uint32_t(n));
        Tag(b, "data"), Put32(b, uint32_t(4 * n));
        for (uint64_t i = 0; i < n; ++i) {
            const float v = frames[i] * scale;
            uint32_t bits;
            std::memcpy(&bits, &v, 4);
            Put32(b, bits);
        }
        auto *out = static_cast<uint8_t *>(std::malloc(b.size()));
        if (!out) Fail(ME_OUT_OF_MEMORY, "host allocation failed");
        std::memcpy(out, b.data(), b.size());
        *bytes = out, *size = b.size();
    });
}

MeStatus me_wav_decode(const uint8_t *bytes, uint64_t size, float **frames, uint64_t *n, uint32_t *sample_rate) {
    return Guard([&] {
        if (!bytes || !frames || !n || !sample_rate) Fail(ME_BAD_ARG, "null argument");
        if (size < 12 || std::memcmp(bytes, "RIFF", 4) || std::memcmp(bytes + 8, "WAVE", 4)) Fail(ME_BAD_ARG, "not a RIFF/WAVE file");
        uint16_t format = 0, channels = 0, bits = 0;
        uint32_t rate = 0;
        const uint8_t *data = nullptr;
        uint64_t data_size = 0;
        for (uint64_t at = 12; at + 8 <= size;) {
            const uint64_t len = Get32(bytes + at + 4), body = at + 8;
            if (body + len > size) Fail(ME_BAD_ARG, "chunk past the end of the file");
            if (!std::memcmp(bytes + at, "fmt ", 4)) {
                if (len < 16) Fail(ME_BAD_ARG, "short fmt chunk");
                format = Get16(bytes + body), channels = Get16(bytes + body + 2), rate = Get32(bytes + body + 4), bits = Get16(bytes + body + 14);
            } else if (!std::memcmp(bytes + at, "data", 4)) {
                data = bytes + body, data_size = len;
            }
            at = body + len + (len & 1); // chunks are word-aligned
        }
        if (!data || !rate) Fail(ME_BAD_ARG, "missing fmt or data chunk");
        if (channels != 1 || !((format == 3 && bits == 32) || (format == 1 && bits == 16))) Fail(ME_BAD_ARG, "only mono float32 or int16 files (format %u, %u channels, %u bits)", format, channels, bits);
        const uint64_t count = data_size / (bits / 8);
        auto *out = static_cast<float *>(std::malloc(std::max<uint64_t>(count, 1) * sizeof(float)));
        if (!out) Fail(ME_OUT_OF_MEMORY, "host allocation failed");
        for (uint64_t i = 0; i < count; ++i) {
            if (format == 3) {
                const uint32_t v = Get32(data + 4 * i);
                std::memcpy(out + i, &v, 4);
            } else {
                out[i] = float(int16_t(Get16(data + 2 * i))) / 32768.f;
            }
        }
        *frames = out, *n = count, *sample_rate = rate;
    });
}

} // extern "C"
