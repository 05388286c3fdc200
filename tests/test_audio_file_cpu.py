"""Rendered audio out (WriteWav over WavWriter, src/audio/AudioSystem.cpp:1244-1250): the reference's own tests/AudioFileTest.cpp
round trip restated against me_wav_encode / me_wav_decode, with scipy's reader as the independent check of the container (the
reference writes through CoreAudio, which exists on macOS only). Host-only."""
import io

import numpy as np
import pytest

import mesheditor_b200 as me
from mesheditor_b200 import MeError

pytestmark = pytest.mark.usefixtures("built_lib")


def test_float_wav_round_trip():
    """tests/AudioFileTest.cpp:16-41: 8,192 frames of a 257-step ramp at 48 kHz come back within 1e-7 (here: exactly)."""
    source = ((np.arange(8192) % 257).astype(np.int64) - 128).astype(np.float32) / np.float32(128)
    data = me.wav_bytes(source, 48000)
    decoded, rate = me.wav_frames(data)
    assert rate == 48000 and len(decoded) == len(source)
    assert float(np.abs(decoded - source).max()) <= 1e-7
    np.testing.assert_array_equal(decoded, source)
    assert len(data) == 56 + 4 * len(source) and data[:4] == b"RIFF" and int.from_bytes(data[4:8], "little") == len(data) - 8  # a well-formed RIFF


def test_container_reads_with_an_independent_decoder():
    from scipy.io import wavfile

    rng = np.random.default_rng(5)
    frames = rng.normal(0, 0.3, 4801).astype(np.float32)  # an odd count: no padding needed for 4-byte samples
    rate, read = wavfile.read(io.BytesIO(me.wav_bytes(frames, 44100)))
    assert rate == 44100 and read.dtype == np.float32
    np.testing.assert_array_equal(read, frames)
    # and the other way: scipy's float32 and int16 files decode here
    for written, want in ((frames, frames), ((frames * 20000).astype(np.int16), (frames * 20000).astype(np.int16).astype(np.float32) / np.float32(32768))):
        buf = io.BytesIO()
        wavfile.write(buf, 96000, written)
        decoded, rate = me.wav_frames(buf.getvalue())
        assert rate == 96000
        np.testing.assert_array_equal(decoded, want)


def test_write_wav_normalisation_and_errors():
    frames = np.array([0.1, -0.8, 0.4, 0.0], np.float32)
    scaled, _ = me.wav_frames(me.wav_bytes(frames, 48000, normalize_max=1.0))
    np.testing.assert_array_equal(scaled, frames * (np.float32(1.0) / np.float32(0.4)))  # WriteWav scales by the largest sample, not the largest magnitude
    empty, rate = me.wav_frames(me.wav_bytes([], 48000))
    assert len(empty) == 0 and rate == 48000
    for bad in (b"", b"RIFF\x04\x00\x00\x00WAVX", me.wav_bytes(frames, 48000)[:30]):
        with pytest.raises(MeError):
            me.wav_frames(bad)
    with pytest.raises(MeError):
        me.wav_bytes(frames, 0)
    from scipy.io import wavfile

    buf = io.BytesIO()
    wavfile.write(buf, 48000, np.zeros((10, 2), np.float32))  # stereo is not what the path renders
    with pytest.raises(MeError):
        me.wav_frames(buf.getvalue())
