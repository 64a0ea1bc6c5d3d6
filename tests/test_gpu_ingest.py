"""-m gpu: device WAV ingest + FFT resampler (csrc/resample.cu) against scipy.signal.resample -- the call
decode_wav_file makes (decoder.py:385-387) -- and decode_wav_file on 48 kHz / stereo WAV files against the oracle chain."""
import wave

import numpy as np
import pytest
from scipy import signal

from oracle import modem_v2 as o2, signals as sig
from oracle.frames import parse_fbp_stream

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,num", [(48000, 96000), (44100, 96000), (96001, 96000), (50001, 100001), (1001, 2003), (4000, 1000),
                                   (123456, 246912), (8, 16), (7, 3)])
def test_resample_matches_scipy(n, num, engine):
    rng = np.random.default_rng(n + num)
    pcm = rng.integers(-20000, 20000, n).astype(np.int16)
    want = signal.resample(pcm.astype(np.float64) / 32768.0, num)
    got = engine.ingest_resample(pcm, num)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))     # two FFT libraries: last-ulp differences only


def test_ingest_channel_select_and_dtypes(engine):
    rng = np.random.default_rng(3)
    st = rng.integers(-30000, 30000, (5000, 2)).astype(np.int16)
    assert np.array_equal(engine.ingest_resample(st, 5000), st[:, 0].astype(np.float64) / 32768.0)    # data[:, 0], exact
    f = rng.standard_normal(4096).astype(np.float32)
    assert np.array_equal(engine.ingest_resample(f, 4096), f.astype(np.float64))
    d = rng.standard_normal((777, 3))
    assert np.array_equal(engine.ingest_resample(d, 777), d[:, 0])


def _write_wav(path, x, sr, stereo=False):
    pcm = np.clip(np.round(x * 32767), -32768, 32767).astype("<i2")
    if stereo:
        pcm = np.stack([pcm, (pcm // 3).astype("<i2")], axis=1).reshape(-1)
    with wave.open(path, "wb") as w:
        w.setnchannels(2 if stereo else 1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(pcm.tobytes())


@pytest.mark.parametrize("sr,stereo", [(48000, False), (48000, True), (96000, True)])
def test_decode_wav_file_resampled(sr, stereo, tmp_path, engine, monkeypatch):
    """A DQPSK recording stored at 48 kHz (the reference's recorder rate) / in stereo: decode_wav_file == the reference
    chain (PCM16 / 32768 -> channel 0 -> scipy.signal.resample -> qpsk_demodulate -> parse), frame recovered."""
    from fbdsp import decoder as fbd
    monkeypatch.chdir(tmp_path)
    _, framed, x96 = sig.kat_signal(sig.qpsk_modulate, 6000 + sr // 1000, 1200, 25, baud=3000, carrier=3000.0)
    x = signal.resample(x96.astype(np.float64), len(x96) * sr // 96000) if sr != 96000 else x96.astype(np.float64)
    x = 0.8 * x / np.max(np.abs(x))
    path = str(tmp_path / "part.wav")
    _write_wav(path, x, sr, stereo)
    # the reference chain, on the host
    with wave.open(path, "rb") as w:
        raw = np.frombuffer(w.readframes(w.getnframes()), "<i2")
    ch0 = (raw.reshape(-1, 2)[:, 0] if stereo else raw).astype(np.float64) / 32768.0
    ref_in = signal.resample(ch0, int(round(len(ch0) * 96000.0 / sr))) if sr != 96000 else ch0
    want_raw = o2.qpsk_demodulate(ref_in, 3000, 3000.0)
    want = [f["data"] for f in parse_fbp_stream(want_raw)]
    assert len(want) == 1                                      # this parameter pair round-trips
    files = fbd.decode_wav_file(path, "QPSK", 3000)
    assert len(files) == 1
    from fbdsp.decoder import _decompress
    assert open(files[0], "rb").read() == _decompress(want[0])


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 8, 64, 360, 2520, 4096, 4099, 48000, 96000, 100003, 1234567, 8640000])
def test_fft_matches_numpy(n, engine):
    """The hand-written float64 FFT behind scipy.signal.hilbert / scipy.signal.resample (csrc/fft.cu): Stockham passes for
    2-3-5-7-smooth lengths, Bluestein for the others (4099 and 100003 are prime, 1234567 = 127 x 9721), forward and inverse,
    against numpy's pocketfft to 1e-12 of the largest bin."""
    import ctypes
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    out = np.empty(n, dtype=np.complex128)
    fn = engine.lib.fb_debug_fft_c2c
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
    for sign, ref in ((-1, np.fft.fft(x)), (1, np.fft.ifft(x) * n)):
        assert fn(engine.handle, x.ctypes.data, out.ctypes.data, n, sign) == 0
        assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
