"""-m gpu: the corpus driver (decoder.decode_corpus: reference dispatch decoder.py:422-434 per recording, one batched launch
sequence per parameter set), byte parity on a FULL-LENGTH 3-minute config-2 recording, WAV encodings, and two host threads
sharing the process-wide engine."""
import os
import struct
import threading
import wave

import numpy as np
import pytest

from oracle import modem_v2 as o2, signals as sig
from oracle.frames import frame_data, parse_fbp_stream

pytestmark = pytest.mark.gpu
FS = 96000


def _noisy(x, snr_db, rng, lead=0):
    x = np.concatenate([np.zeros(lead, np.float64), np.asarray(x, np.float64)])
    return (x + rng.standard_normal(len(x)) * np.sqrt(np.mean(x * x) / 10 ** (snr_db / 10))).astype(np.float32)


def _corpus(n, seed):
    """Recordings of mixed (mode, rate, carrier | tones) incl. the FSK product defaults (reference: ValueError -> [])."""
    rng = np.random.default_rng(seed)
    kinds = [("QPSK", 9600, 9600.0, None), ("8PSK", 9600, 19200.0, None), ("OFDM4", 4800, 9600.0, None), ("BPSK", 4800, 9600.0, None),
             ("QPSK", 9600, None, None), ("FSK1200", 1200, None, (2400.0, 4800.0)), ("FSK9600", 9600, None, (12000.0, 24000.0)),
             ("FSK1200", 1200, None, None), ("FSK19200", 19200, None, None), ("QPSK", 3000, 3000.0, None)]
    recs, modes, rates, carriers, tones, want = [], [], [], [], [], []
    for i in range(n):
        mode, rate, car, tn = kinds[i % len(kinds)]
        nbytes = int(rng.integers(40, 200)) if rate <= 1200 else int(rng.integers(300, 2500))
        payload = rng.integers(0, 256, nbytes, dtype=np.uint8).tobytes()
        framed = frame_data(f"r{i}.bin", payload, 0, 1, nbytes, 0)
        if mode.startswith("FSK"):
            baud = {"FSK1200": 1200, "FSK9600": 9600, "FSK19200": 19200}[mode]
            if tn is None:
                x = (rng.standard_normal(int(rng.integers(5000, 30000))) * 0.1).astype(np.float32)
                w = None                                                     # reference raises
            else:
                x = _noisy(sig.fsk_modulate(framed, baud=baud, mark_freq=tn[0], space_freq=tn[1]), 20, rng)
                w = o2.fsk_demodulate(x, baud, tn[0], tn[1])
        else:
            c = 3000.0 if car is None else car
            sps = FS // rate
            mod = sig.bpsk_modulate if mode == "BPSK" else sig.qpsk_modulate
            x = _noisy(mod(framed, baud=rate, carrier=c), 20, rng, lead=int(rng.integers(0, 20000)) // sps * sps)
            w = (o2.bpsk_demodulate if mode == "BPSK" else o2.qpsk_demodulate)(x, rate, c)
        recs.append(x); modes.append(mode); rates.append(rate); carriers.append(car); tones.append(tn); want.append(w)
    return recs, modes, rates, carriers, tones, want


def test_decode_corpus_mixed_parameter_sets(engine):
    from fbdsp.decoder import corpus_groups, decode_corpus
    recs, modes, rates, carriers, tones, want = _corpus(30, 5100)
    launches0 = engine.kernel_launches
    out = decode_corpus(recs, modes, rates, engine=engine, carriers=carriers, tones=tones)
    n_groups = len(corpus_groups(modes, rates, carriers, tones, [r.dtype for r in recs]))
    # one launch sequence per group, not per recording -- except the Hilbert envelope of a decodable FSK recording, whose two
    # whole-record transforms per tone are ~15 Stockham passes each (csrc/fft.cu)
    n_fsk = sum(1 for m, w in zip(modes, want) if m.startswith("FSK") and w is not None)
    assert n_groups == 10 and engine.kernel_launches - launches0 < 40 * n_groups + 100 * n_fsk
    n_ok = 0
    for i, (r, w) in enumerate(zip(out, want)):
        if w is None:
            assert r.error and r.error.startswith("ValueError") and r.frames == [] and r.raw == b"", i
            with pytest.raises(ValueError) as e:
                o2.fsk_demodulate(recs[i], 1200 if modes[i] == "FSK1200" else 19200 if modes[i] == "FSK19200" else 9600)
            assert r.error == f"ValueError: {e.value}"
            continue
        assert r.error is None and r.raw == w, (i, modes[i], rates[i])
        assert [f["data"] for f in r.frames] == [f["data"] for f in parse_fbp_stream(w)]
        n_ok += len(r.frames)
    assert n_ok >= 12
    # order invariance: the same corpus shuffled gives the same per-recording results
    perm = np.random.default_rng(1).permutation(len(recs))
    out2 = decode_corpus([recs[i] for i in perm], [modes[i] for i in perm], [rates[i] for i in perm], engine=engine,
                         carriers=[carriers[i] for i in perm], tones=[tones[i] for i in perm])
    for j, i in enumerate(perm):
        assert out2[j].raw == out[i].raw and out2[j].error == out[i].error


def test_decode_corpus_short_and_empty(engine):
    from fbdsp.decoder import decode_corpus
    rng = np.random.default_rng(3)
    recs = [rng.standard_normal(n).astype(np.float32) for n in (0, 5, 27, 28, 300, 4000)]
    out = decode_corpus(recs, ["QPSK"] * 6, [9600] * 6, engine=engine)
    for x, r in zip(recs, out):
        try:
            w = o2.qpsk_demodulate(x, 9600, 3000.0)
            assert r.error is None and r.raw == w, len(x)
        except ValueError as e:
            assert r.error == f"ValueError: {e}", len(x)


@pytest.mark.parametrize("fmt", ["f32", "pcm16"])
def test_full_length_config2_recording_bytes(engine, fmt):
    """ONE full 3-minute recording of BASELINE config 2 (17.28 M samples; the bench batch is 256 of these): raw bytes and
    sync index against the oracle's float64 chain, not only CRC-valid payloads."""
    import fbdsp
    n = 180 * FS
    rng = np.random.default_rng(1000)
    payload = rng.integers(0, 256, 431000, dtype=np.uint8).tobytes()
    framed = frame_data("part0.bin", payload, 0, 256, 431000 * 256, 0)
    x = sig.qpsk_modulate(framed, baud=9600, carrier=9600.0).astype(np.float64)
    x = np.concatenate([x, np.zeros(max(0, n - len(x)))])[:n]
    x = (x + rng.standard_normal(n) * np.sqrt(np.mean(x * x) / 100.0)).astype(np.float32)
    d = fbdsp.psk_design(9600.0, 9600.0, float(FS), 1.5, False)
    if fmt == "pcm16":
        pcm = np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)
        res = engine.psk_demod_batch([pcm], d)[0]
        st = o2.qpsk_stages(pcm.astype(np.float64) / 32768.0, 9600, 9600.0)
    else:
        res = engine.psk_demod_batch([x], d)[0]
        st = o2.qpsk_stages(x, 9600, 9600.0)
    assert res.sync_idx == st["sync"] and len(res.raw) == len(st["raw"]) and len(res.raw) > 431900
    assert res.raw == st["raw"]
    assert np.array_equal(engine.last_bits(0), st["bits"])
    (f,) = parse_fbp_stream(res.raw)
    assert f["data"] == payload


def _write_wav(path, frames, sr, width, fmt_tag=1):
    """frames: float array in [-1, 1), shape (n,) or (n, ch)."""
    a = np.atleast_2d(frames.T).T
    nch = a.shape[1]
    if fmt_tag == 3:
        body = a.astype("<f4" if width == 4 else "<f8").tobytes()
    elif width == 1:
        body = (np.round(a * 128.0) + 128).clip(0, 255).astype(np.uint8).tobytes()
    elif width == 2:
        body = np.round(a * 32768.0).clip(-32768, 32767).astype("<i2").tobytes()
    elif width == 3:
        v = np.round(a * 8388608.0).clip(-8388608, 8388607).astype(np.int32).reshape(-1)
        body = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).tobytes()
    else:
        body = np.round(a * 2147483648.0).clip(-2147483648, 2147483647).astype("<i4").tobytes()
    fmt = struct.pack("<HHIIHH", fmt_tag, nch, sr, sr * nch * width, nch * width, 8 * width)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"data" + struct.pack("<I", len(body)) + body)


@pytest.mark.parametrize("width,tag", [(1, 1), (2, 1), (3, 1), (4, 1), (4, 3), (8, 3)])
def test_wav_encodings_match_soundfile_scaling(engine, tmp_path, width, tag):
    """decode_wav_file on every WAV encoding soundfile.read accepts (decoder.py:381): the demodulator must see the
    float64 values soundfile would return."""
    from fbdsp import decoder
    _, framed, x = sig.kat_signal(sig.qpsk_modulate, 6100 + width, 600, 25, baud=9600, carrier=3000.0)
    x = np.clip(x.astype(np.float64) * 0.5, -0.999, 0.999)
    p = str(tmp_path / f"w{width}_{tag}.wav")
    _write_wav(p, x, FS, width, tag)
    got, sr = decoder.read_wav(p)
    assert sr == FS
    if tag == 3:
        want = x.astype(np.float32 if width == 4 else np.float64).astype(np.float64)
    else:
        scale = {1: 128.0, 2: 32768.0, 3: 8388608.0, 4: 2147483648.0}[width]
        q = np.round(x * scale).clip(-scale, scale - 1)
        want = q / scale
    if width == 2 and tag == 1:
        assert got.dtype == np.int16 and np.array_equal(got.astype(np.float64) / 32768.0, want)
    else:
        assert got.dtype == np.float64 and np.array_equal(got, want)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        from fbdsp import modem
        data = decoder._wav_samples(p)
        raw = modem.qpsk_demodulate(decoder._Pcm16(data) if data.dtype == np.int16 else data, 9600)
        assert raw == o2.qpsk_demodulate(want, 9600, 3000.0)
        assert isinstance(decoder.decode_wav_file(p, "QPSK", 9600), list)
    finally:
        os.chdir(cwd)


def test_two_threads_share_the_default_engine():
    """The reference calls decode_from_buffer from the GUI thread and a QThread (filebeep_advanced_v2.py:324,1112): two host
    threads hammering the process-wide engine with different parameter sets must each get the oracle's bytes."""
    from fbdsp import modem
    cases = []
    for seed, baud, car, mod, dem, odem in [(7001, 9600, 9600.0, sig.qpsk_modulate, modem.qpsk_demodulate, o2.qpsk_demodulate),
                                            (7002, 4800, 9600.0, sig.bpsk_modulate, modem.bpsk_demodulate, o2.bpsk_demodulate),
                                            (7003, 3000, 3000.0, sig.qpsk_modulate, modem.qpsk_demodulate, o2.qpsk_demodulate)]:
        _, _, x = sig.kat_signal(mod, seed, 1200, 20, baud=baud, carrier=car)
        cases.append((dem, x, baud, car, odem(x, baud, car)))
    errors = []

    def worker(k):
        try:
            for it in range(12):
                dem, x, baud, car, want = cases[(k + it) % len(cases)]
                if dem(x, baud, car) != want:
                    errors.append((k, it, "bytes differ"))
        except Exception as e:      # noqa: BLE001
            errors.append((k, repr(e)))

    th = [threading.Thread(target=worker, args=(k,)) for k in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert errors == []
