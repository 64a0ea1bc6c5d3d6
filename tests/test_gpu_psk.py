"""-m gpu: the CUDA DPSK path through the C ABI against the golden fixtures and the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import modem_v2 as o2, signals as sig

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PSK_FNS = ("qpsk_demodulate", "psk8_demodulate", "ofdm_demodulate_simple", "bpsk_demodulate", "psk31_demodulate")


def _cases():
    cs = json.load(open(os.path.join(GOLD, "demod_cases.json")))["cases"]
    return [c for c in cs if c["fn"] in PSK_FNS]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_golden_bytes(case, golden, engine):
    """Bit-exact raw bytes (and the same exception type + text) as the reference on the stored inputs."""
    from fbdsp import modem
    _, arrs = golden
    x = arrs[case["name"] + ".x"]
    fn = getattr(modem, case["fn"])
    if case["exc"]:
        with pytest.raises(Exception) as ei:
            fn(x, *case["args"])
        assert type(ei.value).__name__ == case["exc"][0]
        assert str(ei.value) == case["exc"][1]
        return
    raw = fn(x, *case["args"])
    want = arrs[case["name"] + ".raw"].tobytes()
    assert len(raw) == len(want)
    assert raw == want


def _bits_report(got, st, bps):
    want = st["bits"]
    assert len(got) == len(want)
    bad = np.nonzero(got != want)[0]
    margin = (o2.qpsk_margin if bps == 2 else o2.bpsk_margin)(st["diff"])
    return bad, margin


@pytest.mark.parametrize("baud,carrier,snr,seed", [(9600, 9600.0, 20, 101), (9600, 3000.0, 20, 102), (9600, 9600.0, 5, 103),
                                                    (4800, 9600.0, 10, 104), (1200, 3000.0, 20, 105), (3000, 3000.0, 0, 106)])
def test_qpsk_decisions_vs_oracle(baud, carrier, snr, seed, engine):
    """Symbol decisions: >= 99.99 % equal, mismatches only where the oracle's margin < 1e-5 (north_star)."""
    import fbdsp
    nbytes = 20000 if baud >= 4800 else 2500
    _, _, x = sig.kat_signal(sig.qpsk_modulate, seed, nbytes, snr, baud=baud, carrier=carrier)
    d = fbdsp.psk_design(float(baud), float(carrier), 96000.0, 1.5, False)
    res = engine.psk_demod_batch([x], d)[0]
    st = o2.qpsk_stages(x, baud, carrier)
    bad, margin = _bits_report(engine.last_bits(0), st, 2)
    assert len(bad) <= 1e-4 * len(st["bits"])
    assert all(margin[b // 2] < 1e-5 for b in bad), (bad[:10], margin[bad[:10] // 2])
    if len(bad) == 0:
        assert res.raw == st["raw"] and res.sync_idx == st["sync"]


def test_bpsk_decisions_vs_oracle(engine):
    import fbdsp
    _, _, x = sig.kat_signal(sig.bpsk_modulate, 107, 6000, 10, baud=9600, carrier=3000.0)
    d = fbdsp.psk_design(9600.0, 3000.0, 96000.0, 1.0, True)
    res = engine.psk_demod_batch([x], d)[0]
    st = o2.bpsk_stages(x, 9600, 3000.0)
    bad, margin = _bits_report(engine.last_bits(0), st, 1)
    assert all(margin[b] < 1e-5 for b in bad)
    if len(bad) == 0:
        assert res.raw == st["raw"]


def test_ragged_batch_and_dtypes(engine):
    """One launch over recordings of very different lengths (incl. too-short and empty ones) == one by one."""
    import fbdsp
    from fbdsp import _lib
    rng = np.random.default_rng(7)
    d = fbdsp.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
    recs = []
    for i, n in enumerate([50000, 10, 28, 3000, 123457, 40, 2700, 6000, 99999]):
        if n > 5000:
            _, _, x = sig.kat_signal(sig.qpsk_modulate, 200 + i, n // 45, 15, baud=9600, carrier=9600.0)
            x = x[:n]
        else:
            x = (rng.standard_normal(n) * 0.2).astype(np.float32)
        recs.append(x)
    res = engine.psk_demod_batch(recs, d)
    for x, r in zip(recs, res):
        if len(x) <= 27:
            assert r.status == _lib.FB_ST_TOO_SHORT and r.raw == b""
            continue
        want = o2.qpsk_stages(x, 9600, 9600.0)
        assert r.raw == want["raw"], len(x)
        assert r.sync_idx == want["sync"]
    # float64 and PCM16 storage of the same (int16-exact) samples give the same bytes as the oracle on them
    q = np.clip(np.round(recs[0] * 32767), -32768, 32767).astype(np.int16)
    xf = q.astype(np.float32) / np.float32(32768.0)
    want = o2.qpsk_demodulate(xf, 9600, 9600.0)
    for arr in (xf, xf.astype(np.float64), q):
        assert engine.psk_demod_batch([arr], d)[0].raw == want


def test_shard_invariance_of_batching(engine):
    """Any grouping / order of the same recordings gives identical per-recording bytes."""
    import fbdsp
    d = fbdsp.psk_design(9600.0, 3000.0, 96000.0, 1.5, False)
    recs = [sig.kat_signal(sig.qpsk_modulate, 300 + i, 500 + 300 * i, 20, baud=9600, carrier=3000.0)[2] for i in range(6)]
    whole = [r.raw for r in engine.psk_demod_batch(recs, d)]
    rev = [r.raw for r in engine.psk_demod_batch(recs[::-1], d)][::-1]
    single = [engine.psk_demod_batch([x], d)[0].raw for x in recs]
    assert whole == rev == single


@pytest.mark.parametrize("dt", [np.float32, np.int16])
def test_uniform_batch_descriptors(dt, engine, monkeypatch):
    """Equal-length recordings back to back take the arithmetic tile descriptors (no descriptor table read); the same
    batch through the table (FB_PSK_NO_UNIFORM=1) and one recording at a time must give identical bytes."""
    import fbdsp
    rng = np.random.default_rng(31)
    n = 96000 * 2 + 13
    recs = []
    for k in range(5):
        _, _, x = sig.kat_signal(sig.qpsk_modulate, 300 + k, 4000, 15, baud=9600, carrier=9600.0)
        x = np.concatenate([x, (rng.standard_normal(n) * 0.01).astype(np.float32)])[:n]
        recs.append((x * 20000).astype(np.int16) if dt == np.int16 else x)
    d = fbdsp.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
    monkeypatch.delenv("FB_PSK_NO_UNIFORM", raising=False)
    uni = engine.psk_demod_batch(recs, d)
    monkeypatch.setenv("FB_PSK_NO_UNIFORM", "1")
    tab = engine.psk_demod_batch(recs, d)
    for r, (a, b) in enumerate(zip(uni, tab)):
        one = engine.psk_demod_batch([recs[r]], d)[0]
        assert a.raw == b.raw == one.raw and a.sync_idx == b.sync_idx == one.sync_idx and a.status == b.status == 0
        assert len(a.raw) > 4000


def test_sync_beyond_the_search_head(engine):
    """The magic search runs in two launches (the first 131 072 bits, then the rest).  An unmodulated carrier decides
    '00' for 8.5 s, so the first "FB" lies beyond the head; a second recording syncs at once and a third never does."""
    import fbdsp
    from oracle.frames import frame_data
    rng = np.random.default_rng(77)
    t = np.arange(int(8.5 * 96000)) / 96000
    lead = np.sin(2 * np.pi * 9600.0 * t)
    payload = rng.integers(0, 256, 3000, dtype=np.uint8).tobytes()
    body = sig.qpsk_modulate(frame_data("late.bin", payload, 0, 1, 3000, 0), 9600, 9600.0).astype(np.float64)
    late = sig.add_awgn(np.concatenate([lead, body]).astype(np.float32), 25, rng)
    early = sig.add_awgn(body.astype(np.float32), 25, rng)
    never = sig.add_awgn(lead.astype(np.float32), 25, rng)
    d = fbdsp.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
    res = engine.psk_demod_batch([late, early, never], d)
    for x, r in zip((late, early, never), res):
        st = o2.qpsk_stages(x, 9600, 9600.0)
        assert r.sync_idx == st["sync"] and r.raw == st["raw"]
    assert res[0].sync_idx > 131072 + 64 and 0 <= res[1].sync_idx < 200 and res[2].sync_idx == -1


def test_pipelined_host_copy_equals_single_copy(engine, monkeypatch):
    """Host batches of >= 256 MB are copied in up to 16 groups on a copy stream while earlier groups are demodulated
    (csrc/psk_v2.cu); FB_PSK_PIPE_MB lowers the threshold so that a small ragged batch takes that path: the per-recording
    bytes, sync indices and statuses must equal the single-copy path's."""
    import fbdsp
    rng = np.random.default_rng(77)
    d = fbdsp.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
    recs = []
    for i in range(19):
        _, _, x = sig.kat_signal(sig.qpsk_modulate, 7700 + i, int(rng.integers(200, 1500)), 20, baud=9600, carrier=9600.0)
        recs.append(np.concatenate([0.01 * rng.standard_normal(int(rng.integers(0, 3000))).astype(np.float32), x]))
    recs.append(np.zeros(5, np.float32))                        # too short: status, not an exception
    monkeypatch.delenv("FB_PSK_PIPE_MB", raising=False)
    want = engine.psk_demod_batch(recs, d, exact_silence=False)
    monkeypatch.setenv("FB_PSK_PIPE_MB", "0")
    got = engine.psk_demod_batch(recs, d, exact_silence=False)
    assert [(r.raw, r.sync_idx, r.status) for r in got] == [(r.raw, r.sync_idx, r.status) for r in want]
    assert any(len(r.raw) > 100 for r in got)


def test_tiles_outside_the_fp16_split_take_the_fp32_kernel(engine, monkeypatch):
    """psk_mma_kernel hands tiles whose samples do not fit its fp16 hi/lo split (|x| >= 4, or a whole tile below 2^-18) back
    to the fp32 kernel (redo list).  One recording with a normal, a loud (x 6) and a very quiet (x 1e-6) stretch of several
    tiles each: the decisions follow the oracle under the same rule as everywhere (mismatches only at margin < 1e-5 or on
    symbols > 1e-7 below the record's peak), with the tensor-pipe kernel and with the fp32 kernel alone."""
    import fbdsp
    rng = np.random.default_rng(4242)
    parts = []
    for g in (0.3, 6.0, 1e-6, 0.3):
        x = sig.qpsk_modulate(bytes(rng.integers(0, 256, 2400, dtype=np.uint8)), baud=9600, carrier=9600.0)
        parts.append((g * (x + 0.05 * rng.standard_normal(len(x)))).astype(np.float32))
    x = np.concatenate(parts)
    assert len(x) > 16 * 18880 and np.abs(x).max() > 4.0
    d = fbdsp.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
    st = o2.qpsk_stages(x, 9600, 9600.0)
    mag = np.abs(st["diff"])
    margin = o2.qpsk_margin(st["diff"])
    for no_mma in (False, True):
        if no_mma:
            monkeypatch.setenv("FB_PSK_NO_MMA", "1")
        else:
            monkeypatch.delenv("FB_PSK_NO_MMA", raising=False)
        engine.psk_demod_batch([x], d, exact_silence=False)
        got = engine.last_bits(0)
        assert len(got) == len(st["bits"])
        bad = np.unique(np.nonzero(got != st["bits"])[0] // 2)
        assert len(bad) < 1e-4 * len(mag)
        for k in bad:
            assert margin[k] < 1e-5 or mag[k] < 1e-14 * mag.max(), (no_mma, k, margin[k], mag[k] / mag.max())


@pytest.mark.parametrize("dt", [np.float32, np.int16, np.float64])
def test_every_buffer_alignment_of_a_recording(dt, engine):
    """Recordings that start at every element offset mod 8 of the batch buffer (the tensor-pipe kernel keeps one set of B
    fragments per alignment and stages through TMA only float32 tiles whose window start is a multiple of 8 elements), several
    tiles each, all three storage types: raw bytes and sync index equal the oracle's on the same (int16-exact) samples."""
    import fbdsp
    d = fbdsp.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
    recs = []
    for i in range(9):
        _, _, x = sig.kat_signal(sig.qpsk_modulate, 600 + i, 1300, 20, baud=9600, carrier=9600.0)
        q = np.clip(np.round(x[: 60001 + i] * 32767), -32768, 32767).astype(np.int16)      # lengths 60001 .. 60009: offsets walk mod 8
        recs.append(q)
    xs = [q.astype(np.float32) / np.float32(32768.0) for q in recs]
    arrs = recs if dt == np.int16 else [x.astype(dt) for x in xs]
    res = engine.psk_demod_batch(arrs, d, exact_silence=False)
    for x, r in zip(xs, res):
        want = o2.qpsk_stages(x, 9600, 9600.0)
        assert r.raw == want["raw"] and r.sync_idx == want["sync"]
