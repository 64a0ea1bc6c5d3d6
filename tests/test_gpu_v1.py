"""-m gpu: the v1 streaming kernels against oracle/modem_v1.py (SURVEY Appendix B restatement; PARITY UNPINNED --
no executable reference exists for v1, so these tests pin the kernels to the restatement only)."""
import numpy as np
import pytest

from oracle import modem_v1 as v1

pytestmark = pytest.mark.gpu


def _bits(raw):
    return np.unpackbits(np.frombuffer(raw, dtype=np.uint8))


def _check(got_raw, st, margin_per_sym, bps):
    want_bits = st["bits"][: (len(st["bits"]) // 8) * 8]
    got_bits = _bits(got_raw)
    assert len(got_bits) == len(want_bits)
    bad = np.nonzero(got_bits != want_bits)[0]
    assert len(bad) <= 1e-4 * max(1, len(want_bits))
    for b in bad:
        assert margin_per_sym[b // bps] < 1e-5
    if len(bad) == 0:
        assert got_raw == st["raw"]


def _psk_signal(bps, baud, carrier, nsym, snr_db, seed):
    """cos-referenced PSK at sector centres + AWGN (the v1 TX itself is sin-referenced and never round-tripped)."""
    rng = np.random.default_rng(seed)
    sps = int(round(96000 / baud))
    t = np.arange(sps) / 96000
    m = 2 ** bps
    ph = rng.integers(0, m, nsym) * (2 * np.pi / m) + (np.pi / m if bps > 1 else 0.0)
    x = np.cos(2 * np.pi * carrier * t[None, :] - ph[:, None]).reshape(-1)
    x = x + rng.standard_normal(len(x)) * np.sqrt(0.5 / 10 ** (snr_db / 10))
    return np.concatenate([x, rng.standard_normal(7) * 0.1]).astype(np.float32)      # ragged tail


@pytest.mark.parametrize("baud,carrier,snr", [(9600, 9600.0, 10), (1200, 3000.0, 5), (38400, 12000.0, 10), (2400, 12000.0, 0)])
def test_v1_qpsk(baud, carrier, snr, engine):
    from fbdsp import modem_v1 as g
    x = _psk_signal(2, baud, carrier, 40000 if baud >= 9600 else 4000, snr, baud)
    st = v1.qpsk_stages(x, baud, carrier)
    mag = np.hypot(st["i"], st["q"])
    margin = np.minimum(np.abs(st["i"]), np.abs(st["q"])) / np.where(mag > 0, mag, 1)
    _check(g.qpsk_demodulate(x, baud, carrier), st, margin, 2)


def test_v1_bpsk_and_8psk(engine):
    from fbdsp import modem_v1 as g
    x = _psk_signal(1, 9600, 3000.0, 30000, 10, 1)
    st = v1.bpsk_stages(x, 9600, 3000.0)
    _check(g.bpsk_demodulate(x, 9600, 3000.0), st, np.abs(st["i"]) / np.maximum(np.hypot(st["i"], st["q"]), 1e-300), 1)
    x = _psk_signal(3, 2400, 12000.0, 9000, 15, 2)
    st = v1.psk8_stages(x, 2400, 12000.0)
    thr = np.array([1, 3, 5, 7, 9, 11, 13, 16]) * np.pi / 8        # sector edges (0/2pi is not an edge: '111' wraps)
    margin = np.min(np.abs(st["phi"][:, None] - thr[None, :]), axis=1) / (np.pi / 8)
    _check(g.psk8_demodulate(x, 2400, 12000.0), st, margin, 3)


@pytest.mark.parametrize("baud,nsub", [(9600, 8), (9600, 4), (4800, 8), (2400, 4)])
def test_v1_ofdm(baud, nsub, engine):
    from fbdsp import modem_v1 as g
    rng = np.random.default_rng(baud + nsub)
    x = rng.standard_normal(300000 + 3)                           # float64 stays float64 (B.7 has no cast)
    st = v1.ofdm_stages(x, baud, 12000.0, nsub)
    sc = st["sc"].reshape(-1)
    margin = np.minimum(np.abs(sc.real), np.abs(sc.imag)) / np.abs(sc)
    _check(g.ofdm_demodulate_simple(x, baud, 12000.0, nsub), st, margin, 2)
    assert g.ofdm_demodulate_simple(x.astype(np.float32), baud, 12000.0, nsub) == v1.ofdm_demodulate_simple(x.astype(np.float32), baud, 12000.0, nsub)


def test_v1_fsk_uart_roundtrip_and_hs(engine):
    from fbdsp import modem_v1 as g
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, 700, dtype=np.uint8).tobytes()
    x = v1.fsk_modulate(data, 1200)
    x = (x + rng.standard_normal(len(x)).astype(np.float32) * np.float32(0.05)).astype(np.float32)
    st = v1.fsk_stages(x, 1200)
    got = g.fsk_demodulate(x, 1200)
    assert got == st["raw"] == data                                # v1 FSK1200 does round-trip (UART framing)
    x = (rng.standard_normal(200001) * 0.3).astype(np.float32)     # FSK-HS on noise: decisions vs the restatement
    st = v1.fsk_high_speed_stages(x)
    margin = np.abs(st["p_mark"] - st["p_space"]) / np.maximum(st["p_mark"], st["p_space"])
    _check(g.fsk_high_speed_demodulate(x), st, margin, 1)
    with pytest.raises(ValueError):
        g.fsk_demodulate(np.zeros(20, np.float32), 1200)


def test_v1_ragged_batch(engine):
    from fbdsp import modem_v1 as g
    p, table = g.psk_params(g.V1_QPSK, 9600, 9600.0)
    recs = [_psk_signal(2, 9600, 9600.0, n, 12, 50 + n) for n in (1, 7, 256, 257, 5000, 0)]
    recs[-1] = np.zeros(3, np.float32)                             # fewer samples than one symbol -> b''
    out = g.demod_batch(recs, p, table, engine)
    for x, (raw, st) in zip(recs, out):
        assert raw == v1.qpsk_demodulate(x, 9600, 9600.0)


def _uart_gold(bits):
    """B.2 deframer restated once more, straight from the pyc listing (independent of oracle.modem_v1.uart_deframe)."""
    out, i, n = bytearray(), 0, len(bits)
    while i + 10 <= n:
        if bits[i] != 0 or bits[i + 9] != 1:
            i += 1
            continue
        out.append(sum(int(bits[i + 1 + k]) << k for k in range(8)))
        i += 10
    return bytes(out)


@pytest.mark.parametrize("baud,n", [(1200, 30), (1200, 79), (1200, 4000), (1200, 250007), (9600, 28), (9600, 10241), (9600, 1200003),
                                    (4800, 333333)])
def test_v1_fsk_noise_lengths(baud, n, engine):
    """Goertzel FSK on noise at awkward lengths: chunk-parallel pre-filter (short records, ragged last chunk, record
    shorter than one chunk) and the chunk-parallel UART deframer (walks that end inside a chunk, n < 10 bits) against
    the restatement; bits may differ only where the tone powers are within 1e-5 of each other."""
    from fbdsp import modem_v1 as g
    rng = np.random.default_rng(baud + n)
    x = (rng.standard_normal(n) * 0.4).astype(np.float32)
    st = v1.fsk_stages(x, baud)
    got = g.fsk_demodulate(x, baud)
    assert st["raw"] == _uart_gold(st["bits"])
    if got != st["raw"]:
        # a margin flip changes the framing downstream: accept only if some near-tie bit exists, then compare through it
        margin = np.abs(st["p_mark"] - st["p_space"]) / np.maximum(st["p_mark"], st["p_space"])
        assert np.min(margin) < 1e-5, "UART bytes differ from the restatement without any near-tie decision"


def test_v1_uart_deframer_dense(engine):
    """A clean UART stream (every byte framed) plus garbage in front: exercises next(i) = i + 10 runs across many chunks
    and the byte offsets handed from chunk to chunk."""
    from fbdsp import modem_v1 as g
    rng = np.random.default_rng(77)
    data = rng.integers(0, 256, 6000, dtype=np.uint8).tobytes()
    x = v1.fsk_modulate(data, 9600)
    x = np.concatenate([(rng.standard_normal(777) * 0.2).astype(np.float32), x.astype(np.float32)])
    st = v1.fsk_stages(x, 9600)
    assert g.fsk_demodulate(x, 9600) == st["raw"] == _uart_gold(st["bits"])
    assert len(st["raw"]) > 3000                                   # mostly framed bytes: long i += 10 runs


def _sym_cases():
    from fbdsp import modem_v1 as g
    return {
        "bpsk9600": lambda: g.psk_params(g.V1_BPSK, 9600, 3000.0), "qpsk9600": lambda: g.psk_params(g.V1_QPSK, 9600, 9600.0),
        "psk8_9600": lambda: g.psk_params(g.V1_PSK8, 9600, 12000.0), "bpsk38400": lambda: g.psk_params(g.V1_BPSK, 38400, 12000.0),
        "qpsk38400": lambda: g.psk_params(g.V1_QPSK, 38400, 12000.0), "psk8_38400": lambda: g.psk_params(g.V1_PSK8, 38400, 12000.0),
        "bpsk4800": lambda: g.psk_params(g.V1_BPSK, 4800, 3000.0), "qpsk4800": lambda: g.psk_params(g.V1_QPSK, 4800, 9600.0),
        "psk8_4800": lambda: g.psk_params(g.V1_PSK8, 4800, 12000.0), "ofdm8_9600": lambda: g.ofdm_params(9600, 8),
        "ofdm4_4800": lambda: g.ofdm_params(4800, 4), "fsk9600": lambda: g.fsk_params(9600, 8000.0, 16000.0, 7500, 16500, True),
        "fskhs19200": lambda: g.fsk_params(19200, 12000.0, 18000.0, 8000, 22000, False),
        "fsk4800": lambda: g.fsk_params(4800, 8000.0, 16000.0, 7500, 16500, True),
        "bpsk2400": lambda: g.psk_params(g.V1_BPSK, 2400, 3000.0), "qpsk2400": lambda: g.psk_params(g.V1_QPSK, 2400, 3000.0),
        "psk8_2400": lambda: g.psk_params(g.V1_PSK8, 2400, 12000.0),
        "bpsk1200": lambda: g.psk_params(g.V1_BPSK, 1200, 3000.0), "qpsk1200": lambda: g.psk_params(g.V1_QPSK, 1200, 3000.0),
        "psk8_1200": lambda: g.psk_params(g.V1_PSK8, 1200, 12000.0),
    }


@pytest.mark.parametrize("name", ["bpsk9600", "qpsk9600", "psk8_9600", "bpsk38400", "qpsk38400", "psk8_38400", "bpsk4800", "qpsk4800",
                                  "psk8_4800", "ofdm8_9600", "ofdm4_4800", "fsk9600", "fskhs19200", "fsk4800", "bpsk2400", "qpsk2400", "psk8_2400", "bpsk1200", "qpsk1200", "psk8_1200"])
def test_v1_sym_kernel_equals_generic(name, engine, monkeypatch):
    """Every compile-time geometry of v1_sym_kernel against the generic v1_corr_kernel (itself checked against the
    restatement above) on a ragged batch whose recordings start at every residue mod 4 (all register-shift phases of
    the vector symbol loads), span several tiles, end inside a tile, or hold less than one symbol: byte-exact -- both
    kernels evaluate the same FMA chains."""
    from fbdsp import modem_v1 as g
    p, table = _sym_cases()[name]()
    rng = np.random.default_rng(len(name) * 1000 + p.sps)
    lens = [30001, 30, 40002, 1, 61443, 8191, 29, 123457, 31, 50000]
    recs = [(rng.standard_normal(n) * 0.4).astype(np.float32) for n in lens]
    monkeypatch.delenv("FB_V1_GENERIC", raising=False)
    fast = g.demod_batch(recs, p, table, engine)
    monkeypatch.setenv("FB_V1_GENERIC", "1")
    slow = g.demod_batch(recs, p, table, engine)
    assert [s for _, s in fast] == [s for _, s in slow]
    for (a, _), (b, _), n in zip(fast, slow, lens):
        assert a == b, f"{name}: recording of {n} samples differs"
    assert sum(len(a) for a, _ in fast) > 300


def test_v1_sym_kernel_oracle_multi(engine):
    """The specialised kernels directly against the restatement on a ragged batch (QPSK-9600, 8PSK-38400, OFDM8)."""
    from fbdsp import modem_v1 as g
    rng = np.random.default_rng(99)
    lens = [20001, 3, 40962, 12345, 77777]
    for mode, baud, carrier, bps, stages in ((g.V1_QPSK, 9600, 9600.0, 2, v1.qpsk_stages), (g.V1_PSK8, 38400, 12000.0, 3, v1.psk8_stages)):
        recs = [_psk_signal(bps, baud, carrier, max(1, n // int(round(96000 / baud))), 12, n) for n in lens]
        out = g.demod_batch(recs, *g.psk_params(mode, baud, carrier), engine)
        for x, (raw, _) in zip(recs, out):
            st = stages(x, baud, carrier)
            if mode == g.V1_QPSK:
                mag = np.hypot(st["i"], st["q"])
                margin = np.minimum(np.abs(st["i"]), np.abs(st["q"])) / np.where(mag > 0, mag, 1)
            else:
                thr = np.array([1, 3, 5, 7, 9, 11, 13, 16]) * np.pi / 8
                margin = np.min(np.abs(st["phi"][:, None] - thr[None, :]), axis=1) / (np.pi / 8)
            _check(raw, st, margin, bps)
    recs = [rng.standard_normal(n).astype(np.float32) for n in lens]
    out = g.demod_batch(recs, *g.ofdm_params(9600, 8), engine)
    for x, (raw, _) in zip(recs, out):
        st = v1.ofdm_stages(x, 9600, 12000.0, 8)
        sc = st["sc"].reshape(-1)
        margin = np.minimum(np.abs(sc.real), np.abs(sc.imag)) / np.maximum(np.abs(sc), 1e-300)
        _check(raw, st, margin, 2)


def test_v1_sym_kernel_specials(engine):
    """Digital silence, exact sector centres / edges and non-finite samples through the sign-bit slicers."""
    from fbdsp import modem_v1 as g
    z = np.zeros(4000, np.float32)
    for mode, baud, carrier, f in ((g.V1_QPSK, 9600, 9600.0, v1.qpsk_demodulate), (g.V1_BPSK, 9600, 3000.0, v1.bpsk_demodulate),
                                   (g.V1_PSK8, 38400, 12000.0, v1.psk8_demodulate), (g.V1_PSK8, 9600, 12000.0, v1.psk8_demodulate)):
        raw, _ = g.demod_batch([z], *g.psk_params(mode, baud, carrier), engine)[0]
        assert raw == f(z, baud, carrier)
    raw, _ = g.demod_batch([z], *g.ofdm_params(9600, 8), engine)[0]
    assert raw == v1.ofdm_demodulate_simple(z, 9600, 12000.0, 8)
    # noiseless 8PSK steered (2x2 solve for the non-orthogonal references) to sector centres, to 3e-5 and 2e-6 rad either
    # side of every edge (must match exactly: the second is inside the float32 pre-slicer's guard band at sps 2, so it
    # exercises the float64 fall-back) and onto the edges themselves (a 1-ulp atan2 difference may flip those: margin rule)
    for baud in (9600, 38400):
        sps = int(round(96000 / baud))
        t = np.arange(sps) / 96000
        rc, rs = np.cos(2 * np.pi * 12000.0 * t), np.sin(2 * np.pi * 12000.0 * t)
        gram = np.array([[rc @ rc, rc @ rs], [rc @ rs, rs @ rs]])
        edges = np.arange(1, 16, 2) * np.pi / 8
        for delta, exact in ((3e-5, True), (-3e-5, True), (2e-6, True), (-2e-6, True), (0.0, False)):
            th = np.concatenate([edges + delta, np.arange(8) * np.pi / 4])
            ab = np.linalg.solve(gram, np.stack([np.cos(th), np.sin(th)]))
            x = (ab[0][:, None] * rc[None, :] + ab[1][:, None] * rs[None, :]).reshape(-1)
            x = np.tile(x, 24).astype(np.float32)
            raw, _ = g.demod_batch([x], *g.psk_params(g.V1_PSK8, baud, 12000.0), engine)[0]
            st = v1.psk8_stages(x, baud, 12000.0)
            thr = np.array([1, 3, 5, 7, 9, 11, 13, 16]) * np.pi / 8
            margin = np.min(np.abs(st["phi"][:, None] - thr[None, :]), axis=1) / (np.pi / 8)
            if exact:
                assert margin.min() > 1e-6 and raw == st["raw"], (baud, delta)
            else:
                want, got = st["bits"][: len(st["bits"]) // 8 * 8], _bits(raw)
                assert len(want) == len(got)
                for b_ in np.nonzero(want != got)[0]:
                    assert margin[b_ // 3] < 1e-5
    # NaN samples: the QPSK comparison chain of B.5 falls through to '10'
    y = (np.random.default_rng(3).standard_normal(4000) * 0.3).astype(np.float32)
    y[105] = np.nan
    raw, _ = g.demod_batch([y], *g.psk_params(g.V1_QPSK, 9600, 9600.0), engine)[0]
    assert raw == v1.qpsk_demodulate(y, 9600, 9600.0)


def test_v1_ofdm_prepass_fallbacks(engine, monkeypatch):
    """The float32 pre-pass of the OFDM demap must hand every symbol it cannot certify to the float64 path: constant
    input (all demapped bins exactly zero), tiny bins next to large ones, a NaN sample -- same bytes as the generic
    float64 kernel (a direct DFT leaves 1e-17 residues where the restatement's FFT has exact zeros, so the restatement is
    not the yardstick for these inputs)."""
    from fbdsp import modem_v1 as g
    rng = np.random.default_rng(12)
    n = 40000
    const = np.full(n, 0.25, np.float32)
    mixed = (rng.standard_normal(n) * 0.3).astype(np.float32)
    mixed[1000:3000] = 0.5                                          # a run of symbols whose bins are exactly zero
    mixed[5000:7000] += np.float32(1e-7) * rng.standard_normal(2000).astype(np.float32)
    tone = np.cos(2 * np.pi * 12000.0 * np.arange(n) / 96000).astype(np.float32)   # energy in one bin only, others ~1e-8
    nan = mixed.copy()
    nan[12345] = np.nan
    for baud, nsub in ((9600, 8), (4800, 4)):
        p, table = g.ofdm_params(baud, nsub)
        recs = [const, mixed, tone, nan]
        monkeypatch.delenv("FB_V1_GENERIC", raising=False)
        fast = g.demod_batch(recs, p, table, engine)
        monkeypatch.setenv("FB_V1_GENERIC", "1")
        slow = g.demod_batch(recs, p, table, engine)
        for (a, _), (b, _) in zip(fast, slow):
            assert a == b
        monkeypatch.delenv("FB_V1_GENERIC", raising=False)
