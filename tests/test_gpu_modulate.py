"""-m gpu: batch modulators against waveforms recorded from the UNMODIFIED reference modulators
(tests/golden/modulators.npz, tools/make_golden_mod.py) and against oracle/signals.py on larger seeded payloads.
Tolerance (float path): samples are float32(sin(float64)); CUDA's and numpy's float64 sin may differ in the last place,
which can flip the float32 rounding of a sample: |diff| <= 1 float32 ulp of 1.0 (1.2e-7), on at most 1e-5 of the samples."""
import json
import os

import numpy as np
import pytest

from oracle import signals as sig

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _close(got, want):
    assert got.dtype == np.float32 and got.shape == want.shape
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    assert d.max(initial=0.0) <= 1.2e-7
    assert np.count_nonzero(d) <= max(1, 1e-5 * len(want))


def test_modulators_golden(engine):
    from fbdsp import modem as fb
    z = np.load(os.path.join(GOLD, "modulators.npz"))
    meta = json.load(open(os.path.join(GOLD, "modulators.json")))
    for c in meta["cases"]:
        data = z[c["name"] + "__data"].tobytes()
        got = getattr(fb, c["fn"])(data, *c["args"])
        _close(got, z[c["name"] + "__y"])
    for e in meta["errors"]:
        with pytest.raises(ValueError) as ei:
            getattr(fb, e["fn"])(b"abc", *e["args"])
        assert str(ei.value) == e["msg"]
    assert len(fb.qpsk_modulate(b"abc", 200000)) == 0            # sps == 0: empty waveform, as in the reference
    with pytest.raises(ZeroDivisionError):
        fb.qpsk_modulate(b"abc", 0)


@pytest.mark.parametrize("kind", ["qpsk", "bpsk", "fsk"])
def test_modulate_batch_vs_oracle(kind, engine):
    """Ragged batch (empty payload included) at the BASELINE parameter sets; phases reach ~1e5 rad."""
    from fbdsp import modulate as m
    rng = np.random.default_rng({"qpsk": 1, "bpsk": 2, "fsk": 3}[kind])
    payloads = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (30000, 0, 1, 4097, 12345)]
    if kind == "qpsk":
        prm, ref = m.psk_mod_params(m.FB_MOD_DQPSK, 9600, 9600.0, 96000), lambda b: sig.qpsk_modulate(b, 9600, 9600.0)
    elif kind == "bpsk":
        prm, ref = m.psk_mod_params(m.FB_MOD_DBPSK, 4800, 9600.0, 96000), lambda b: sig.bpsk_modulate(b, 4800, 9600.0)
    else:
        prm, ref = m.fsk_mod_params(9600, 12000.0, 24000.0, 96000), lambda b: sig.fsk_modulate(b, 9600, 12000.0, 24000.0)
    out = m.modulate_batch(payloads, *prm, engine)
    for b, y in zip(payloads, out):
        _close(y, ref(b))


def test_modulate_demodulate_on_device(engine):
    """TX and RX both on the device: frame -> DQPSK waveform (device) -> demodulate -> parse -> payload."""
    from fbdsp import modem as fb
    from oracle import frames as ofr
    from fbdsp import frames as fr
    rng = np.random.default_rng(8)
    payload = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()
    framed = ofr.frame_data("tx.bin", payload, 0, 1, len(payload), 0)
    x = fb.qpsk_modulate(framed, 9600, 9600.0)
    raw = fb.qpsk_demodulate(x, 9600, 9600.0)
    got = fr.parse_fbp_stream_enhanced(raw)
    assert len(got) == 1 and got[0]["data"] == payload and got[0]["name"] == "tx.bin"
    x = fb.fsk_modulate(framed, 9600, 12000.0, 24000.0)
    got = fr.parse_fbp_stream_enhanced(fb.fsk_demodulate(x, 9600, 12000.0, 24000.0))
    assert len(got) == 1 and got[0]["data"] == payload
