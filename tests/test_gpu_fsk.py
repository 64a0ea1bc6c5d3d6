"""-m gpu: v2 FSK device path against the golden fixtures (reference outputs) and the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import modem_v2 as o2, signals as sig

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FSK_FNS = ("fsk_demodulate", "fsk_high_speed_demodulate", "ft8_demodulate")


def _cases():
    cs = json.load(open(os.path.join(GOLD, "demod_cases.json")))["cases"]
    return [c for c in cs if c["fn"] in FSK_FNS]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_golden_bytes(case, golden, engine):
    from fbdsp import modem
    _, arrs = golden
    x = arrs[case["name"] + ".x"]
    fn = getattr(modem, case["fn"])
    if case["exc"]:
        with pytest.raises(Exception) as ei:
            fn(x, *case["args"])
        assert type(ei.value).__name__ == case["exc"][0] and str(ei.value) == case["exc"][1]
        return
    assert fn(x, *case["args"]) == arrs[case["name"] + ".raw"].tobytes()


@pytest.mark.parametrize("baud,mark,space,nbytes,snr,seed", [(9600, 12000.0, 24000.0, 8192, 20, 601), (9600, 14400.0, 33600.0, 3000, 10, 602),
                                                              (4800, 8000.0, 16000.0, 2000, 20, 603), (600, 1200.0, 2200.0, 100, 20, 604),
                                                              (19200, 24000.0, 26000.0, 2000, 15, 605)])
def test_fsk_decisions_vs_oracle(baud, mark, space, nbytes, snr, seed, engine):
    """Config-1 working variant (9600 Bd, 12k/24k) and friends: decided bits equal the oracle's except where the
    envelope margin |mark-space|/max is below 1e-5 somewhere in the vote window."""
    from fbdsp import fsk
    mb = min(baud, 9600)
    _, _, x = sig.kat_signal(sig.fsk_modulate, seed, nbytes, snr, baud=mb, mark_freq=mark, space_freq=space)
    if len(x) % 2 == 0:
        x = x[:-1]                                          # odd N exercises the other hilbert mask branch
    d = fsk.fsk_design(baud, mark, space, 96000, len(x))
    res = fsk.fsk_demod_batch([x], d, engine)[0]
    st = o2.fsk_stages(x, baud, mark, space)
    got = engine.last_bits(0)
    assert len(got) == len(st["bits"])
    bad = np.nonzero(got != st["bits"])[0]
    spb = int(96000 / baud)
    margin = np.abs(st["mark_env"] - st["space_env"]) / np.maximum(st["mark_env"], st["space_env"])
    for b in bad:
        c = spb // 2 + b * spb
        assert margin[c - spb // 4: c + spb // 4].min() < 1e-5
    assert len(bad) <= 1e-4 * len(got)
    if len(bad) == 0:
        assert res.raw == st["raw"] and res.sync_idx == st["sync"]


def test_fsk_roundtrip_64k_config1_variant(engine):
    """BASELINE configs[0] with tones that design: 64 KiB file comes back bit-exact."""
    from fbdsp import modem, frames
    from oracle.frames import frame_data
    rng = np.random.default_rng(7)
    payload = rng.integers(0, 256, 65536, dtype=np.uint8).tobytes()
    x = sig.fsk_modulate(frame_data("f64k.bin", payload, 0, 1, 65536, 0), 9600, 12000.0, 24000.0)
    raw = modem.fsk_demodulate(x, 9600, 12000.0, 24000.0)
    fr = frames.parse_batch([raw], engine)[0]
    assert len(fr) == 1 and fr[0]["data"] == payload
