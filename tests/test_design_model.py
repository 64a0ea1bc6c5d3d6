"""Host logic: the product's filter design (fbdsp.design), evaluated by a numpy model of the kernels
(tests/model_psk.py), must reproduce the oracle's decided bits exactly.  CPU only."""
import json
import os

import numpy as np
import pytest

import model_psk as M
from fbdsp import design
from oracle import modem_v2 as o2

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PSK_FNS = {"qpsk_demodulate": (1.5, False), "psk8_demodulate": (1.5, False), "ofdm_demodulate_simple": (1.5, False),
           "bpsk_demodulate": (1.0, True)}


def _cases():
    cs = json.load(open(os.path.join(GOLD, "demod_cases.json")))["cases"]
    return [c for c in cs if c["fn"] in PSK_FNS and not c["exc"] and c["n"] > 27]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_model_bits_equal_oracle(case, golden):
    _, arrs = golden
    x = arrs[case["name"] + ".x"]
    band_k, n0_is_sps = PSK_FNS[case["fn"]]
    baud, carrier = case["args"][0], case["args"][1]
    d = design.psk_design(float(baud), float(carrier), 96000.0, band_k, n0_is_sps)
    st = (o2.bpsk_stages if n0_is_sps else o2.qpsk_stages)(x, baud, carrier)
    want = st["bits"] if st["bits"] is not None else np.zeros(0, np.uint8)
    got = M.demod_bits(x, d)
    assert len(got) == len(want)
    assert np.array_equal(got, want)


def test_design_raises_like_scipy():
    with pytest.raises(ValueError, match="Digital filter critical frequencies must be 0 < Wn < 1"):
        design.psk_design(48000.0, 3000.0, 96000.0, 1.5, False)


def test_design_geometry():
    d = design.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
    assert d.sps == 10 and d.n0 == 5 and d.bits_per_sym == 2 and not d.emulate_only
    assert d.taps.shape == (10, d.nt) and d.nt == d.dl + d.dh + 1
    assert d.diag["leak"] <= design.FIR_TOL
    d = design.psk_design(31.25, 1000.0, 96000.0, 1.0, True)       # psk31: far too narrow for the FIR table
    assert d.emulate_only
