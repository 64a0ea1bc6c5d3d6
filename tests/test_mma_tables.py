"""CPU: the tensor-core DPSK formulation (fbdsp/mma_tables.py, csrc/psk_mma.cu).
  * the numpy model of the kernel (tests/model_psk_mma.py: fp16 hi/lo operands, fp32 accumulation with truncation, group-level
    slow-pole states) reproduces the oracle's decisions on interior symbols;
  * the C++ table builder inside libfbdsp.so (host code, no GPU needed) produces the band matrix of the Python statement."""
import ctypes

import numpy as np
import pytest

import fbdsp
from fbdsp import _lib, mma_tables as mt
from oracle import modem_v2 as o2, signals as sig

import model_psk
import model_psk_mma


@pytest.mark.parametrize("carrier", [9600.0, 3000.0])
def test_model_matches_oracle_decisions(carrier):
    d = fbdsp.psk_design(9600.0, carrier, 96000.0, 1.5, False)
    t = mt.tables(d)
    assert t["ksteps"] == 15 and t["hh_blocks"].sum() <= 26 and t["lo_blocks"].sum() <= 18
    _, _, x = sig.kat_signal(sig.qpsk_modulate, 21, 1500, 20, baud=9600, carrier=carrier)
    k_lo, k_hi = 1024, 1024 + 512 - 1
    u, _ = model_psk_mma.interior_mma(x, 3, d, k_lo, k_hi)
    y = u / (t["sx"] * t["st"])
    yref = model_psk.interior(x, d, k_lo, k_hi)
    assert np.max(np.abs(y - yref)) < 2e-6 * np.max(np.abs(yref))
    st = o2.qpsk_stages(x, 9600, carrier)
    dd = y[1:] * np.conj(y[:-1]) * complex(d.c_struct.rho[0], d.c_struct.rho[1])
    assert np.array_equal(o2.qpsk_slice(dd), st["bits"][2 * k_lo: 2 * k_hi])


def test_cpp_band_matrix_equals_python():
    lib = _lib.load()
    for carrier in (9600.0, 3000.0, 19200.0):
        d = fbdsp.psk_design(9600.0, carrier, 96000.0, 1.5, False)
        out = np.zeros((8, 240, 16), dtype=np.float64)
        ks = lib.fb_debug_mma_band(ctypes.byref(d.c_struct), d.taps.ctypes.data, out.ctypes.data)
        if ks < 0:
            continue
        assert ks == 15
        scale = max(np.max(np.abs(mt.band_matrices(d, 0)[0])), 1e-30)
        for sh in range(8):
            want = mt.band_matrices(d, sh)[0]
            # the C++ side starts from the float32 taps / residues the ABI carries: 24-bit inputs
            assert np.max(np.abs(out[sh] - want)) < 3e-7 * scale, (carrier, sh)
    d = fbdsp.psk_design(4800.0, 9600.0, 96000.0, 1.0, True)              # sps 20: not this kernel's class
    assert lib.fb_debug_mma_band(ctypes.byref(d.c_struct), d.taps.ctypes.data, np.zeros((8, 240, 16)).ctypes.data) == -1
