"""libfbdsp.so loads and exports every symbol include/fbdsp.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from fbdsp import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fbdsp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fbdsp.h but not exported"
    assert sorted(_lib.SYMBOLS) == names
    assert lib.fb_abi_version() == 1
    assert lib.fb_strerror(0) == b"ok"


def test_struct_layout_matches_header():
    # sizeof(fb_psk_design) as laid out by the C compiler == the ctypes mirror
    import subprocess, tempfile
    from fbdsp.design import fb_psk_design
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "sz.c")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "fbdsp.h"\n'
                             'int main(){printf("%zu %zu %zu\\n", sizeof(fb_psk_design), offsetof(fb_psk_design, rho), offsetof(fb_psk_design, slow_rmc));return 0;}\n')
        exe = os.path.join(td, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    assert int(out[0]) == ctypes.sizeof(fb_psk_design)
    assert int(out[1]) == fb_psk_design.rho.offset
    assert int(out[2]) == fb_psk_design.slow_rmc.offset


def test_no_gpu_fails_loudly():
    lib = _lib.load()
    if lib.fb_device_count() > 0:
        pytest.skip("GPU present")
    import fbdsp
    with pytest.raises(fbdsp.FbdspError):
        fbdsp.Engine(0)
    import numpy as np
    from fbdsp import modem
    with pytest.raises(fbdsp.FbdspError):
        modem.qpsk_demodulate(np.zeros(1000, np.float32), 9600, 9600.0)
    # parameter errors surface before any device work, exactly as in the reference
    with pytest.raises(ValueError, match="filter critical frequencies must be greater than 0"):
        modem.fsk_demodulate(np.zeros(1000, np.float32))
    with pytest.raises(TypeError):
        modem.psk8_demodulate(np.zeros(10), baud=9600, carrier=3000.0)


def test_host_only_bounds_and_struct_mirrors():
    """Pure host entry points (no device work): output bounds of the modulators / v1 demodulators, and the ctypes
    mirrors of fb_mod_params / fb_v1_params against the C compiler's layout."""
    import subprocess, tempfile
    from fbdsp import modulate as m, modem_v1 as g
    lib = _lib.load()
    p, base, env = m.psk_mod_params(m.FB_MOD_DQPSK, 9600, 9600.0, 96000)
    assert lib.fb_mod_out_samples(ctypes.byref(p), 100) == (40 + 4 * 100) * 10          # modem.py:150-186
    p, base, env = m.psk_mod_params(m.FB_MOD_DBPSK, 4800, 9600.0, 96000)
    assert lib.fb_mod_out_samples(ctypes.byref(p), 100) == (80 + 8 * 100) * 20          # modem.py:33-65
    p, t, _ = m.fsk_mod_params(1200, 1200.0, 2200.0, 96000)
    assert lib.fb_mod_out_samples(ctypes.byref(p), 0) == 8 * 4 * 80                     # preamble only (modem.py:281)
    q, table = g.psk_params(g.V1_PSK8, 38400, 12000.0)
    assert q.sps == 2 and lib.fb_v1_out_bound(ctypes.byref(q), 2001) == 1000 * 3 // 8   # App. B.6: bits truncated to x8
    q, table = g.ofdm_params(9600, 8)
    assert (q.sps, q.off0, q.len, q.nf, q.bits_per_sym) == (10, 2, 8, 7, 14)             # App. B.7: 7 carriers for OFDM8 @ 9600
    assert table[3, :, 1].tolist() == [0.0] * 8                                          # Im row of the Nyquist bin is exactly zero
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "sz.c")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "fbdsp.h"\n'
                             'int main(){printf("%zu %zu %zu %zu\\n", sizeof(fb_mod_params), offsetof(fb_mod_params, wfreq), sizeof(fb_v1_params), offsetof(fb_v1_params, bp_b));return 0;}\n')
        exe = os.path.join(td, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = [int(v) for v in subprocess.check_output([exe]).decode().split()]
    assert out[0] == ctypes.sizeof(m.fb_mod_params) and out[1] == m.fb_mod_params.wfreq.offset
    assert out[2] == ctypes.sizeof(g.fb_v1_params) and out[3] == g.fb_v1_params.bp_b.offset
