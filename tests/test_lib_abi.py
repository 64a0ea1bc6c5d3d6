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
