"""Host-side helpers of the engine wrapper (no GPU): exact-silence detection and the emulating design copy."""
import ctypes

import numpy as np

from fbdsp import design
from fbdsp.engine import emulating_copy, has_zero_run


def test_has_zero_run_positions():
    x = np.ones(1000, np.float32)
    assert not has_zero_run(x, 10)
    for a, b in ((0, 10), (495, 505), (990, 1000)):                 # leading, interior, trailing runs of exactly min_run zeros
        y = x.copy(); y[a:b] = 0
        assert has_zero_run(y, 10) and not has_zero_run(y, 11)
    assert has_zero_run(np.zeros(10, np.float32), 10) and not has_zero_run(np.zeros(9, np.float32), 10)
    y = x.copy(); y[100:105] = 0; y[200:209] = 0                    # several short runs do not add up
    assert not has_zero_run(y, 10)
    z = np.zeros(50, np.int16); z[25] = -1                          # integer PCM, negative sample counts as non-zero
    assert has_zero_run(z, 24) and not has_zero_run(z, 26)
    assert has_zero_run(np.array([0.0, -0.0, 0.0]), 3)              # -0.0 is a zero sample


def test_emulating_copy_is_independent():
    d = design.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
    e = emulating_copy(d)
    assert e.emulate_only and e.c_struct.emulate_only == 1
    assert not d.emulate_only and d.c_struct.emulate_only == 0       # the original is untouched (ctypes structures copy by value)
    assert ctypes.addressof(e.c_struct) != ctypes.addressof(d.c_struct)
    for f in ("sps", "n0", "bits_per_sym", "pad_bp", "pad_lp", "w_bp", "w_lp"):
        assert getattr(e.c_struct, f) == getattr(d.c_struct, f)
    assert list(e.c_struct.bp_b) == list(d.c_struct.bp_b) and list(e.c_struct.lp_a) == list(d.c_struct.lp_a)


def test_has_zero_run_against_brute_force():
    rng = np.random.default_rng(12345)
    for _ in range(1500):
        n, run = int(rng.integers(1, 300)), int(rng.integers(1, 40))
        x = rng.integers(-1, 2, n).astype(np.float32)
        if rng.random() < 0.5:
            a = int(rng.integers(0, n))
            x[a: min(n, a + int(rng.integers(0, 80)))] = 0
        z = (x == 0).astype(np.int64)
        want = n >= run and bool((np.convolve(z, np.ones(run, dtype=np.int64), "valid") >= run).any())
        assert has_zero_run(x, run) == want, (n, run, x.tolist())
