"""CPU: the hand-written FFT's butterflies, radix schedule and Bluestein chirps (csrc/fft.cuh; `__host__ __device__`), executed on
the host through a debug hook of libfbdsp.so and compared with numpy.  The device execution of the same code is
tests/test_gpu_ingest.py::test_fft_matches_numpy."""
import ctypes

import numpy as np
import pytest

from fbdsp import _lib


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 16, 25, 27, 49, 64, 100, 105, 360, 625, 1000, 2520, 4096, 6561, 10080,
                               11, 13, 97, 1001, 4099, 17280])
def test_host_fft_matches_numpy(n):
    lib = _lib.load()
    fn = lib.fb_debug_fft_host
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    out = np.empty(n, dtype=np.complex128)
    for sign, ref in ((-1, np.fft.fft(x)), (1, np.fft.ifft(x) * n)):
        assert fn(x.ctypes.data, out.ctypes.data, n, sign) == 0
        assert np.abs(out - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()) * max(1.0, np.log2(n))
