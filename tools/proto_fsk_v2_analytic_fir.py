#!/usr/bin/env python
"""CPU prototype behind DESIGN.md section 9 item 2 (numpy / scipy only, no GPU): how far a LOCAL complex FIR per tone
(band-pass o Hilbert of modem.py:306-309 as one kernel) is from the reference's filtfilt + FFT-hilbert, as a function of the
distance from the record ends.  Prints the kernel supports at several truncation levels and the residual, which is the
far field of the circular Hilbert transform (scipy.signal.hilbert is an FFT over the whole record): ~ 1 / distance.

usage: python tools/proto_fsk_v2_analytic_fir.py [payload_bytes]"""
import os, sys
import numpy as np
from scipy import signal
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import signals as sig

fs, baud, mark, space = 96000, 9600, 12000.0, 24000.0
rng = np.random.default_rng(1)
data = bytes(rng.integers(0, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 6000, dtype=np.uint8))
x = sig.add_awgn(sig.fsk_modulate(data, baud=baud, mark_freq=mark, space_freq=space), 20, rng).astype(np.float64)
N, nyq = len(x), fs / 2
print("samples", N)


def exact(freq):                                            # modem.py:306-309
    b, a = signal.butter(3, [(freq - baud) / nyq, (freq + baud) / nyq], btype="band")
    return signal.hilbert(signal.filtfilt(b, a, x)), (b, a)


def kernel(b, a, L=1 << 15):                                # impulse response of filtfilt o hilbert far from any edge
    imp = np.zeros(2 * L + 1)
    imp[L] = 1.0
    f = signal.filtfilt(b, a, imp, padlen=0)
    return signal.hilbert(np.concatenate([f, [0.0]]))[:-1], L


for name, freq in (("mark", mark), ("space", space)):
    aex, (b, a) = exact(freq)
    g, L = kernel(b, a)
    mag = np.abs(g) / np.abs(g).max()
    for tol in (1e-7, 1e-10, 1e-13):
        idx = np.nonzero(mag > tol)[0]
        print(f"{name}: kernel support at {tol:g} of its peak: {idx[0] - L} .. {idx[-1] - L}")
    idx = np.nonzero(mag > 1e-13)[0]
    lo, hi = idx[0] - L, idx[-1] - L
    aloc = np.convolve(x, g[L + lo: L + hi + 1])[-lo: -lo + N]
    err = np.abs(aloc - aex) / np.abs(aex).max()
    for dist in (0, 10, 100, 1000, 10000, 100000, N // 2):
        if dist + 50 < N:
            print(f"{name}: |local FIR - reference| / peak at {dist:7d} samples from the start {err[dist:dist + 50].max():.2e}, from the end {err[N - 1 - dist - 50: N - dist].max():.2e}")

# ---- the exact decomposition (validated here to 3e-13 of the peak, everywhere including the record ends) ---------------------
#   a_ref = hilbert(filtfilt(x)) = a~ + d' + j H_circ{d'},
#   a~  = the local complex FIR applied to the zero-extended record, its tails beyond [0, N) wrapped around (periodised),
#   d'  = filtfilt(x) - Re(a~): non-zero only within ~T samples of the two record ends (1e-13 inside),
#   H_circ{d'}[n] far from the ends = (2 / N) * sum over the m of opposite parity of d'[m] cot(pi (n - m) / N): one moment of d'
#   per parity is accurate to 2e-9 of the peak at 20 000 samples from an end, two moments to 5e-12 (checked with /tmp runs).
print()
for name, freq in (("mark", mark), ("space", space)):
    aex, (b, a) = exact(freq)
    g, L = kernel(b, a)
    mag = np.abs(g) / np.abs(g).max()
    idx = np.nonzero(mag > 1e-13)[0]
    lo, hi = idx[0] - L, idx[-1] - L
    full = np.convolve(x, g[L + lo: L + hi + 1])             # linear convolution; index 0 <-> output sample `lo`
    aper = full[-lo: -lo + N].copy()
    aper[:hi] += full[-lo + N: -lo + N + hi]                  # right tail wraps to the start
    aper[N + lo:] += full[:-lo]                               # left tail wraps to the end
    filt = aex.real
    peak = np.abs(aex).max()
    delta = filt - aper.real
    T = 4000
    d2 = np.zeros(N)
    d2[:T] = delta[:T]
    d2[-T:] = delta[-T:]
    res = np.abs(aper + signal.hilbert(d2) - aex) / peak
    print(f"{name}: max|d'| / peak more than {T} samples from an end {np.abs(delta[T:N - T]).max() / peak:.1e}; "
          f"residual of periodised local FIR + circular analytic signal of the edge term {res.max():.1e}")
