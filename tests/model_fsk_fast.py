"""numpy model of the factorised v2 FSK receive chain: what a kernel would evaluate instead of two whole-record
float64 `filtfilt`s and four whole-record FFTs per recording (modem.py:306-323).  Test infrastructure, like model_psk*.py.

    analytic(filtfilt(x)) = a~ + d' + j H_circ{d'}
      a~   local complex FIR g (band-pass o Hilbert far from any edge) on the zero-extended record, tails wrapped around;
      d'   = filtfilt(x) - Re a~ : non-zero only near the two record ends (needs the exact filtfilt on two edge windows only);
      H_circ{d'}[n] = (2 / N) sum_{m: n - m odd} d'[m] cot(pi (n - m) / N)       (N even; scipy.signal.hilbert is circular)
                    evaluated directly within `near` samples of an end, by `moments` moments of d' per parity beyond.
Only the samples inside the vote windows are evaluated."""
import numpy as np
from scipy import signal


def analytic_kernel(b, a, tol=1e-13, half=1 << 15):
    """Impulse response of filtfilt o hilbert far from any edge, truncated where it falls below tol of its peak."""
    imp = np.zeros(2 * half + 1)
    imp[half] = 1.0
    g = signal.hilbert(np.concatenate([signal.filtfilt(b, a, imp, padlen=0), [0.0]]))[:-1]
    idx = np.nonzero(np.abs(g) > tol * np.abs(g).max())[0]
    lo, hi = idx[0] - half, idx[-1] - half
    return g[half + lo: half + hi + 1], lo, hi


def _edge_term(x, b, a, g, lo, hi, T, W):
    """d' on [0, T) and [N - T, N): exact filtfilt on edge windows (warm-up W beyond the cut) minus Re of the wrapped local FIR."""
    N = len(x)
    head = signal.filtfilt(b, a, x[: T + W])[:T]
    tail = signal.filtfilt(b, a, x[N - T - W:])[-T:]
    # local FIR on the zero-extended record (a kernel evaluates it at the vote-window samples only), tails wrapped around
    full = np.convolve(x, g)                                                 # index 0 <-> output sample lo
    aper = full[-lo: -lo + N].copy()
    aper[:hi] += full[-lo + N: -lo + N + hi]
    aper[N + lo:] += full[:-lo]
    d = np.zeros(N)
    d[:T] = head - aper.real[:T]
    d[N - T:] = tail - aper.real[N - T:]
    return d, aper


def _hilbert_circ_at(d, idx_nz, n, N, near, moments):
    """H_circ{d}[n] for the sample indices n: direct sum near the ends, multipole expansion elsewhere."""
    out = np.zeros(len(n))
    m = idx_nz
    dm = d[m]
    mm = np.where(m < N // 2, m, m - N).astype(np.float64)                   # signed position around the seam
    dist = np.minimum(n, N - 1 - n)
    is_near = dist < near
    nn = n[is_near]
    if len(nn):
        acc = np.zeros(len(nn))
        for s in range(0, len(nn), 2048):                                    # direct sum, blocked
            blk = nn[s: s + 2048]
            k = blk[:, None] - m[None, :]
            odd = (k & 1) == 1
            with np.errstate(divide="ignore", invalid="ignore"):
                c = np.where(odd, 1.0 / np.tan(np.pi * k / N), 0.0)
            acc[s: s + 2048] = (c * dm[None, :]).sum(axis=1)
        out[is_near] = (2.0 / N) * acc
    nf = n[~is_near]
    if len(nf):
        res = np.zeros(len(nf))
        for par in (0, 1):
            sel = (m & 1) == par
            e = np.pi * mm[sel] / N
            mom = [np.sum(dm[sel] * e ** k) for k in range(moments)]
            pick = (nf & 1) != par
            xx = np.pi * nf[pick] / N
            cot, csc2 = 1.0 / np.tan(xx), 1.0 / np.sin(xx) ** 2
            terms = [cot, csc2, cot * csc2, csc2 * (csc2 + 2.0 * cot ** 2) / 3.0]      # Taylor coefficients of cot(x - e) in e
            res[pick] = sum(terms[k] * mom[k] for k in range(moments))
        out[~is_near] = (2.0 / N) * res
    return out


def fsk_fast_bits(x, baud, mark, space, fs=96000, T=4000, W=1500, near=30000, moments=2):
    """Decided bits of fsk_demodulate (before the sync search), evaluated through the factorisation.  N must be even and
    longer than 2 (T + W) (a kernel would send shorter records to the exact path)."""
    x = np.asarray(x, dtype=np.float64)
    N = len(x)
    assert N % 2 == 0 and N > 2 * (T + W) + 4096
    spb = int(fs / baud)
    q = spb // 4
    centres = np.arange(spb // 2, N, spb)
    lo_w, hi_w = centres - q, np.minimum(centres + q, N)
    n = np.concatenate([np.arange(a_, b_) for a_, b_ in zip(lo_w, hi_w)])    # vote-window samples only
    env2 = []
    for f in (mark, space):
        b, a = signal.butter(3, [(f - baud) / (fs / 2), (f + baud) / (fs / 2)], btype="band")
        g, lo, hi = analytic_kernel(b, a)
        d, aper = _edge_term(x, b, a, g, lo, hi, T, W)
        nz = np.nonzero(d)[0]
        h = _hilbert_circ_at(d, nz, n, N, near, moments)
        an = aper[n] + d[n] + 1j * h
        env2.append(an.real ** 2 + an.imag ** 2)
    cmp_ = (env2[0] > env2[1]).astype(np.int64)
    csum = np.concatenate(([0], np.cumsum(cmp_)))
    length = hi_w - lo_w
    ends = np.cumsum(length)
    ones = csum[ends] - csum[ends - length]
    return (2 * ones > length).astype(np.uint8), env2
