"""numpy model of the GPU DPSK algorithm (NOT a product path; test helper only).

It evaluates the product's host-side design (fbdsp.design) the way the CUDA kernels do --
polyphase FIR at symbol instants + decimated slow recursions for interior symbols, windowed
step-by-step evaluation of the reference recurrences for the edge zones -- so the geometry
(zones, warm-ups, initial states) can be checked against the oracle on a CPU-only box.
"""
import numpy as np
from scipy import signal


def plan(N, d):
    """Symbol/dsym ranges exactly as csrc/plan (fb_psk_plan) computes them."""
    sps, n0 = d.sps, d.n0
    nsym = -(-(N - n0) // sps) if N > n0 else 0
    ndsym = max(nsym - 1, 0)
    if d.emulate_only or nsym < 2:
        return nsym, ndsym, 0, 0
    kmin = max(0, -(-(d.zone_left - n0) // sps))
    kmax = (N - 1 - d.zone_right - n0) // sps if N - 1 - d.zone_right - n0 >= 0 else -1
    dl32 = -(-kmin // 32) * 32
    dr32 = (kmax // 32) * 32 if kmax >= 0 else 0
    if dr32 <= dl32 or dr32 > ndsym:
        return nsym, ndsym, 0, 0
    return nsym, ndsym, dl32, dr32


def interior(x, d, k_lo, k_hi):
    """y'_k for k in [k_lo, k_hi] by the interior formula (float64 here; fp32 on the GPU)."""
    cs = d.c_struct
    sps, n0, N = d.sps, d.n0, len(x)
    x = np.asarray(x, dtype=np.float64)
    ks = np.arange(k_lo, k_hi + 1)
    nk = n0 + ks * sps
    y = np.zeros(len(ks), dtype=np.complex128)
    xp = np.concatenate([x, np.zeros(d.dl * sps + 2 * sps)])
    taps = d.taps.astype(np.complex128)
    for j in range(sps):
        for t in range(d.nt):
            q = (t - d.dl) * sps - j
            y += taps[j, t] * xp[nk - q]
    pad = cs.pad_bp
    for i in range(d.nslow):
        p = complex(cs.slow_p[2 * i], cs.slow_p[2 * i + 1])
        rp = complex(cs.slow_rp[2 * i], cs.slow_rp[2 * i + 1])
        rpc = complex(cs.slow_rpc[2 * i], cs.slow_rpc[2 * i + 1])
        rm = complex(cs.slow_rm[2 * i], cs.slow_rm[2 * i + 1])
        rmc = complex(cs.slow_rmc[2 * i], cs.slow_rmc[2 * i + 1])
        # forward: virtual left extension = scipy's odd extension then constant (lfilter_zi start)
        ext = 2 * x[0] - x[pad:0:-1]
        f_state = ext[0] * p / (1 - p)                      # sum_{n < -pad} p^(-pad - n) * ext[0]
        xl = np.concatenate([ext, x])
        F = signal.lfilter([1.0], [1.0, -p], xl.astype(np.complex128), zi=[f_state * 1.0])[0]
        Fst = p * F[pad + nk - 1]                           # sum_{n<n_k} p^(n_k-n) x[n]
        xr = np.concatenate([x, np.zeros(8)])[::-1].astype(np.complex128)
        B = signal.lfilter([1.0], [1.0, -p], xr)[::-1]      # B[n] = sum_{m>=n} p^(m-n) x[m], zero extension
        Bst = B[nk] - x[nk]
        y += rp * Fst + rpc * np.conj(Fst) + rm * Bst + rmc * np.conj(Bst)
    return y


def _lfilter_win(b, a, zi, seq, exact_start):
    z0 = zi * seq[0] if exact_start else np.zeros(len(zi), dtype=seq.dtype)
    return signal.lfilter(b, a, seq, zi=z0)[0]


def edge_window(x, d, k_lo, k_hi):
    """y'_k for k in [k_lo, k_hi] by windowed evaluation of the reference recurrences
    (what the CUDA edge kernel does, with scipy standing in for the DF2T loops)."""
    cs = d.c_struct
    x = np.asarray(x, dtype=np.float64)
    N, sps, n0 = len(x), d.sps, d.n0
    pb, pl = cs.pad_bp, cs.pad_lp
    bp_b, bp_a, bp_zi = np.array(cs.bp_b), np.array(cs.bp_a), np.array(cs.bp_zi)
    lp_b, lp_a, lp_zi = np.array(cs.lp_b), np.array(cs.lp_a), np.array(cs.lp_zi)
    n_lo, n_hi = n0 + k_lo * sps, n0 + k_hi * sps
    la, lb = max(-pl, n_lo - d.w_lp), min(N - 1 + pl, n_hi + d.w_lp)
    fa, fb = max(0, la), min(N - 1, lb)                      # f needed on [fa, fb]
    if la < 0:
        fb = max(fb, min(N - 1, pl))
    if lb > N - 1:
        fa = min(fa, max(0, N - 1 - pl))
    wa, wb = max(-pb, fa - d.w_bp), min(N - 1 + pb, fb + d.w_bp)
    ext = np.concatenate([2 * x[0] - x[pb:0:-1], x, 2 * x[-1] - x[-2:-pb - 2:-1]])   # index n -> ext[n + pb]
    seg = ext[wa + pb: wb + pb + 1]
    yf = _lfilter_win(bp_b, bp_a, bp_zi, seg, wa == -pb)
    f = _lfilter_win(bp_b, bp_a, bp_zi, yf[::-1], wb == N - 1 + pb)[::-1]            # f[n] = f_arr[n - wa]
    n = np.arange(fa, fb + 1)
    ph = np.mod(cs.cycles_per_sample * n, 1.0)
    u = f[fa - wa: fb - wa + 1] * np.exp(-2j * np.pi * ph)                            # u[n] -> u_arr[n - fa]

    def u_ext(m):                                                                    # LP-stage odd extension
        m = np.asarray(m)
        out = np.empty(len(m), dtype=np.complex128)
        mid = (m >= 0) & (m <= N - 1)
        out[mid] = u[m[mid] - fa]
        lo, hi = m < 0, m > N - 1
        out[lo] = 2 * u[0 - fa] - u[-m[lo] - fa] if lo.any() else 0
        out[hi] = 2 * u[N - 1 - fa] - u[2 * (N - 1) - m[hi] - fa] if hi.any() else 0
        return out

    useq = u_ext(np.arange(la, lb + 1))
    vf = _lfilter_win(lp_b, lp_a, lp_zi.astype(np.complex128), useq, la == -pl)
    bb = _lfilter_win(lp_b, lp_a, lp_zi.astype(np.complex128), vf[::-1], lb == N - 1 + pl)[::-1]
    nk = n0 + np.arange(k_lo, k_hi + 1) * sps
    s = bb[nk - la]
    return s * np.exp(2j * np.pi * np.mod(cs.cycles_per_sample * nk, 1.0))           # un-rotate: y'_k


def slice_bits(y, d):
    cs = d.c_struct
    rho = complex(cs.rho[0], cs.rho[1])
    dd = y[1:] * np.conj(y[:-1]) * rho
    if d.bits_per_sym == 1:
        return (dd.real < 0).astype(np.uint8)
    a, b = dd.real + dd.imag, dd.real - dd.imag
    zero = (a == 0) & (b == 0)
    hi = np.where(a > 0, 0, np.where(zero, 0, 1)).astype(np.uint8)
    lo = np.where(a > 0, np.where(b > 0, 0, 1), np.where(b < 0, 1, 0)).astype(np.uint8)
    bits = np.empty(2 * len(dd), dtype=np.uint8)
    bits[0::2], bits[1::2] = hi, lo
    return bits


def demod_bits(x, d):
    """Decided bit stream for one recording, stitched from edge + interior parts."""
    N = len(x)
    nsym, ndsym, dl32, dr32 = plan(N, d)
    if nsym < 2:
        return np.zeros(0, dtype=np.uint8)
    bps = d.bits_per_sym
    if dr32 <= dl32:
        return slice_bits(edge_window(x, d, 0, nsym - 1), d)
    parts = [slice_bits(edge_window(x, d, 0, dl32), d),
             slice_bits(interior(x, d, dl32, dr32), d),
             slice_bits(edge_window(x, d, dr32, nsym - 1), d)]
    bits = np.concatenate(parts)
    assert len(bits) == ndsym * bps
    return bits
