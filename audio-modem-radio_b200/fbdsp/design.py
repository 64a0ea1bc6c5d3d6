"""Host-side filter design for the v2 DPSK receive chain (float64, cached per parameter set).

The reference computes, per recording (modem.py:73-93 / 194-209):

    f  = filtfilt(butter(4, band), x)            zero-phase band-pass        |H_bp|^2
    u  = f * exp(-j w n)                         continuous LO
    bb = filtfilt(butter(4, baud/nyq), u)        zero-phase low-pass         |H_lp|^2
    s_k = bb[n0 + k*sps]

Away from the record edges this is ONE linear time-invariant map followed by a rotation:

    s_k = exp(-j w n_k) * sum_q c[q] x[n_k - q],    C(f) = |H_bp(f)|^2 |H_lp(f - fc)|^2

and the differential detector only ever sees  s_{k+1} conj(s_k) = y_{k+1} conj(y_k) * rho
with  y_k = sum_q c[q] x[n_k - q]  and the constant  rho = exp(-j w sps).  c[q] is an
exponentially decaying two-sided kernel.  Its long tail belongs to the few band-pass poles
close to the unit circle (the 0.01*nyq clamp of modem.py:76,197 puts a 480 Hz edge in most
configurations); everything else dies within ~9 symbols.  The design therefore splits

    c[q] = c_fast[q]  (|q| <= ~9 sps, evaluated as a polyphase FIR at symbol instants only)
         + sum_i R+_i p_i^q [q>0] + R-_i p_i^-q [q<0]      (slow conjugate pole pairs p_i)

The slow part is a first-order complex recursion per pole pair and direction, which
decimates exactly:  F[n+sps] = p^sps F[n] + sum_j p^(sps-j) x[n+j].

The record edges (scipy filtfilt's odd extension + lfilter_zi start-up, twice) are NOT
time-invariant; the first/last few hundred symbols are produced on the GPU by a float64
step-by-step evaluation of the reference recurrences on a window (the "edge" kernel).  This
module also sizes those zones.

Filter *coefficients* come from scipy.signal.butter exactly as in the reference, so invalid
parameters raise the same ValueError with the same text.
"""
from __future__ import annotations

import ctypes
import functools
import math
from dataclasses import dataclass, field

import numpy as np
from scipy import signal

MAX_SLOW = 4           # conjugate pole pairs handled by the recursive path (band-pass has 4 pairs)
MAX_TAPS = 8192        # complex taps (NT * sps) the FIR table can hold
FIR_TOL = 1.0e-7       # relative size of the neglected fast tail (actual is 6e-9..3e-8 after rounding NT up; fp32 evaluation adds ~1e-7)
SLOW_TOL = 1.0e-8      # warm-up truncation of the slow recursion: neglected state, relative to max|c| (residue size included)
EDGE_TOL = 1.0e-10     # decay demanded before the interior formula takes over from the edge kernel


class fb_psk_design(ctypes.Structure):
    """Mirror of `struct fb_psk_design` in include/fbdsp.h (plain C, no pointers)."""
    _fields_ = [
        ("sps", ctypes.c_int32), ("n0", ctypes.c_int32), ("bits_per_sym", ctypes.c_int32),
        ("nt", ctypes.c_int32), ("dl", ctypes.c_int32), ("dh", ctypes.c_int32),
        ("nslow", ctypes.c_int32), ("wcols", ctypes.c_int32),
        ("zone_left", ctypes.c_int32), ("zone_right", ctypes.c_int32),
        ("w_bp", ctypes.c_int32), ("w_lp", ctypes.c_int32),
        ("emulate_only", ctypes.c_int32), ("pad_bp", ctypes.c_int32), ("pad_lp", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("cycles_per_sample", ctypes.c_double),
        ("bp_b", ctypes.c_double * 9), ("bp_a", ctypes.c_double * 9), ("bp_zi", ctypes.c_double * 8),
        ("lp_b", ctypes.c_double * 5), ("lp_a", ctypes.c_double * 5), ("lp_zi", ctypes.c_double * 4),
        ("rho", ctypes.c_float * 2),
        ("slow_p", ctypes.c_double * (2 * MAX_SLOW)),       # pole (re, im), one per conjugate pair
        ("slow_lam", ctypes.c_float * (2 * MAX_SLOW)),      # p^sps
        ("slow_rp", ctypes.c_float * (2 * MAX_SLOW)),       # residue on F        (pole p,  q>0)
        ("slow_rpc", ctypes.c_float * (2 * MAX_SLOW)),      # residue on conj(F)  (pole p*, q>0)
        ("slow_rm", ctypes.c_float * (2 * MAX_SLOW)),       # residue on B        (pole p,  q<0)
        ("slow_rmc", ctypes.c_float * (2 * MAX_SLOW)),      # residue on conj(B)  (pole p*, q<0)
    ]


@dataclass
class PskDesign:
    baud: float
    carrier: float
    samp_rate: float
    band_k: float
    sps: int
    n0: int
    bits_per_sym: int
    nt: int = 0
    dl: int = 0
    dh: int = 0
    nslow: int = 0
    wcols: int = 0
    zone_left: int = 0          # samples: symbols with n_k < zone_left belong to the left edge kernel
    zone_right: int = 0         # samples: symbols with n_k > N-1-zone_right belong to the right edge kernel
    w_bp: int = 0
    w_lp: int = 0
    emulate_only: bool = False
    taps: np.ndarray = field(default=None, repr=False)        # complex64 [sps][nt]
    slow_w: np.ndarray = field(default=None, repr=False)      # complex64 [nslow][sps+1]  p^j, j=0..sps
    c_struct: fb_psk_design = field(default=None, repr=False)
    diag: dict = field(default_factory=dict, repr=False)
    taps64: np.ndarray = field(default=None, repr=False)      # complex128 [sps][nt]: the same taps before rounding
    res: list = field(default_factory=list, repr=False)       # [(p, R+, R+', R-, R-')] complex128 per slow pole pair


def _tf_eval(z, zeros, poles, gain):
    num = np.ones_like(z)
    den = np.ones_like(z)
    for q in zeros:
        num = num * (z - q)
    for q in poles:
        den = den * (z - q)
    return gain * num / den


def _tf_scalar(z, zeros, poles, gain):
    return gain * np.prod([z - q for q in zeros]) / np.prod([z - q for q in poles])


def _decay_len(radius: float, tol: float) -> int:
    if radius <= 0.0:
        return 1
    if radius >= 1.0:
        return 1 << 30
    return int(math.ceil(math.log(tol) / math.log(radius)))


@functools.lru_cache(maxsize=256)
def psk_design(baud: float, carrier: float, samp_rate: float, band_k: float, n0_is_sps: bool) -> PskDesign:
    """Design for bpsk_demodulate (band_k=1, n0=sps; modem.py:70-93) or qpsk_demodulate
    (band_k=1.5, n0=sps//2; modem.py:191-209)."""
    sps = int(samp_rate / baud)                                       # modem.py:70,191 (truncation)
    nyq = samp_rate / 2
    low = (carrier - baud * band_k) / nyq
    high = (carrier + baud * band_k) / nyq
    band = [max(0.01, low), min(0.99, high)]
    bp_b, bp_a = signal.butter(4, band, btype="band")                 # raises like the reference
    lp_b, lp_a = signal.butter(4, baud / nyq, btype="low")
    zb, pb, kb = signal.butter(4, band, btype="band", output="zpk")
    zl, pl, kl = signal.butter(4, baud / nyq, btype="low", output="zpk")
    if sps < 1:
        raise ValueError("slice step cannot be zero")                 # bb[0::0] in the reference
    d = PskDesign(baud=baud, carrier=carrier, samp_rate=samp_rate, band_k=band_k, sps=sps,
                  n0=(sps if n0_is_sps else sps // 2), bits_per_sym=(1 if n0_is_sps else 2))
    w = 2 * np.pi * carrier / samp_rate

    r_lp = float(np.max(np.abs(pl)))
    r_bp = float(np.max(np.abs(pb)))
    d.w_lp = _decay_len(r_lp, EDGE_TOL)
    d.w_bp = _decay_len(r_bp, EDGE_TOL)
    hw_min = _decay_len(r_lp, FIR_TOL)
    dl = max(1, -(-(hw_min - (sps - 1)) // sps))                      # Hneg = dl*sps + sps-1 >= hw_min
    emulate_only = False
    c = None
    for _attempt in range(6):
        dh = dl + 1
        nt = dl + dh + 1
        hneg, hpos = dl * sps + sps - 1, dh * sps
        if nt * sps > MAX_TAPS:
            emulate_only = True
            break
        # composite kernel by frequency sampling (alias-free: tails are ~0 at NF/2)
        nf = 1 << int(math.ceil(math.log2(max(64 * max(hpos, 1), 16 * min(d.w_bp, 1 << 17), 1 << 16))))
        if c is None or len(c) != nf:
            fgrid = np.arange(nf) / nf
            g_bp = np.abs(_tf_eval(np.exp(2j * np.pi * fgrid), zb, pb, kb)) ** 2
            g_lp = np.abs(_tf_eval(np.exp(2j * np.pi * fgrid - 1j * w), zl, pl, kl)) ** 2
            c = np.fft.ifft(g_bp * g_lp)
        cmax = float(np.max(np.abs(c)))
        # slow poles: band-pass poles whose memory outlives the FIR window (upper half-plane member of each pair)
        slow = [p for p in pb if p.imag > 0 and abs(p) ** hneg > FIR_TOL * 0.1]
        if any(abs(p.imag) < 1e-9 for p in pb if abs(p) ** hneg > FIR_TOL * 0.1):
            emulate_only = True                                        # real slow pole: not a conjugate pair
            break
        if len(slow) > MAX_SLOW:
            emulate_only = True
            break

        def g_lp_z(z):
            return _tf_scalar(z, zl, pl, kl) * _tf_scalar(1 / z, zl, pl, kl)

        def residues(p):
            r = kb * np.prod([p - q for q in zb]) / (p * np.prod([p - q for q in pb if q != p]))
            rho_ = r * _tf_scalar(1 / p, zb, pb, kb)                  # g_bp[k] = sum rho_i p_i^|k|
            return rho_ * g_lp_z(p * np.exp(-1j * w)), rho_ * g_lp_z(p * np.exp(1j * w))

        res = []
        for p in slow:
            pc = [q for q in pb if abs(q - np.conj(p)) < 1e-12][0]
            rp, rm = residues([q for q in pb if q == p][0])
            rpc, rmc = residues(pc)
            res.append((p, rp, rpc, rm, rmc))
        q = np.arange(-hneg - 4 * sps, hpos + 4 * sps + 1)
        cq = c[q % nf].copy()
        for p, rp, rpc, rm, rmc in res:
            pos, neg = q > 0, q < 0
            cq[pos] -= rp * p ** q[pos] + rpc * np.conj(p) ** q[pos]
            cq[neg] -= rm * p ** (-q[neg]) + rmc * np.conj(p) ** (-q[neg])
        inside = (q >= -hneg) & (q <= hpos)
        leak = float(np.max(np.abs(cq[~inside]))) / cmax
        cond = max([abs(v) for r in res for v in r[1:]] + [0.0]) / cmax
        if cond > 1e3:
            emulate_only = True                                        # modal form ill-conditioned (very narrow band)
            break
        if leak <= FIR_TOL:
            break
        dl += 1
    else:
        emulate_only = True

    d.emulate_only = emulate_only
    cs = fb_psk_design()
    cs.sps, cs.n0, cs.bits_per_sym = d.sps, d.n0, d.bits_per_sym
    cs.cycles_per_sample = carrier / samp_rate
    cs.pad_bp, cs.pad_lp = 3 * max(len(bp_a), len(bp_b)), 3 * max(len(lp_a), len(lp_b))
    for i in range(9):
        cs.bp_b[i], cs.bp_a[i] = bp_b[i], bp_a[i]
    for i in range(5):
        cs.lp_b[i], cs.lp_a[i] = lp_b[i], lp_a[i]
    for i, v in enumerate(signal.lfilter_zi(bp_b, bp_a)):
        cs.bp_zi[i] = v
    for i, v in enumerate(signal.lfilter_zi(lp_b, lp_a)):
        cs.lp_zi[i] = v
    rho = np.exp(-1j * w * sps)
    cs.rho[0], cs.rho[1] = rho.real, rho.imag
    cs.w_bp, cs.w_lp = min(d.w_bp, 1 << 28), min(d.w_lp, 1 << 28)
    cs.emulate_only = int(emulate_only)
    if not emulate_only:
        d.nt, d.dl, d.dh, d.nslow = nt, dl, dh, len(res)
        taps = np.zeros((sps, nt), dtype=np.complex128)
        for j in range(sps):
            for t in range(nt):
                qq = (t - dl) * sps - j
                taps[j, t] = cq[qq - q[0]]
        d.taps = taps.astype(np.complex64)
        d.taps64 = taps
        d.res = list(res)
        d.slow_w = np.zeros((max(1, d.nslow), sps + 1), dtype=np.complex64)
        r_slow = 0.0
        for i, (p, rp, rpc, rm, rmc) in enumerate(res):
            d.slow_w[i] = (p ** np.arange(sps + 1)).astype(np.complex64)
            lam = p ** sps
            cs.slow_p[2 * i], cs.slow_p[2 * i + 1] = p.real, p.imag
            cs.slow_lam[2 * i], cs.slow_lam[2 * i + 1] = lam.real, lam.imag
            for name, v in (("slow_rp", rp), ("slow_rpc", rpc), ("slow_rm", rm), ("slow_rmc", rmc)):
                arr = getattr(cs, name)
                arr[2 * i], arr[2 * i + 1] = v.real, v.imag
            r_slow = max(r_slow, abs(p))
        # truncating the boundary sums after w samples leaves |R| r^w of the state: size it against max|c| like the FIR tail
        d.wcols = 0 if not res else -(-_decay_len(r_slow, min(0.5, SLOW_TOL / max(cond, 1e-12))) // sps)
        # interior formula valid for symbols with  zone_left <= n_k <= N-1-zone_right
        d.zone_left = max(hpos, d.w_lp)
        d.zone_right = max(hneg, d.w_lp) + d.w_bp
        cs.nt, cs.dl, cs.dh, cs.nslow, cs.wcols = nt, dl, dh, d.nslow, d.wcols
        cs.zone_left, cs.zone_right = d.zone_left, d.zone_right
        d.diag = dict(leak=leak, cond=cond, cmax=cmax, hneg=hneg, hpos=hpos, r_lp=r_lp, r_bp=r_bp,
                      slow_radii=[abs(r[0]) for r in res])
    d.c_struct = cs
    return d
