"""Host-side tables of the tensor-core DPSK kernel (csrc/psk_mma.cu): the interior formula of fbdsp/design.py written as
ONE banded-Toeplitz contraction per group of G symbols, evaluated by warp-level MMAs on fp16 hi/lo split operands.

A group = G = 8 consecutive symbols k0 .. k0+7 (G*sps samples).  With n_k = n0 + k*sps:

    y[k0+s] = sum_kk  x[w0' + kk] * B[kk][s]                 in-group part (FIR + the slow part of in-group sources)
            + sum_i  af[i][s] (x) F_i  +  ab[i][s] (x) Bk_i   out-of-group sources through the slow-pole states

    w0' = n_k0 - H+ - sh        window origin, moved sh (0..7) samples earlier so that it sits on a 16-byte boundary of the
                                sample buffer: the staging copy is then a plain aligned vector copy, and the shift lives in B
    B[kk][s] = c_fast[q] + [source inside the group and q != 0] c_slow[q],   q = H+ + sps*s + sh - kk
    F_i  = sum_{n <  n_k0}          p_i^(n_k0 - n) x[n]          state at the group start   F[r+1] = lam F[r] + Zf[r]
    Bk_i = sum_{n >= n_k0 + G*sps}  p_i^(n - n_k0 - G*sps) x[n]  state at the group end     Bk[r]  = Zb[r+1] + lam Bk[r+1]
    Zf_i = sum_j p_i^(G*sps - j) x[n_k0 + j],  Zb_i = sum_j p_i^j x[n_k0 + j]   (j < G*sps): one more 8-column MMA on the same rows

MMA rows are groups (16 per m-tile), the K axis is the sample window (KP = 16 * k-steps), the N axis is (symbol, re/im).
Precision: operands are split x = xh + xl, B = Bh + Bl with fp16 pieces (22 significant bits), products xh*Bh, xh*Bl, xl*Bh
are accumulated in fp32 by the tensor cores; lo products are skipped for the 16 x 8 blocks of B whose taps are below
LO_BLOCK_TOL of the largest (their 2^-11 correction is below the 1e-7 budget of the fp32 kernel).
"""
from __future__ import annotations

import functools
import math

import numpy as np

G = 8                       # symbols per group = MMA row
SX_LOG2 = 14                # samples are scaled by 2^14 before the fp16 split (|x| < 4 representable)
LO_BLOCK_TOL = 2.0e-5       # blocks of B whose largest entry is below this fraction of max|B| get no lo products


def _cfast(d, q):
    """c_fast[q] from the polyphase table taps[j][t] = c_fast[(t - dl) * sps - j]; 0 outside."""
    sps = d.sps
    j = (-q) % sps
    t = d.dl + (q + j) // sps
    if 0 <= t < d.nt:
        return d.taps64[j, t]
    return 0.0


def _cslow(d, q):
    v = 0.0
    for (p, rp, rpc, rm, rmc) in d.res:
        if q > 0:
            v += rp * p ** q + rpc * np.conj(p) ** q
        elif q < 0:
            v += rm * p ** (-q) + rmc * np.conj(p) ** (-q)
    return v


def geometry(d):
    sps = d.sps
    hpos, hneg = d.dh * sps, d.dl * sps + sps - 1
    Hp = max(hpos, (G - 1) * sps)
    Hn = max(hneg, G * sps - 1)
    kmax = 7 + (G - 1) * sps + Hp + Hn + 1               # sh <= 7
    return dict(Hp=Hp, Hn=Hn, ksteps=-(-kmax // 16))


def band_matrices(d, sh: int):
    """(B_fir [KP][2G], B_feat [KP][4*nslow]) in float64 for window shift sh."""
    sps = d.sps
    g = geometry(d)
    Hp, KP = g["Hp"], 16 * g["ksteps"]
    hpos, hneg = d.dh * sps, d.dl * sps + sps - 1
    gs = G * sps
    bfir = np.zeros((KP, 2 * G))
    nsl = len(d.res)
    bfeat = np.zeros((KP, max(4 * nsl, 1)))
    for kk in range(KP):
        jj = kk - Hp - sh                                   # source sample relative to the group start
        for s in range(G):
            q = Hp + sps * s + sh - kk
            v = 0.0
            if -hneg <= q <= hpos:
                v += _cfast(d, q)
            if 0 <= jj < gs and q != 0:
                v += _cslow(d, q)
            bfir[kk, 2 * s], bfir[kk, 2 * s + 1] = np.real(v), np.imag(v)
        if 0 <= jj < gs:
            for i, (p, *_r) in enumerate(d.res):
                zf, zb = p ** (gs - jj), p ** jj
                bfeat[kk, 4 * i: 4 * i + 4] = zf.real, zf.imag, zb.real, zb.imag
    return bfir, bfeat


def split16(a, scale):
    """fp16 hi/lo split of a * scale (both pieces unscaled relative to each other: the absolute floor of fp16, 2^-25 after
    scaling the largest entry to ~2^14, is far below the budget)."""
    v = np.asarray(a, dtype=np.float64) * scale
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def _pow2_scale(m):
    return 2.0 ** math.floor(math.log2(16384.0 / m)) if m > 0 else 1.0


def tables(d):
    """Cached on the design object (psk_design itself is cached per parameter set)."""
    t = getattr(d, "_mma_tables", None)
    if t is None:
        t = _tables(d)
        d._mma_tables = t
    return t


def _tables(d):
    """Everything the kernel needs for design d, for all 8 window shifts:
       bh, bl [8][KP][2G]  fp16 FIR band (hi, lo)      fh, fl [8][KP][4 nslow]  fp16 feature weights
       lo_blocks[ksteps][2]  bool: which 16 x 8 blocks of the FIR band need the lo products (union over shifts)
       hh_blocks[ksteps][2]  bool: which blocks are non-zero at all; feat_steps[ksteps] bool
       maps [nslow][G][2][4]  float32: real 2x2 maps of the forward / backward states for symbol s (units of the accumulators)
       lam [nslow] complex: p^(G*sps)."""
    sps = d.sps
    g = geometry(d)
    ks = g["ksteps"]
    mats = [band_matrices(d, sh) for sh in range(8)]
    mfir = max(np.max(np.abs(m[0])) for m in mats)
    st = _pow2_scale(mfir)
    sf = 16384.0 / 2                                        # |p^k| <= 1: feature weights at <= 2^13
    nsl = len(d.res)
    bh = np.zeros((8, 16 * ks, 2 * G), np.float16); bl = np.zeros_like(bh)
    fh = np.zeros((8, 16 * ks, max(4 * nsl, 1)), np.float16); fl = np.zeros_like(fh)
    hh_blocks = np.zeros((ks, 2 * G // 8), bool); lo_blocks = np.zeros_like(hh_blocks); feat_steps = np.zeros(ks, bool)
    for sh, (bf, bz) in enumerate(mats):
        bh[sh], bl[sh] = split16(bf, st)
        fh[sh], fl[sh] = split16(bz, sf)
        for k in range(ks):
            for n in range(2 * G // 8):
                blk = np.abs(bf[16 * k: 16 * k + 16, 8 * n: 8 * n + 8])
                hh_blocks[k, n] |= bool(np.any(blk != 0))
                lo_blocks[k, n] |= bool(np.max(blk) >= LO_BLOCK_TOL * mfir)
            feat_steps[k] |= bool(np.any(bz[16 * k: 16 * k + 16] != 0))
    gs = G * sps
    maps = np.zeros((max(nsl, 1), G, 2, 4), np.float64)
    lam = np.zeros(max(nsl, 1), np.complex128)
    for i, (p, rp, rpc, rm, rmc) in enumerate(d.res):
        lam[i] = p ** gs
        for s in range(G):
            for dirn, (a, b) in enumerate(((rp * p ** (sps * s), rpc * np.conj(p) ** (sps * s)),
                                           (rm * p ** (gs - sps * s), rmc * np.conj(p) ** (gs - sps * s)))):
                # a F + b conj(F) as a real 2x2 map of (Re F, Im F), in accumulator units: (Sx St y) from (Sx Sf state)
                maps[i, s, dirn] = np.array([a.real + b.real, b.imag - a.imag, a.imag + b.imag, a.real - b.real]) * (st / sf)
    return dict(ksteps=ks, KP=16 * ks, Hp=g["Hp"], Hn=g["Hn"], st=st, sf=sf, sx=2.0 ** SX_LOG2, bh=bh, bl=bl, fh=fh, fl=fl,
                hh_blocks=hh_blocks, lo_blocks=lo_blocks, feat_steps=feat_steps, maps=maps.astype(np.float32), lam=lam, nslow=nsl)


def b_fragments(mat16: np.ndarray, k: int, n: int) -> np.ndarray:
    """The mma.m16n8k16 B fragment (col layout) of block (k-step k, n-tile n) of mat16 [KP][N]: uint32 [32 lanes][2]:
    register j of lane l holds B[16k + 2 (l % 4) + 8 j + {0, 1}][8 n + l / 4], the lower k in the low half."""
    out = np.zeros((32, 2), np.uint32)
    bits = mat16.view(np.uint16)
    for l in range(32):
        col = 8 * n + l // 4
        for j in range(2):
            r = 16 * k + 2 * (l % 4) + 8 * j
            out[l, j] = int(bits[r, col]) | (int(bits[r + 1, col]) << 16)
    return out
