"""numpy model of the tensor-core DPSK kernel (csrc/psk_mma.cu) -- NOT a product path; a test helper that evaluates
fbdsp.mma_tables the way the kernel does (fp16 hi/lo operands, fp32 accumulation with truncation per MMA, group-level
slow-pole recursion) so that the formulation and its precision can be checked against the oracle on a CPU-only box."""
import numpy as np

from fbdsp import mma_tables as mt


def _rz32(v):
    """float64 -> float32 rounding toward zero (what the tensor cores do when they add a k-step to the accumulator)."""
    f = np.asarray(v, dtype=np.float64).astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(v)
    return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float32)


def split_samples(x):
    """Loader semantics: xs = float32(x) * 2^14; hi = xs with the low 13 mantissa bits cleared (exact in fp16);
    lo = fp16(xs - hi)."""
    xs = (np.asarray(x, dtype=np.float32) * np.float32(2.0 ** mt.SX_LOG2)).astype(np.float32)
    hi = (xs.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (xs - hi).astype(np.float16)
    return hi.astype(np.float16), lo


def group_outputs(x, off, d, k0_first, n_groups, use_lo_skip=True):
    """Accumulator values (acc_fir [n_groups][2G], acc_feat [n_groups][4 nslow]) for groups starting at symbols
    k0_first + G r.  x = the whole recording (float), off = its element offset in the batch buffer (alignment)."""
    t = mt.tables(d)
    sps, n0 = d.sps, d.n0
    xh, xl = split_samples(x)
    xh64, xl64 = xh.astype(np.float64), xl.astype(np.float64)
    N = len(x)
    KP, ks = t["KP"], t["ksteps"]
    nn = 2 * mt.G // 8
    acc_h = np.zeros((n_groups, 2 * mt.G), np.float32)
    acc_l = np.zeros((n_groups, 2 * mt.G), np.float32)
    nf = 4 * t["nslow"]
    fac_h = np.zeros((n_groups, max(nf, 1)), np.float32)
    fac_l = np.zeros((n_groups, max(nf, 1)), np.float32)
    for r in range(n_groups):
        k0 = k0_first + mt.G * r
        w0 = n0 + k0 * sps - t["Hp"]
        sh = (off + w0) % 8
        idx = w0 - sh + np.arange(KP)
        ok = (idx >= 0) & (idx < N)
        ah = np.where(ok, xh64[np.clip(idx, 0, N - 1)], 0.0)
        al = np.where(ok, xl64[np.clip(idx, 0, N - 1)], 0.0)
        bh, bl = t["bh"][sh].astype(np.float64), t["bl"][sh].astype(np.float64)
        fh, fl = t["fh"][sh].astype(np.float64), t["fl"][sh].astype(np.float64)
        for k in range(ks):
            sl = slice(16 * k, 16 * k + 16)
            for n in range(nn):
                cs = slice(8 * n, 8 * n + 8)
                if t["hh_blocks"][k, n]:
                    acc_h[r, cs] = _rz32(acc_h[r, cs].astype(np.float64) + ah[sl] @ bh[sl, cs])
                if t["lo_blocks"][k, n] or not use_lo_skip:
                    acc_l[r, cs] = _rz32(acc_l[r, cs].astype(np.float64) + ah[sl] @ bl[sl, cs])
                    acc_l[r, cs] = _rz32(acc_l[r, cs].astype(np.float64) + al[sl] @ bh[sl, cs])
            if nf and t["feat_steps"][k]:
                fac_h[r, :nf] = _rz32(fac_h[r, :nf].astype(np.float64) + ah[sl] @ fh[sl, :nf])
                fac_l[r, :nf] = _rz32(fac_l[r, :nf].astype(np.float64) + ah[sl] @ fl[sl, :nf])
                fac_l[r, :nf] = _rz32(fac_l[r, :nf].astype(np.float64) + al[sl] @ fh[sl, :nf])
    return (acc_h + acc_l).astype(np.float32), (fac_h + fac_l).astype(np.float32)


def interior_mma(x, off, d, k_lo, k_hi, f_init=None):
    """u_k (accumulator units: Sx St y_k) for symbols k_lo .. k_hi (k_lo, k_hi + 1 multiples of G) by the kernel's
    formulation.  Slow-pole states by an exact (float64) recursion over ALL groups of the record here; the kernel carries
    the forward state along its tile range and takes the backward state from a look-ahead of one m-tile."""
    t = mt.tables(d)
    sps, n0 = d.sps, d.n0
    assert k_lo % mt.G == 0 and (k_hi + 1) % mt.G == 0
    n_groups = (k_hi + 1 - k_lo) // mt.G
    acc, feat = group_outputs(x, off, d, k_lo, n_groups)
    u = acc[:, 0::2].astype(np.complex128) + 1j * acc[:, 1::2]           # [n_groups][G]
    x64 = np.asarray(x, np.float64)
    gs = mt.G * sps
    for i, (p, *_r) in enumerate(d.res):
        # exact states in accumulator units: Sx * Sf * state
        scale = t["sx"] * t["sf"]
        F = np.zeros(n_groups, np.complex128)
        B = np.zeros(n_groups, np.complex128)
        for r in range(n_groups):
            g0 = n0 + (k_lo + mt.G * r) * sps
            n = np.arange(max(0, g0 - 4000), g0)
            F[r] = np.sum(p ** (g0 - n) * x64[n]) * scale
            m = np.arange(g0 + gs, min(len(x64), g0 + gs + 4000))
            B[r] = np.sum(p ** (m - g0 - gs) * x64[m]) * scale
        for s in range(mt.G):
            mf, mb = t["maps"][i, s, 0].astype(np.float64), t["maps"][i, s, 1].astype(np.float64)
            u[:, s] += (mf[0] * F.real + mf[1] * F.imag) + 1j * (mf[2] * F.real + mf[3] * F.imag)
            u[:, s] += (mb[0] * B.real + mb[1] * B.imag) + 1j * (mb[2] * B.real + mb[3] * B.imag)
    return u.reshape(-1), feat
