"""The reference's "v1" demodulators (bytecode-only) on the device: Goertzel FSK (+UART deframe),
FSK-HS, I/Q integrate-and-dump BPSK/QPSK/8PSK and per-symbol DFT OFDM demap -- the kernels the north_star names.
Signatures and defaults follow the disassembled v1 modem.py.  Parity: oracle/modem_v1.py only (unpinned)."""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence

import numpy as np
from scipy import signal

from . import _lib
from .engine import PADLEN_MSG, Engine, _DT, default_engine

SAMPLE_RATE = 96000
V1_BPSK, V1_QPSK, V1_PSK8, V1_OFDM, V1_FSK = range(5)


class fb_v1_params(ctypes.Structure):
    """Mirror of `struct fb_v1_params` in include/fbdsp.h."""
    _fields_ = [("mode", ctypes.c_int32), ("sps", ctypes.c_int32), ("off0", ctypes.c_int32), ("len", ctypes.c_int32),
                ("nf", ctypes.c_int32), ("bits_per_sym", ctypes.c_int32), ("uart", ctypes.c_int32), ("prefilter", ctypes.c_int32),
                ("bp_w", ctypes.c_int32), ("bp_pad", ctypes.c_int32),
                ("bp_b", ctypes.c_double * 9), ("bp_a", ctypes.c_double * 9), ("bp_zi", ctypes.c_double * 8)]


def _sps(fs, baud):
    return int(round(fs / baud))


def psk_params(mode: int, baud, carrier, fs=SAMPLE_RATE):
    p = fb_v1_params()
    p.mode, p.sps, p.off0, p.len, p.nf = mode, _sps(fs, baud), 0, _sps(fs, baud), 1
    p.bits_per_sym = {V1_BPSK: 1, V1_QPSK: 2, V1_PSK8: 3}[mode]
    t = np.arange(p.sps) / fs                                            # references restart every symbol (B.4)
    table = np.empty((1, p.sps, 2))
    table[0, :, 0], table[0, :, 1] = np.cos(2 * np.pi * carrier * t), np.sin(2 * np.pi * carrier * t)
    return p, np.ascontiguousarray(table)


def ofdm_params(baud, num_subcarriers, fs=SAMPLE_RATE):
    p = fb_v1_params()
    sps = _sps(fs, baud)
    cp = sps // 4
    useful = sps - cp
    nf = min(int(num_subcarriers), useful - 1)                           # F[1 : n_sub + 1] silently shorter (B.7)
    if nf < 1 or nf > 8:
        raise _lib.FbdspError("v1 OFDM device path supports 1..8 demapped subcarriers")
    p.mode, p.sps, p.off0, p.len, p.nf, p.bits_per_sym = V1_OFDM, sps, cp, useful, nf, 2 * nf
    # rows of the DFT matrix exactly as an FFT evaluates them: e^{-j 2 pi k m / useful} with *exact* 0 / +-1 at the
    # quarter turns, so that e.g. the Nyquist bin of a real symbol has an exactly zero imaginary part (its quadrant
    # decision in B.7 would otherwise hang on the sign of a 1e-16 rounding residue)
    dft = np.fft.fft(np.eye(useful), axis=0)
    table = np.empty((nf, useful, 2))
    for k in range(nf):
        table[k, :, 0], table[k, :, 1] = dft[k + 1].real, dft[k + 1].imag
    return p, np.ascontiguousarray(table)


def fsk_params(baud, mark, space, band_lo, band_hi, uart: bool, fs=SAMPLE_RATE):
    p = fb_v1_params()
    sps = _sps(fs, baud)
    p.mode, p.sps, p.off0, p.len, p.nf, p.bits_per_sym, p.uart, p.prefilter = V1_FSK, sps, 0, sps, 2, 1, int(uart), 1
    nyq = fs / 2
    b, a = signal.butter(4, [band_lo / nyq, band_hi / nyq], btype="band")   # B.1 bandpass_filter, may raise like scipy
    p.bp_pad = 3 * max(len(a), len(b))
    for i in range(9):
        p.bp_b[i], p.bp_a[i] = b[i], a[i]
    for i, v in enumerate(signal.lfilter_zi(b, a)):
        p.bp_zi[i] = v
    r = float(np.max(np.abs(np.roots(a))))
    p.bp_w = int(min(1 << 24, math.ceil(math.log(1e-12) / math.log(r)))) if r < 1 else 1 << 24
    j = np.arange(sps)
    table = np.empty((2, sps, 2))
    for k, f in enumerate((mark, space)):                                # Goertzel power == |sum x e^{-j w k}|^2
        w = 2 * np.pi * f / fs
        table[k, :, 0], table[k, :, 1] = np.cos(w * j), -np.sin(w * j)
    return p, np.ascontiguousarray(table)


def demod_batch(recordings: Sequence[np.ndarray], p: fb_v1_params, table: np.ndarray, engine: Optional[Engine] = None):
    """[(raw bytes, status)] per recording."""
    eng = engine or default_engine()
    n = len(recordings)
    if n == 0:
        return []
    dt = np.dtype(recordings[0].dtype)
    if any(r.dtype != dt for r in recordings) or dt not in _DT:
        raise ValueError("all recordings of one batch must share a supported dtype")
    lengths = [len(r) for r in recordings]
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    flat = np.ascontiguousarray(recordings[0] if n == 1 else np.concatenate(recordings))
    if len(flat) == 0:
        flat = np.zeros(1, dtype=dt)
    sizes = np.array([int(eng.lib.fb_v1_out_bound(ctypes.byref(p), int(m))) for m in lengths], dtype=np.uint64)
    out_offsets = np.concatenate([[np.uint64(0)], np.cumsum((sizes + np.uint64(7)) // np.uint64(4) * np.uint64(4), dtype=np.uint64)]).astype(np.uint64)
    out = np.zeros(int(out_offsets[-1]) + 8, dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    rc = eng.lib.fb_v1_demod_batch(eng.handle, ctypes.byref(p), table.ctypes.data, n, flat.ctypes.data, offsets.ctypes.data_as(u64p),
                                   _DT[dt], 0, out.ctypes.data, out_offsets.ctypes.data_as(u64p), out_len.ctypes.data, status.ctypes.data)
    _lib.check(eng.lib, eng.handle, rc, "fb_v1_demod_batch")
    return [(out[int(out_offsets[r]): int(out_offsets[r]) + int(out_len[r])].tobytes(), int(status[r])) for r in range(n)]


def _one(x, p, table, padlen=None):
    raw, st = demod_batch([x], p, table)[0]
    if st == _lib.FB_ST_TOO_SHORT:
        raise ValueError(PADLEN_MSG % (padlen or 27))
    return raw


def _f32(samples):
    return np.ascontiguousarray(np.asarray(samples, dtype=np.float32))   # every v1 demod but OFDM casts first (App. B)


def bpsk_demodulate(samples, baud=1200, carrier=3000.0, samp_rate=SAMPLE_RATE) -> bytes:
    """B.4 (pyc src 348-369)."""
    return _one(_f32(samples), *psk_params(V1_BPSK, baud, carrier, samp_rate))


def qpsk_demodulate(samples, baud=1200, carrier=3000.0, samp_rate=SAMPLE_RATE) -> bytes:
    """B.5 (pyc src 396-424)."""
    return _one(_f32(samples), *psk_params(V1_QPSK, baud, carrier, samp_rate))


def psk8_demodulate(samples, baud=2400, carrier=12000.0, samp_rate=SAMPLE_RATE) -> bytes:
    """B.6 (pyc src 454-494)."""
    return _one(_f32(samples), *psk_params(V1_PSK8, baud, carrier, samp_rate))


def ofdm_demodulate_simple(samples, baud=4800, carrier=12000.0, num_subcarriers=8, samp_rate=SAMPLE_RATE) -> bytes:
    """B.7 (pyc src 860-904): no float32 cast, carrier unused."""
    a = np.asarray(samples)
    x = np.ascontiguousarray(a if a.dtype in (np.float32, np.float64) else a.astype(np.float64))
    return _one(x, *ofdm_params(baud, num_subcarriers, samp_rate))


def fsk_demodulate(samples, baud=1200, mark=1200.0, space=2200.0, samp_rate=SAMPLE_RATE) -> bytes:
    """B.2 (pyc src 272-327): band-pass, Goertzel per bit, UART deframe."""
    if baud >= 9600:
        mark, space = 8000.0, 16000.0
    a = np.asarray(samples)
    x = np.ascontiguousarray(a if a.dtype in (np.float32, np.float64) else a.astype(np.float64))
    p, table = fsk_params(baud, mark, space, min(mark, space) - 500, max(mark, space) + 500, True, samp_rate)
    if len(x) <= p.bp_pad:
        raise ValueError(PADLEN_MSG % p.bp_pad)
    return _one(x, p, table, p.bp_pad)


def fsk_high_speed_demodulate(samples, baud=19200, mark=12000.0, space=18000.0, samp_rate=SAMPLE_RATE) -> bytes:
    """B.3 (pyc src 771-807): band-pass 8-22 kHz, Goertzel, MSB-first bytes without framing."""
    a = np.asarray(samples)
    x = np.ascontiguousarray(a if a.dtype in (np.float32, np.float64) else a.astype(np.float64))
    p, table = fsk_params(baud, mark, space, 8000, 22000, False, samp_rate)
    if len(x) <= p.bp_pad:
        raise ValueError(PADLEN_MSG % p.bp_pad)
    return _one(x, p, table, p.bp_pad)
