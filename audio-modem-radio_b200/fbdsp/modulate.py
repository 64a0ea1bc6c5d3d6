"""Batch modulators on the device: the reference's bpsk_modulate / qpsk_modulate / fsk_modulate
(modem.py:28-65, 138-186, 270-295) for many payloads per call.  The per-symbol tables are built here with the
reference's own numpy expressions; phases and waveforms are computed by csrc/modulate.cu.  No CPU path."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .engine import Engine, default_engine

FB_MOD_DBPSK, FB_MOD_DQPSK, FB_MOD_CPFSK = 0, 1, 2


class fb_mod_params(ctypes.Structure):
    """Mirror of `struct fb_mod_params` in include/fbdsp.h."""
    _fields_ = [("kind", ctypes.c_int32), ("sps", ctypes.c_int32), ("inc", ctypes.c_double * 4), ("wfreq", ctypes.c_double * 2),
                ("gain", ctypes.c_float), ("pad", ctypes.c_float)]


def psk_mod_params(kind: int, baud, carrier, samp_rate):
    """(params, base, env) for DBPSK / DQPSK.  Raises what the reference raises when int(sps * 0.1) == 0
    (modem.py:59-61,181-183: `envelope[-0:] = np.linspace(1, 0, 0)`)."""
    sps = int(samp_rate / baud)                                  # modem.py:37,154
    t_symbol = np.arange(sps) / samp_rate
    base = np.ascontiguousarray(2 * np.pi * carrier * t_symbol, dtype=np.float64)      # modem.py:54,178
    env = np.ones(sps, dtype=np.float64)
    ramp = int(sps * 0.1)
    env[:ramp] = np.linspace(0, 1, ramp)
    env[-ramp:] = np.linspace(1, 0, ramp)                        # ramp == 0 and sps > 0: numpy's broadcast ValueError
    p = fb_mod_params()
    p.kind, p.sps = kind, sps
    inc = (0.0, np.pi, 0.0, 0.0) if kind == FB_MOD_DBPSK else (0.0, np.pi / 2, -np.pi / 2, np.pi)   # modem.py:46,162-167
    for i, v in enumerate(inc):
        p.inc[i] = v
    return p, base, env


def fsk_mod_params(baud, mark_freq, space_freq, samp_rate):
    bit_dur = 1.0 / baud
    spb = int(round(samp_rate * bit_dur))                        # modem.py:271-272
    t = np.ascontiguousarray(np.arange(spb) / samp_rate, dtype=np.float64)
    p = fb_mod_params()
    p.kind, p.sps, p.gain = FB_MOD_CPFSK, spb, 0.9
    for i, f in enumerate((space_freq, mark_freq)):              # index = bit
        p.wfreq[i] = 2 * np.pi * f                               # modem.py:288
        p.inc[i] = 2 * np.pi * f * (spb / samp_rate)             # modem.py:292
    return p, t, None


def out_samples(p: fb_mod_params, n_bytes: int, engine: Optional[Engine] = None) -> int:
    eng = engine or default_engine()
    return int(eng.lib.fb_mod_out_samples(ctypes.byref(p), int(n_bytes)))


def modulate_batch(payloads: Sequence[bytes], p: fb_mod_params, base: np.ndarray, env: Optional[np.ndarray],
                   engine: Optional[Engine] = None) -> List[np.ndarray]:
    """One float32 waveform per payload (host arrays)."""
    eng = engine or default_engine()
    n = len(payloads)
    if n == 0:
        return []
    if p.sps == 0:
        return [np.zeros(0, np.float32) for _ in payloads]
    lens = [len(b) for b in payloads]
    doff = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data = np.frombuffer(b"".join(bytes(b) for b in payloads) or b"\0", dtype=np.uint8)
    sizes = [int(eng.lib.fb_mod_out_samples(ctypes.byref(p), m)) for m in lens]
    ooff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    out = np.empty(max(1, int(ooff[-1])), dtype=np.float32)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    rc = eng.lib.fb_modulate_batch(eng.handle, ctypes.byref(p), base.ctypes.data, env.ctypes.data if env is not None else None, n,
                                   data.ctypes.data, doff.ctypes.data_as(u64p), out.ctypes.data, ooff.ctypes.data_as(u64p), 0)
    _lib.check(eng.lib, eng.handle, rc, "fb_modulate_batch")
    return [out[int(ooff[r]): int(ooff[r + 1])].copy() if n > 1 else out[: sizes[0]] for r in range(n)]


def modulate_batch_device(p: fb_mod_params, base: np.ndarray, env: Optional[np.ndarray], data_ptr: int, data_offsets: np.ndarray,
                          out_ptr: int, out_offsets: np.ndarray, engine: Optional[Engine] = None, data_on_device: bool = True):
    """Payload bytes -> float32 waveforms in caller-provided DEVICE memory (asynchronous on the engine's stream)."""
    eng = engine or default_engine()
    u64p = ctypes.POINTER(ctypes.c_uint64)
    doff = np.ascontiguousarray(data_offsets, dtype=np.uint64)
    ooff = np.ascontiguousarray(out_offsets, dtype=np.uint64)
    flags = _lib.FB_OUT_ON_DEVICE | _lib.FB_ASYNC | (_lib.FB_SAMPLES_ON_DEVICE if data_on_device else 0)
    rc = eng.lib.fb_modulate_batch(eng.handle, ctypes.byref(p), base.ctypes.data, env.ctypes.data if env is not None else None,
                                   len(doff) - 1, data_ptr, doff.ctypes.data_as(u64p), out_ptr, ooff.ctypes.data_as(u64p), flags)
    _lib.check(eng.lib, eng.handle, rc, "fb_modulate_batch")
