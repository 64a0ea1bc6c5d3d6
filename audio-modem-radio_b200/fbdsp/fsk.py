"""v2 FSK host side (modem.py:298-341): Butterworth design with the reference's exact error behaviour, then the
device path (csrc/fsk_v2.cu).  There is no CPU path."""
from __future__ import annotations

import ctypes
import functools
import math
from typing import List, Optional, Sequence

import numpy as np
from scipy import signal

from . import _lib
from .engine import PADLEN_MSG, DemodResult, Engine, _DT, _as_samples, default_engine, flat_view


class fb_fsk_design(ctypes.Structure):
    """Mirror of `struct fb_fsk_design` in include/fbdsp.h."""
    _fields_ = [("spb", ctypes.c_int32), ("pad", ctypes.c_int32), ("w", ctypes.c_int32 * 2),
                ("b", (ctypes.c_double * 7) * 2), ("a", (ctypes.c_double * 7) * 2), ("zi", (ctypes.c_double * 6) * 2)]


def _tone(baud, freq, samp_rate):
    nyq = samp_rate / 2
    return signal.butter(3, [(freq - baud) / nyq, (freq + baud) / nyq], btype="band")       # modem.py:307


@functools.lru_cache(maxsize=128)
def _design(baud: float, mark: float, space: float, samp_rate: float, only_mark: bool = False):
    d = fb_fsk_design()
    d.spb = int(samp_rate / baud)                                                              # modem.py:301
    tones = [(_tone(baud, mark, samp_rate))]
    if not only_mark:
        tones.append(_tone(baud, space, samp_rate))
    for t, (b, a) in enumerate(tones):
        d.pad = 3 * max(len(a), len(b))
        for i in range(7):
            d.b[t][i], d.a[t][i] = b[i], a[i]
        for i, v in enumerate(signal.lfilter_zi(b, a)):
            d.zi[t][i] = v
        r = float(np.max(np.abs(np.roots(a))))
        d.w[t] = int(min(1 << 24, math.ceil(math.log(1e-12) / math.log(r)))) if r < 1 else 1 << 24
    return d


def fsk_design(baud, mark_freq, space_freq, samp_rate, n_samples: Optional[int] = None) -> fb_fsk_design:
    """Same call order as the reference: mark design -> mark filtfilt (length check) -> space design."""
    dm = _design(float(baud), float(mark_freq), float(space_freq), float(samp_rate), True)    # may raise (mark)
    if n_samples is not None and n_samples <= dm.pad:
        raise ValueError(PADLEN_MSG % dm.pad)
    return _design(float(baud), float(mark_freq), float(space_freq), float(samp_rate), False)  # may raise (space)


def fsk_demod_batch(recordings: Sequence[np.ndarray], d: fb_fsk_design, engine: Optional[Engine] = None) -> List[DemodResult]:
    eng = engine or default_engine()
    if not len(recordings):
        return []
    dt = np.dtype(recordings[0].dtype)
    if any(r.dtype != dt for r in recordings) or dt not in _DT:
        raise ValueError("all recordings of one batch must share a supported dtype")
    lengths = [len(r) for r in recordings]
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    flat = flat_view(list(recordings))
    if len(flat) == 0:
        flat = np.zeros(1, dtype=dt)
    sizes = np.array([int(eng.lib.fb_fsk_out_bound(ctypes.byref(d), int(n))) for n in lengths], dtype=np.uint64)
    out_offsets = np.concatenate([[np.uint64(0)], np.cumsum((sizes + np.uint64(3)) // np.uint64(4) * np.uint64(4), dtype=np.uint64)]).astype(np.uint64)
    n = len(recordings)
    out = np.empty(int(out_offsets[-1]) + 4, dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    sync = np.zeros(n, dtype=np.int64)
    status = np.zeros(n, dtype=np.int32)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    rc = eng.lib.fb_fsk_demod_batch(eng.handle, ctypes.byref(d), n, flat.ctypes.data, offsets.ctypes.data_as(u64p), _DT[dt], 0,
                                    out.ctypes.data, out_offsets.ctypes.data_as(u64p), out_len.ctypes.data, sync.ctypes.data,
                                    status.ctypes.data)
    _lib.check(eng.lib, eng.handle, rc, "fb_fsk_demod_batch")
    return [DemodResult(out[int(out_offsets[r]): int(out_offsets[r]) + int(out_len[r])].tobytes(), int(sync[r]), int(status[r]))
            for r in range(n)]


def demod_fsk(samples, baud, mark_freq, space_freq, samp_rate, engine: Optional[Engine] = None) -> bytes:
    x = _as_samples(samples)
    d = fsk_design(baud, mark_freq, space_freq, samp_rate, len(x))
    res = fsk_demod_batch([x], d, engine)[0]
    res.raise_for_status(d.pad)
    return res.raw
