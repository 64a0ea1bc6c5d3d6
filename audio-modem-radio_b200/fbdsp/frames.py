"""FBPC frame parsing on the device: drop-in for decoder.parse_fbp_stream_enhanced (decoder.py:142-208)."""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import numpy as np

from . import _lib
from .engine import Engine, default_engine

MAX_FRAMES = 8


def parse_batch(raws: Sequence[bytes], engine: Engine = None, full: bool = False) -> List[List[dict]]:
    """Frames of every raw stream, each as the reference's {'name', 'data', 'final_crc'} dict
    (full=True adds part / total / file_size / offset / data_len)."""
    eng = engine or default_engine()
    n = len(raws)
    if n == 0:
        return []
    lens = np.fromiter((len(b) for b in raws), dtype=np.uint64, count=n)
    off = np.concatenate([[np.uint64(0)], np.cumsum((lens + np.uint64(15)) // np.uint64(16) * np.uint64(16), dtype=np.uint64)]).astype(np.uint64)
    flat = np.zeros(int(off[-1]) + 16, dtype=np.uint8)
    for i, b in enumerate(raws):
        flat[int(off[i]): int(off[i]) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    pbytes = np.zeros(n, dtype=np.uint64)
    max_frames = MAX_FRAMES
    while True:
        frames = (_lib.fb_frame * (n * max_frames))()
        nfr = np.zeros(n, dtype=np.int32)
        rc = eng.lib.fb_parse_frames_batch(eng.handle, n, flat.ctypes.data, off.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                           lens.ctypes.data, max_frames, ctypes.addressof(frames), nfr.ctypes.data,
                                           pbytes.ctypes.data, 0)
        _lib.check(eng.lib, eng.handle, rc, "fb_parse_frames_batch")
        if int(nfr.max()) <= max_frames:
            break
        max_frames = int(nfr.max())        # a recording holds more frames than the table: the reference has no limit, so neither do we
    out = []
    for i, b in enumerate(raws):
        fl = []
        for k in range(int(nfr[i])):
            f = frames[i * max_frames + k]
            rec = {"name": b[f.name_off: f.name_off + f.name_len].decode("utf-8", "ignore"),
                   "data": b[f.payload_off: f.payload_off + f.data_len], "final_crc": int(f.file_crc)}
            if full:
                rec.update(part=int(f.part), total=int(f.total), file_size=int(f.file_size), offset=int(f.offset),
                           data_len=int(f.data_len))
            fl.append(rec)
        out.append(fl)
    return out


def parse_fbp_stream_enhanced(raw: bytes) -> list:
    """decoder.py:142-208 (same return value; the reference's progress print()s stay out of the engine)."""
    return parse_batch([bytes(raw)])[0]
