"""Drop-in for the reference's fec.py *decode* side (same class names, constructor arguments and method
signatures: fec.py:7-9, 34, 114-116, 126), backed by libfbdsp.so.  Batch variants take lists of blocks.
The reference's codes are not real Reed-Solomon / Viterbi codes; the device kernels compute what it computes."""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from .engine import Engine, default_engine


def _csr(blocks: Sequence[bytes]):
    lens = np.fromiter((len(b) for b in blocks), dtype=np.uint64, count=len(blocks))
    off = np.concatenate([[np.uint64(0)], np.cumsum(lens, dtype=np.uint64)]).astype(np.uint64)
    flat = np.frombuffer(b"".join(blocks), dtype=np.uint8) if int(off[-1]) else np.zeros(1, np.uint8)
    return np.ascontiguousarray(flat), off, lens


def _u64p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def rs_decode_batch(blocks: Sequence[bytes], engine: Engine = None) -> List[Tuple[bytes, bool]]:
    """[(decoded, crc_ok)] for every block: ReedSolomonFEC.decode semantics (fec.py:34-69)."""
    eng = engine or default_engine()
    if not blocks:
        return []
    flat, off, lens = _csr(blocks)
    bound = np.array([eng.lib.fb_rs_out_bound(int(n)) for n in lens], dtype=np.uint64)
    ooff = np.concatenate([[np.uint64(0)], np.cumsum((bound + np.uint64(15)) // np.uint64(16) * np.uint64(16), dtype=np.uint64)]).astype(np.uint64)
    out = np.zeros(int(ooff[-1]) + 16, dtype=np.uint8)
    olen = np.zeros(len(blocks), dtype=np.uint64)
    ok = np.zeros(len(blocks), dtype=np.int32)
    rc = eng.lib.fb_rs_decode_batch(eng.handle, len(blocks), flat.ctypes.data, _u64p(off), out.ctypes.data, _u64p(ooff),
                                    olen.ctypes.data, ok.ctypes.data, 0)
    _lib.check(eng.lib, eng.handle, rc, "fb_rs_decode_batch")
    return [(out[int(ooff[i]): int(ooff[i]) + int(olen[i])].tobytes(), bool(ok[i] == 1)) for i in range(len(blocks))]


def viterbi_decode_batch(blocks: Sequence[bytes], engine: Engine = None) -> List[bytes]:
    """ViterbiDecoder.decode semantics (fec.py:126-155) for every block."""
    eng = engine or default_engine()
    if not blocks:
        return []
    flat, off, lens = _csr(blocks)
    bound = np.array([eng.lib.fb_viterbi_out_bound(int(n)) for n in lens], dtype=np.uint64)
    ooff = np.concatenate([[np.uint64(0)], np.cumsum((bound + np.uint64(15)) // np.uint64(16) * np.uint64(16), dtype=np.uint64)]).astype(np.uint64)
    out = np.zeros(int(ooff[-1]) + 16, dtype=np.uint8)
    olen = np.zeros(len(blocks), dtype=np.uint64)
    rc = eng.lib.fb_viterbi_decode_batch(eng.handle, len(blocks), flat.ctypes.data, _u64p(off), out.ctypes.data, _u64p(ooff),
                                         olen.ctypes.data, 0)
    _lib.check(eng.lib, eng.handle, rc, "fb_viterbi_decode_batch")
    return [out[int(ooff[i]): int(ooff[i]) + int(olen[i])].tobytes() for i in range(len(blocks))]


def crc32_batch(blocks: Sequence[bytes], engine: Engine = None) -> List[int]:
    """zlib.crc32 of every block on the device."""
    eng = engine or default_engine()
    if not blocks:
        return []
    flat, off, _ = _csr(blocks)
    crc = np.zeros(len(blocks), dtype=np.uint32)
    rc = eng.lib.fb_crc32_batch(eng.handle, len(blocks), flat.ctypes.data, _u64p(off), crc.ctypes.data, 0)
    _lib.check(eng.lib, eng.handle, rc, "fb_crc32_batch")
    return [int(c) for c in crc]


class ReedSolomonFEC:
    """fec.py:7-69 (decode side)."""

    def __init__(self, nsym=32):
        self.nsym = nsym

    def decode(self, data: bytes) -> bytes:
        out, ok = rs_decode_batch([bytes(data)])[0]
        if len(data) >= 4 and not ok:
            print("Aviso: CRC não corresponde - dados podem estar corrompidos")     # fec.py:66-67
        return out


class ViterbiDecoder:
    """fec.py:114-155."""

    def __init__(self, constraint_length=7):
        self.constraint_length = constraint_length
        self.g1 = 0b1111001
        self.g2 = 0b1011011
        self.trellis = {}

    def decode(self, data: bytes) -> bytes:
        return viterbi_decode_batch([bytes(data)])[0]
