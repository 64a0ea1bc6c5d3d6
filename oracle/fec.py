"""Oracle: restatement of the reference's fec.py.  TEST INFRASTRUCTURE ONLY.

Pinned against /root/reference/fec.py by tools/make_golden.py (tests/golden/fec.json holds
the reference's own encode/decode outputs for SURVEY Appendix C.3 vectors plus seeded
random blocks).  Neither code is a real Reed-Solomon / Viterbi decoder -- the oracle
restates what the reference computes, not what its class names promise.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


def rs_encode(data: bytes) -> bytes:
    """fec.py:11-32.  (b1, b2, b1^b2) triples; odd tail -> (b, 0xFF); + CRC32(data) LE."""
    d = np.frombuffer(data, dtype=np.uint8)
    npairs = len(d) // 2
    out = np.empty(npairs * 3 + (2 if len(d) % 2 else 0), dtype=np.uint8)
    pairs = d[: npairs * 2].reshape(-1, 2)
    body = out[: npairs * 3].reshape(-1, 3)
    body[:, 0] = pairs[:, 0]
    body[:, 1] = pairs[:, 1]
    body[:, 2] = pairs[:, 0] ^ pairs[:, 1]
    if len(d) % 2:
        out[-2] = d[-1]
        out[-1] = 0xFF
    return out.tobytes() + struct.pack("<I", zlib.crc32(data) & 0xFFFFFFFF)


def rs_decode_ex(data: bytes):
    """fec.py:34-69 -> (decoded bytes, crc_ok).  crc_ok mirrors the reference's warning print."""
    if len(data) < 4:                                              # fec.py:36-37
        return bytes(data), True
    crc_expected = struct.unpack("<I", data[-4:])[0]               # fec.py:40
    body = np.frombuffer(data[:-4], dtype=np.uint8)
    m = len(body)
    # loop fec.py:46-62: a full triple is consumed while i + 2 < m, i.e. floor(m/3) triples;
    # the remaining (m mod 3) bytes are appended raw one at a time.
    ntr = m // 3
    tri = body[: ntr * 3].reshape(-1, 3)
    dec = np.empty(ntr * 2 + (m - ntr * 3), dtype=np.uint8)
    pairs = dec[: ntr * 2].reshape(-1, 2)
    ok = (tri[:, 0] ^ tri[:, 1]) == tri[:, 2]
    pairs[:, 0] = tri[:, 0]
    pairs[:, 1] = np.where(ok, tri[:, 1], 0x3F)                    # fec.py:53-57
    dec[ntr * 2 :] = body[ntr * 3 :]
    out = dec.tobytes()
    return out, (zlib.crc32(out) & 0xFFFFFFFF) == crc_expected     # fec.py:65-67


def rs_decode(data: bytes) -> bytes:
    """ReedSolomonFEC.decode, fec.py:34-69."""
    return rs_decode_ex(data)[0]


_G1 = 0b1111001   # fec.py:76
_G2 = 0b1011011   # fec.py:77


def conv_encode(data: bytes) -> bytes:
    """ConvolutionalEncoder.encode, fec.py:79-111 (rate 1/2, K=7, 6 flush steps)."""
    bits = np.unpackbits(np.frombuffer(data, dtype=np.uint8))
    stream = np.concatenate([np.zeros(6, np.uint8), bits, np.zeros(6, np.uint8)]).astype(np.int64)
    n = len(bits) + 6
    # shift register after consuming input k: bit j of the register is stream[k + 6 - j]
    out1 = np.zeros(n, dtype=np.int64)
    out2 = np.zeros(n, dtype=np.int64)
    for j in range(7):
        tap = stream[6 - j : 6 - j + n]
        if (_G1 >> j) & 1:
            out1 ^= tap
        if (_G2 >> j) & 1:
            out2 ^= tap
    enc = np.empty(2 * n, dtype=np.uint8)
    enc[0::2] = out1
    enc[1::2] = out2
    # fec.py:103-109: the final partial byte is NOT left-aligned (missing bits are skipped)
    full = (len(enc) // 8) * 8
    out = bytearray(np.packbits(enc[:full]).tobytes())
    rem = enc[full:]
    if len(rem):
        v = 0
        for b in rem:
            v = (v << 1) | int(b)
        out.append(v)
    return bytes(out)


def viterbi_decode(data: bytes) -> bytes:
    """ViterbiDecoder.decode, fec.py:126-155: unpack MSB-first, drop the last 12 bits when
    there are >= 12, keep even-index bits, repack MSB-first with the final partial byte
    right-aligned."""
    bits = np.unpackbits(np.frombuffer(data, dtype=np.uint8))
    if len(bits) >= 12:                                            # fec.py:138-139
        bits = bits[:-12]
    used = bits[0::2]                                              # fec.py:144-146
    full = (len(used) // 8) * 8
    out = bytearray(np.packbits(used[:full]).tobytes())
    rem = used[full:]
    if len(rem):                                                   # fec.py:148-153
        v = 0
        for b in rem:
            v = (v << 1) | int(b)
        out.append(v)
    return bytes(out)
