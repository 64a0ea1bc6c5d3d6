"""Oracle: FBPC frame build / parse.  TEST INFRASTRUCTURE ONLY.

frame_data restates encoder._frame_data (encoder.py:94-114); parse_fbp_stream restates
decoder.parse_fbp_stream_enhanced (decoder.py:142-208) without its print() side effects.
Pinned by tools/make_golden.py against the reference functions themselves.
"""
from __future__ import annotations

import binascii
import struct


def frame_data(fname: str, data: bytes, part_number: int = 0, total_parts: int = 1,
               file_size: int = 0, file_crc: int = 0) -> bytes:
    """encoder.py:94-114."""
    fname_b = fname.encode("utf-8")[:255]
    part_crc = binascii.crc32(data) & 0xFFFFFFFF
    return (b"FBPC" + bytes([len(fname_b)]) + fname_b
            + struct.pack("<IIIIII", part_number, total_parts, file_size, file_crc, len(data), part_crc)
            + data)


def parse_fbp_stream(raw: bytes, full: bool = False) -> list:
    """decoder.py:142-208.  Returns [{'name','data','final_crc'}] like the reference; with
    full=True each dict also carries the parsed-then-dropped fields and the frame offset."""
    parsed = []
    offset = 0
    starts = []
    while True:                                                  # decoder.py:155-159 (overlapping scan)
        idx = raw.find(b"FBPC", offset)
        if idx == -1:
            break
        starts.append(idx)
        offset = idx + 1
    for start in starts:
        if start + 30 > len(raw):                                # decoder.py:166
            continue
        name_len = raw[start + 4]
        if name_len == 0:                                        # decoder.py:170
            continue
        name_start = start + 5
        fname = raw[name_start: name_start + name_len].decode("utf-8", "ignore")
        meta_start = name_start + name_len
        if meta_start + 24 > len(raw):                           # decoder.py:179
            continue
        part, total, fsize, fcrc, dlen, pcrc = struct.unpack("<IIIIII", raw[meta_start: meta_start + 24])
        if dlen > 50_000_000 or dlen == 0:                       # decoder.py:184
            continue
        payload_start = meta_start + 24
        if payload_start + dlen > len(raw):                      # decoder.py:187-189
            continue
        payload = raw[payload_start: payload_start + dlen]
        if (binascii.crc32(payload) & 0xFFFFFFFF) == pcrc:       # decoder.py:194-201
            rec = {"name": fname, "data": payload, "final_crc": fcrc}
            if full:
                rec.update(part=part, total=total, file_size=fsize, offset=start, data_len=dlen)
            parsed.append(rec)
    return parsed
