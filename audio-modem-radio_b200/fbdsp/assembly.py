"""Multi-part file assembly: the reference's FileAssembly semantics (decoder.py:20-116) as the batch driver's final
gather.  In the reference this logic is unreachable (save_decoded_files unpacks 7-tuples from a list of
dicts, decoder.py:249), but it is the specification of how parts of one file are merged:

  * a file is keyed by "<name>_<file_crc>" (decoder.py:251); parts are slots 0 .. total_parts-1, out-of-range part
    numbers are ignored (decoder.py:58)
  * every part gets a "signal quality" in [0, 1] from its payload alone -- (1 - zero ratio) * (distinct bytes / 256) *
    (1 - 0.5 if the payload is one 5-byte pattern repeated) -- and a later copy of a part replaces the stored one only
    when its quality is STRICTLY higher (decoder.py:33-75)
  * a complete file is the parts joined in order; size and CRC32 mismatches against the frame header are reported, not
    fatal (decoder.py:90-104); assembling an incomplete file raises ValueError with the reference's message

Host-side integer / bytes logic: no device work here (the payload CRCs were already checked by parse_frames_kernel).
"""
from __future__ import annotations

import binascii
import re
from typing import Dict, Iterable, List, Optional


def signal_quality(data: bytes) -> float:
    """decoder.py:33-54."""
    n = len(data)
    if n == 0:
        return 0.0
    zero_ratio = data.count(0) / n
    diversity = len(set(data)) / 256
    penalty = 0.0
    if n > 10:
        whole = n - n % 5
        if data[:5] * (n // 5) == data[:whole]:
            penalty = 0.5
    q = (1 - zero_ratio) * diversity * (1 - penalty)
    return max(0.0, min(1.0, q))


class FileAssembly:
    """Same constructor, methods and return values as decoder.FileAssembly (prints and wall-clock fields dropped)."""

    def __init__(self, filename: str, total_parts: int, file_size: int, file_crc: int):
        self.filename = filename
        self.total_parts = total_parts
        self.file_size = file_size
        self.expected_crc = file_crc
        self.parts: List[Optional[bytes]] = [None] * total_parts
        self.parts_quality = [0.0] * total_parts
        self.received_parts = 0
        self.replaced = 0                       # bookkeeping the reference only prints

    calculate_signal_quality = staticmethod(signal_quality)

    def add_part(self, part_number: int, data: bytes, quality: Optional[float] = None) -> bool:
        """True when the file is complete after this part (decoder.py:56-80)."""
        if not (0 <= part_number < self.total_parts):
            return False
        if quality is None:
            quality = signal_quality(data)
        if self.parts[part_number] is None:
            self.received_parts += 1
        elif quality > self.parts_quality[part_number]:
            self.replaced += 1
        else:
            return self.received_parts == self.total_parts
        self.parts[part_number] = data
        self.parts_quality[part_number] = quality
        return self.received_parts == self.total_parts

    def get_progress(self) -> float:
        return (self.received_parts / self.total_parts) * 100 if self.total_parts > 0 else 0

    def get_missing_parts(self) -> list:
        return [i for i, p in enumerate(self.parts) if p is None]

    def assemble_file(self) -> bytes:
        if self.received_parts != self.total_parts:
            raise ValueError(f"Partes insuficientes: {self.received_parts}/{self.total_parts}. "
                             f"Faltando: {self.get_missing_parts()}")
        return b"".join(self.parts)

    def check(self, data: bytes) -> dict:
        """What assemble_file / save_decoded_files only print (decoder.py:97-102, 262-265)."""
        return {"size_ok": len(data) == self.file_size, "crc_ok": (binascii.crc32(data) & 0xFFFFFFFF) == self.expected_crc}

    def get_quality_report(self) -> dict:
        q = self.parts_quality
        return {"average_quality": sum(q) / len(q) if q else 0, "min_quality": min(q) if q else 0,
                "max_quality": max(q) if q else 0, "completed_parts": self.received_parts, "total_parts": self.total_parts}


_PART_RE = re.compile(r"\.part\d+$")


def file_key(fr: dict):
    """(key, file name) of the file a frame belongs to.  The sender names the parts of a split file "<file>.part<i+1>"
    (encoder.py:149) and stamps every part with the whole file's CRC32, so the file is "<file>_<file_crc>": the
    reference's own key (decoder.py:251) keeps the suffix, which is why its multi-part branch can never join anything."""
    name = fr["name"]
    base = _PART_RE.sub("", name) if int(fr.get("total", 1)) > 1 else name
    return f"{base}_{fr['final_crc']}", base


def part_payload(fr: dict, decompress: bool = True) -> bytes:
    """The bytes of the ORIGINAL file this frame carries.  The sender compresses every part on its own before framing
    (encoder.py:165-168, adaptive_compress, on by default) and the one live receive path decompresses per frame
    (decoder.py:446), so size / CRC32 of the joined file only hold on the decompressed parts."""
    if not decompress:
        return fr["data"]
    from .decoder import _decompress
    return _decompress(fr["data"])


def assemble_stream(frames: Iterable[dict], decompress: bool = True) -> Dict[str, dict]:
    """Feed parsed frames ({'name','data','final_crc','part','total','file_size'}, e.g. fbdsp.frames.parse_batch(...,
    full=True) over every recording of a job, in arrival order) through FileAssembly objects the way save_decoded_files
    does (decoder.py:247-275): a file is emitted -- and its assembly dropped -- the moment its last part arrives; a later
    copy of a part then starts a fresh assembly.  Parts are decompressed one by one (part_payload) and keyed by
    file_key(), the same rule fbdsp.shard.assemble_parts uses.  Returns {key: {...}} for every file emitted or pending."""
    live: Dict[str, FileAssembly] = {}
    out: Dict[str, dict] = {}
    emitted = 0
    for fr in frames:
        total = int(fr.get("total", 1))
        key, base = file_key(fr)
        asm = live.get(key)
        if asm is None:
            asm = live[key] = FileAssembly(base, total, int(fr.get("file_size", 0)), int(fr["final_crc"]))
        if asm.add_part(int(fr.get("part", 0)), part_payload(fr, decompress)):
            data = asm.assemble_file()
            rec = {"name": asm.filename, "data": data, "complete": True, "missing": [], "replaced": asm.replaced,
                   "quality": asm.get_quality_report(), **asm.check(data)}
            out[key if key not in out else f"{key}#{emitted}"] = rec
            emitted += 1
            del live[key]
    for key, asm in live.items():
        out.setdefault(key, {"name": asm.filename, "data": None, "complete": False, "missing": asm.get_missing_parts(),
                             "replaced": asm.replaced, "quality": asm.get_quality_report(), "size_ok": False, "crc_ok": False})
    return out
