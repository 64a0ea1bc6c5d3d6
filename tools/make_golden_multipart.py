#!/usr/bin/env python
"""tests/golden/multipart.json: a multi-part transfer built by the UNMODIFIED reference sender
(encoder.split_file_for_transmission -> encoder.adaptive_compress per part -> encoder._frame_data, encoder.py:117-168),
plus two parser streams the reference's parse_fbp_stream_enhanced (decoder.py:142-208) handles without any table limit:
more than 64 "FBPC" occurrences and more than 8 back-to-back frames.  Run in the build container only."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import GOLD, import_reference, quiet      # noqa: E402


def main():
    modem, fec, decoder, encoder = import_reference()
    rng = np.random.default_rng(4242)
    words = [b"filebeep", b"modem", b"FBPC", b"radio", b"part", b"crc32", b"audio", b"\n", b" ", b"0123456789"]
    blob = b"".join(words[i] for i in rng.integers(0, len(words), 2400))          # compressible, contains the magic
    path = os.path.join(os.getcwd(), "notes.txt")
    open(path, "wb").write(blob)
    out = {"file": blob.hex(), "modes": {}}
    for mode, rate in (("QPSK", 100), ("BPSK", 200)):
        parts = encoder.split_file_for_transmission(path, mode, rate, 60)
        frames = []
        for fname, data, pn, tot, fsize, fcrc in parts:
            comp = quiet(encoder.adaptive_compress, data, mode)
            frames.append({"name": fname, "part": pn, "total": tot, "file_size": fsize, "file_crc": fcrc,
                           "compressed": len(comp) != len(data), "framed": encoder._frame_data(fname, comp, pn, tot, fsize, fcrc).hex()})
        assert len(frames) > 2 and any(f["compressed"] for f in frames)
        # what the one live receive path makes of each frame (decoder.py:440-446): parse, then intelligent_decompress
        for f in frames:
            (fr,) = quiet(decoder.parse_fbp_stream_enhanced, bytes.fromhex(f["framed"]))
            from utils.compression import intelligent_decompress
            f["decoded"] = quiet(intelligent_decompress, fr["data"]).hex()
        assert b"".join(bytes.fromhex(f["decoded"]) for f in frames) == blob
        out["modes"][mode] = frames
    # parser streams beyond the device tables
    many = b"".join(encoder._frame_data(f"f{i}.bin", bytes([i]) * (5 + i), i, 12, 0, i) + b"\x00" * (i % 3) for i in range(12))
    magic = (b"FBPC" * 40 + b"xx" + encoder._frame_data("a.bin", b"FBPC" * 50 + b"tail", 0, 1, 204, 7) + b"FBPCFBPC" * 30
             + encoder._frame_data("b.bin", b"hello", 0, 1, 5, 9))
    out["streams"] = []
    for name, raw in (("twelve_frames", many), ("many_magics", magic)):
        fr = quiet(decoder.parse_fbp_stream_enhanced, raw)
        out["streams"].append({"name": name, "raw": raw.hex(), "n_magic": raw.count(b"FBPC"),
                               "frames": [{"name": f["name"], "data": f["data"].hex(), "final_crc": f["final_crc"]} for f in fr]})
    assert len(out["streams"][0]["frames"]) == 12 and out["streams"][1]["n_magic"] > 64
    json.dump(out, open(os.path.join(GOLD, "multipart.json"), "w"))
    print("wrote multipart.json:", {m: len(v) for m, v in out["modes"].items()}, [(s["name"], s["n_magic"], len(s["frames"])) for s in out["streams"]])


if __name__ == "__main__":
    main()
