"""CPU: fbdsp.assembly against vectors recorded from the UNMODIFIED reference's decoder.FileAssembly
(decoder.py:20-116): 40 signal-quality vectors and 40 random add_part sequences (out-of-range parts, duplicates of
higher / equal / lower quality, complete and incomplete files).  Generator: import the reference through
tools/make_golden.import_reference(), drive FileAssembly.add_part / assemble_file with seeded payloads
(np.random.default_rng(8100)), dump every observable to tests/golden/assembly.json."""
import binascii
import json
import os

import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "assembly.json")


def _cases(kind):
    return [c for c in json.load(open(GOLD)) if c["type"] == kind]


def test_signal_quality_vectors():
    from fbdsp.assembly import signal_quality
    cs = _cases("quality")
    assert len(cs) == 40
    for c in cs:
        assert signal_quality(bytes.fromhex(c["data"])) == c["quality"]


def test_add_part_sequences():
    from fbdsp.assembly import FileAssembly
    cs = _cases("sequence")
    assert len(cs) == 40
    for c in cs:
        fa = FileAssembly("f.bin", c["total"], 1234, 0xDEADBEEF)
        for op in c["ops"]:
            assert fa.add_part(op["part"], bytes.fromhex(op["data"])) == op["ret"]
        assert [None if p is None else p.hex() for p in fa.parts] == c["parts"]
        assert fa.parts_quality == c["quality"]
        assert fa.received_parts == c["received"] and fa.get_missing_parts() == c["missing"]
        assert fa.get_progress() == c["progress"]
        assert fa.get_quality_report() == c["report"]
        if c["error"] is None:
            assert fa.assemble_file().hex() == c["joined"]
        else:
            with pytest.raises(ValueError) as e:
                fa.assemble_file()
            assert str(e.value) == c["error"]


def test_assemble_stream_multi_part_job():
    """A 3-part file with a duplicated, lower-quality copy of part 1 and a second, incomplete file."""
    from fbdsp.assembly import assemble_stream
    parts = [bytes(range(i, i + 60)) for i in (0, 60, 120)]
    whole = b"".join(parts)
    crc = binascii.crc32(whole) & 0xFFFFFFFF
    fr = lambda p, d, name="a.bin", total=3, c=crc, size=len(whole): {"name": name, "data": d, "final_crc": c, "part": p,
                                                                      "total": total, "file_size": size}
    out = assemble_stream([fr(1, bytes(60)), fr(0, parts[0]), fr(1, parts[1]), fr(0, bytes(60)), fr(2, parts[2]),
                           fr(0, b"x" * 10, name="b.bin", total=2, c=7, size=20)])
    a = out[f"a.bin_{crc}"]
    assert a["complete"] and a["data"] == whole and a["size_ok"] and a["crc_ok"] and a["replaced"] == 1
    b = out["b.bin_7"]
    assert not b["complete"] and b["missing"] == [1] and b["data"] is None


MULTI = os.path.join(os.path.dirname(__file__), "golden", "multipart.json")


@pytest.mark.parametrize("mode", ["QPSK", "BPSK"])
def test_multipart_transfer_built_by_the_reference_sender(mode):
    """Parts split, compressed one by one and framed by the unmodified reference sender (tools/make_golden_multipart.py:
    encoder.split_file_for_transmission -> adaptive_compress -> _frame_data).  Both join functions must give back the
    original file -- i.e. decompress per part and key on the base name -- whatever the arrival order."""
    from fbdsp import shard
    from fbdsp.assembly import assemble_stream, file_key
    from oracle.frames import parse_fbp_stream
    g = json.load(open(MULTI))
    blob = bytes.fromhex(g["file"])
    frames = [fr for f in g["modes"][mode] for fr in parse_fbp_stream(bytes.fromhex(f["framed"]), full=True)]
    assert len(frames) == len(g["modes"][mode]) and any(f["compressed"] for f in g["modes"][mode])
    assert len({file_key(fr)[0] for fr in frames}) == 1                      # ".partN" names, one file
    for order in (list(range(len(frames))), list(reversed(range(len(frames)))), [3, 0, 1, 1, 2] + list(range(4, len(frames)))):
        for join in (assemble_stream, shard.assemble_parts):
            (f,) = join([frames[i] for i in order]).values()
            assert f["complete"] and f["data"] == blob and f["size_ok"] and f["crc_ok"] and f["name"] == "notes.txt"
    (f,) = shard.assemble_parts(frames, decompress=False).values()          # the joined compressed blobs are NOT the file
    assert f["complete"] and not f["crc_ok"]


def test_oracle_parser_has_no_table_limits():
    from oracle.frames import parse_fbp_stream
    for s in json.load(open(MULTI))["streams"]:
        got = parse_fbp_stream(bytes.fromhex(s["raw"]))
        assert [(f["name"], f["data"].hex(), f["final_crc"]) for f in got] == [(f["name"], f["data"], f["final_crc"]) for f in s["frames"]]
