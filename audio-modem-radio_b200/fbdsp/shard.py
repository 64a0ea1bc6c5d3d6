"""Multi-GPU shard scheduler: one process per GPU, independent recordings per rank, a final gather.

The receive path shards embarrassingly (decode_from_buffer is stateless per recording, decoder.py:417-464), so there
is NO collective in the data path: every rank demodulates its own recordings on its own GPU and only the small
per-recording results (raw bytes <= 0.6 % of the input, frames, status) are gathered to rank 0 -- over NCCL on the
GPU box (torch.distributed.gather_object), over gloo in the CPU tests.  Results are returned in the caller's
recording order, identical for every world size and shard order (tests/test_shard_gloo.py).

Partitioning is longest-processing-time-first bin packing on sample counts (cost is linear in samples).
The multi-part join that follows is the reference's FileAssembly semantics (decoder.py:20-116, fbdsp/assembly.py): parts
ordered by part number, duplicates replaced only by strictly higher-quality copies, size and CRC32 of the joined file
checked against the frame header.
"""
from __future__ import annotations

import binascii
import heapq
from typing import Callable, Dict, List, Optional, Sequence


def lpt_partition(lengths: Sequence[int], n_shards: int) -> List[List[int]]:
    """Indices per shard: longest first into the currently lightest bin (ties -> lowest bin id: deterministic)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    heap = [(0, s) for s in range(n_shards)]
    heapq.heapify(heap)
    shards: List[List[int]] = [[] for _ in range(n_shards)]
    for i in order:
        load, s = heapq.heappop(heap)
        shards[s].append(i)
        heapq.heappush(heap, (load + int(lengths[i]), s))
    for s in shards:
        s.sort()
    return shards


def decode_sharded(recordings: Sequence, lengths: Sequence[int], decode_fn: Callable[[List], List],
                   rank: int = 0, world_size: int = 1, dist=None, load_fn: Optional[Callable[[int], object]] = None):
    """Decode `recordings` across `world_size` ranks.

    recordings   the recordings themselves, or None entries when `load_fn(i)` fetches recording i on the rank
                 that owns it (so no rank ever touches another rank's samples)
    lengths      sample counts of ALL recordings (known to every rank: the plan is computed identically everywhere)
    decode_fn    list of recordings -> list of per-recording results (e.g. fbdsp.decoder.decode_batch on this GPU)
    dist         torch.distributed (already initialised) or None for a single process
    Returns the full result list in recording order on rank 0, None elsewhere.
    """
    shards = lpt_partition(lengths, world_size)
    mine = shards[rank]
    local_in = [recordings[i] if load_fn is None else load_fn(i) for i in mine]
    local_out = decode_fn(local_in) if mine else []
    if len(local_out) != len(mine):
        raise RuntimeError("decode_fn must return one result per recording")
    payload = list(zip(mine, local_out))
    if dist is None or world_size == 1:
        gathered = [payload]
    else:
        gathered = [None] * world_size if rank == 0 else None
        dist.gather_object(payload, gathered, dst=0)           # the only communication: small per-recording results
        if rank != 0:
            return None
    out = [None] * len(lengths)
    for part in gathered:
        for i, r in part:
            out[i] = r
    return out


def assemble_parts(frames: Sequence[dict], decompress: bool = True) -> Dict[str, dict]:
    """Join multi-part files from parsed frames ({'name','data','final_crc','part','total','file_size'}) gathered from
    all ranks, in any order: fbdsp.assembly.FileAssembly (the reference's semantics, decoder.py:20-116: part slots, a
    later copy of a part replaces the stored one only when its signal quality is strictly higher, size / CRC32 of the
    joined file checked against the frame header).  Every part is decompressed on its own first (the sender compresses
    per part, encoder.py:165-168) and files are keyed by assembly.file_key() -- one rule for this function and
    assembly.assemble_stream.  Returns {key: {'name','data' | None,'complete','size_ok','crc_ok','missing'}}."""
    from .assembly import FileAssembly, file_key, part_payload
    files: Dict[str, FileAssembly] = {}
    for fr in frames:
        key, base = file_key(fr)
        asm = files.get(key)
        if asm is None:
            asm = files[key] = FileAssembly(base, int(fr.get("total", 1)), int(fr.get("file_size", 0)), int(fr["final_crc"]))
        asm.add_part(int(fr.get("part", 0)), part_payload(fr, decompress))
    out = {}
    for key, asm in files.items():
        missing = asm.get_missing_parts()
        if missing:
            out[key] = {"name": asm.filename, "data": None, "complete": False, "size_ok": False, "crc_ok": False, "missing": missing}
            continue
        data = asm.assemble_file()
        out[key] = {"name": asm.filename, "data": data, "complete": True, "missing": [], **asm.check(data)}
    return out
