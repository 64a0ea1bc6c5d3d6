"""Host logic of the multi-GPU path on CPU: LPT partition, world_size-2 gloo gather, shard invariance, part assembly.
(The decode function is the oracle here -- the scheduler is what is under test; the GPU tests run the real engine.)"""
import os
import socket

import numpy as np
import pytest

from fbdsp import shard
from oracle import frames as ofr, modem_v2 as o2, signals as sig


def test_lpt_partition_properties():
    rng = np.random.default_rng(0)
    lengths = rng.integers(192_000, 17_280_000, 1000).tolist()
    for g in (1, 2, 4, 8):
        parts = shard.lpt_partition(lengths, g)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(1000))                       # a partition
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)         # LPT balance bound
    assert shard.lpt_partition([], 4) == [[], [], [], []]
    assert shard.lpt_partition([5], 2) == [[0], []]


def _recordings():
    recs = []
    for i, n in enumerate([300, 900, 150, 1200, 40, 700, 500]):
        _, _, x = sig.kat_signal(sig.qpsk_modulate, 400 + i, n, 20, baud=9600, carrier=9600.0, name=f"f{i}.bin")
        recs.append(x)
    return recs


def _decode(recs):
    return [o2.qpsk_stages(x, 9600, 9600.0)["raw"] for x in recs]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    recs = _recordings()
    lengths = [len(r) for r in recs]
    loaded = []

    def load(i):                                               # each rank only ever loads its own recordings
        loaded.append(i)
        return recs[i]

    out = shard.decode_sharded([None] * len(recs), lengths, _decode, rank, world, dist, load_fn=load)
    q.put((rank, out, sorted(loaded)))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_matches_single_process():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, out, loaded = q.get(timeout=180)
        got[r] = (out, loaded)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    recs = _recordings()
    single = shard.decode_sharded(recs, [len(r) for r in recs], _decode)
    assert got[0][0] == single                                   # shard invariance: same bytes, same order
    assert got[1][0] is None                                     # only rank 0 holds the gathered result
    assert sorted(got[0][1] + got[1][1]) == list(range(len(recs)))   # disjoint ownership, nothing loaded twice
    assert not set(got[0][1]) & set(got[1][1])


def test_assemble_parts():
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()
    import binascii
    crc = binascii.crc32(data) & 0xFFFFFFFF
    chunks = [data[i:i + 1800] for i in range(0, 5000, 1800)]
    streams = [ofr.frame_data(f"big.bin.part{i + 1}", c, i, len(chunks), len(data), crc) for i, c in enumerate(chunks)]
    frames = [fr for s in reversed(streams + [streams[1]]) for fr in ofr.parse_fbp_stream(s, full=True)]   # shuffled + duplicate
    files = shard.assemble_parts(frames)
    (f,) = files.values()
    assert f["complete"] and f["size_ok"] and f["crc_ok"] and f["data"] == data and f["name"] == "big.bin"
    files = shard.assemble_parts(frames[1:2])
    (f,) = files.values()
    assert not f["complete"] and f["missing"]
