"""BASELINE.json configs 3, 4 and 5 as sharded workloads (bench.py --config N).

Every config is a list of independent *units* (one recording each: a WAV part, an SNR-sweep record, a corpus recording).
All ranks build the same unit list; fbdsp.shard.decode_sharded splits it (LPT on sample counts), every rank synthesises and
decodes ONLY its own units on its own GPU -- there is no collective in the data path -- and the small per-unit results are
gathered to rank 0, which joins / audits them.  Per rank the units run in waves that fit HBM:

    synthesise the wave on the device (untimed: modulators + AWGN; set-up, not the product)
    timed, CUDA events on the engine's stream:  one batched launch sequence per parameter set (fb_psk_demod_batch /
        fb_fsk_demod_batch, device-resident), frame parse + CRC32 (fb_parse_frames_batch), and for config 3 the per-part
        fec.py decode (fb_rs_decode_spans) -- only the frame table leaves the device inside the timed region
    e2e (first wave): the same recordings as PCM16 in pinned host memory through the public Python API
        (fbdsp.decoder.decode_corpus [+ fbdsp.fec.rs_decode_batch]), wall clock, H2D and D2H inside

`value` = samples of ALL ranks / max-over-ranks of the timed device time.  The oracle (oracle/) is used outside every timed
region only: to frame / FEC-encode the inputs (TX side) and to audit a bounded sample of the outputs.
"""
from __future__ import annotations

import binascii
import ctypes
import hashlib
import os
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

FS = 96000
MAXF = 4


@dataclass
class Unit:
    idx: int
    mode: str                        # what decode_from_buffer is told (decoder.py:422-434)
    rate: int
    carrier: Optional[float] = None  # None -> the reference's hard-wired 3000.0
    tones: Optional[tuple] = None    # (mark, space[, baud]); None -> the reference's hard-wired 1200 / 2200 (design raises)
    kind: str = "qpsk"               # qpsk | bpsk | fsk | noise (unmodulated: the reference raises on its parameters anyway)
    seed: int = 0
    nbytes: int = 0                  # payload bytes (before FEC / framing); kind == noise: the record length in samples
    lead: int = 0                    # leading silence, samples
    snr: float = 20.0
    n_samples: int = 0               # record length (finish_units)
    name: str = "u.bin"
    part: int = 0
    total: int = 1
    fsize: int = 0
    fcrc: int = 0
    fec: bool = False
    ramp_free: bool = False
    payload: Optional[bytes] = field(default=None, repr=False)

    @property
    def carrier_eff(self) -> float:
        return 3000.0 if self.carrier is None else self.carrier

    @property
    def baud(self) -> int:
        return int(self.tones[2]) if (self.tones is not None and len(self.tones) > 2) else self.rate


def unit_payload(u: Unit) -> bytes:
    if u.payload is not None:
        return u.payload
    return np.random.default_rng(u.seed).integers(0, 256, u.nbytes, dtype=np.uint8).tobytes()


def unit_framed(u: Unit) -> bytes:
    """TX side (out of product scope): optional ReedSolomonFEC.encode (fec.py:11-32), then encoder._frame_data, plus four
    filler bytes -- the last differential symbol of a record has no successor to be decided against."""
    from oracle import fec as ofec
    from oracle.frames import frame_data
    data = unit_payload(u)
    if u.fec:
        data = ofec.rs_encode(data)
    return frame_data(u.name, data, u.part, u.total, u.fsize, u.fcrc) + b"\x00" * 4


def framed_len(u: Unit) -> int:
    n = u.nbytes
    if u.fec:
        n = (n // 2) * 3 + (2 if n % 2 else 0) + 4
    return 4 + 1 + len(u.name.encode()) + 24 + n + 4


def finish_units(units):
    for u in units:
        if u.kind == "noise":
            n = u.nbytes
        else:
            sps = FS // u.baud if u.kind != "fsk" else int(round(FS / u.baud))
            nsym = {"qpsk": 40 + 4 * framed_len(u), "bpsk": 80 + 8 * framed_len(u), "fsk": 8 * (4 + framed_len(u))}[u.kind]
            n = u.lead + nsym * sps
        u.n_samples = (n + 7) // 8 * 8              # records are back to back in 16-byte aligned slots (trailing silence)
    return units


# ----------------------------------------------------------------------------------------- workloads
def config3_units(file_bytes: int, part_seconds: int = 60, fec: bool = True, seed: int = 3000):
    """OFDM8 + fec.py decode of a multi-part file: split exactly as encoder.split_file_for_transmission does for mode OFDM8
    at 9600 sym/s (encoder.py:117-151: part_size = int(9600 * 60 * 0.9) bytes, names "<file>.part<i+1>"), every part
    ReedSolomonFEC-encoded, framed, sent as ofdm_modulate_simple(.., 9600, 9600.0, 8) (= DQPSK, modem.py:371).  The file
    is a seeded byte stream generated part by part, so no rank holds more than its own parts until the join on rank 0."""
    part_size = int(9600 * part_seconds * 0.9)
    total = -(-file_bytes // part_size)
    units = [Unit(idx=i, mode="OFDM8", rate=9600, carrier=9600.0, kind="qpsk", seed=seed + i, nbytes=min(part_size, file_bytes - i * part_size),
                  name=f"big.bin.part{i + 1}" if total > 1 else "big.bin", part=i, total=total, fsize=file_bytes, fec=fec) for i in range(total)]
    crc = 0
    for u in units:
        crc = binascii.crc32(unit_payload(u), crc)
    for u in units:
        u.fcrc = crc & 0xFFFFFFFF
    return finish_units(units)


def config4_units(symbols_per_point: int = 10_000_000, recs_per_point: int = 5, snrs=range(0, 31)):
    """8PSK 38 400 sym/s AWGN sweep, algo v2: psk8_demodulate == qpsk_demodulate at sps = int(2.5) = 2 (modem.py:348), carrier
    12 kHz.  The reference modulator cannot emit sps 2 (ramp = 0 crash), so the input is the ramp-free restatement of
    modem.py:138-186."""
    units = []
    nb = symbols_per_point // recs_per_point // 4
    for snr in snrs:
        for k in range(recs_per_point):
            units.append(Unit(idx=len(units), mode="8PSK", rate=38400, carrier=12000.0, kind="qpsk", seed=4000 + snr * 64 + k, nbytes=nb,
                              snr=float(snr), name=f"s{snr}_{k}.bin", ramp_free=True))
    return finish_units(units)


CORPUS_KINDS = [     # (mode, rate, carrier, tones, kind): round-tripping pairs of SURVEY 8c + the product defaults that raise
    ("QPSK", 9600, 9600.0, None, "qpsk"), ("QPSK", 3000, 3000.0, None, "qpsk"), ("8PSK", 9600, 19200.0, None, "qpsk"),
    ("OFDM4", 4800, 9600.0, None, "qpsk"), ("BPSK", 4800, 9600.0, None, "bpsk"), ("QPSK", 9600, None, None, "qpsk"),
    ("FSK1200", 1200, None, (2400.0, 4800.0), "fsk"), ("FSK19200", 19200, None, (8000.0, 16000.0, 4800), "fsk"),
    ("FSK1200", 1200, None, None, "noise"), ("FSK19200", 19200, None, None, "noise"),
]


def config5_units(n: int, min_s: float = 2.0, max_s: float = 180.0):
    """Mixed-scheme corpus: lengths log-uniform 2 s .. 3 min, leading silence 0 .. 0.5 s (whole symbols), 20 dB, seeds
    default_rng(5000 + i) ("FSK19200-class" = 4800 Bd m8000 / s16000 as the survey says)."""
    units = []
    for i in range(n):
        rng = np.random.default_rng(5000 + i)
        mode, rate, car, tones, kind = CORPUS_KINDS[int(rng.integers(0, len(CORPUS_KINDS)))]
        secs = float(np.exp(rng.uniform(np.log(min_s), np.log(max_s))))
        lead = int(rng.integers(0, FS // 2))
        u = Unit(idx=i, mode=mode, rate=rate, carrier=car, tones=tones, kind=kind, seed=50_000 + i, name=f"c{i}.bin")
        if kind == "noise":
            u.nbytes = int(secs * FS)
        else:
            sps = FS // u.baud if kind != "fsk" else int(round(FS / u.baud))
            u.lead = lead // sps * sps
            u.nbytes = max(16, int((secs * FS - u.lead) / sps / {"qpsk": 4, "bpsk": 8, "fsk": 8}[kind]) - 64)
        units.append(u)
    return finish_units(units)


# ----------------------------------------------------------------------------------------- device synthesis
def synth_group(torch, dev, eng, group: List[Unit], buf, offsets):
    """Waveforms of one parameter set into buf (float32, device), record r at offsets[r].  PSK: the phase walk of
    modem.py:44-48 / 170-174 as a float64 cumsum, sin(base + phase) * envelope (torch); CPFSK: the engine's batch modulator
    (fb_modulate_batch).  AWGN over the whole record, power from mean(x^2) of the record, one seeded
    generator per unit so that a unit's samples do not depend on which rank made them."""
    from fbdsp import modulate as fmod
    u0 = group[0]
    if u0.kind == "fsk":
        p, base, env = fmod.fsk_mod_params(u0.baud, u0.tones[0], u0.tones[1], FS)
        framed = [unit_framed(u) for u in group]
        data = torch.from_numpy(np.frombuffer(b"".join(framed), dtype=np.uint8).copy()).to(dev)
        doff = np.concatenate([[0], np.cumsum([len(f) for f in framed])]).astype(np.uint64)
        slots = np.array([int(offsets[r]) + u.lead for r, u in enumerate(group)] + [int(offsets[len(group)])], dtype=np.uint64)
        fmod.modulate_batch_device(p, base, env, data.data_ptr(), doff, buf.data_ptr(), slots, eng, data_on_device=True)
        eng.sync()
    elif u0.kind in ("qpsk", "bpsk"):
        sps = FS // u0.baud
        base = 2 * np.pi * u0.carrier_eff * torch.arange(sps, device=dev, dtype=torch.float64) / FS
        env = torch.ones(sps, dtype=torch.float64, device=dev)
        ramp = int(sps * 0.1)
        if ramp and not u0.ramp_free:
            env[:ramp] = torch.linspace(0, 1, ramp, dtype=torch.float64, device=dev)
            env[-ramp:] = torch.linspace(1, 0, ramp, dtype=torch.float64, device=dev)
        for r, u in enumerate(group):
            bits = np.unpackbits(np.frombuffer(unit_framed(u), dtype=np.uint8))
            if u.kind == "qpsk":
                bits = np.concatenate([np.array([0, 0] * 30 + [1, 1] * 10, dtype=np.uint8), bits])          # modem.py:148
                dphi = np.array([0.0, np.pi / 2, -np.pi / 2, np.pi])[bits[0::2].astype(np.int64) * 2 + bits[1::2]]   # :160-165
            else:
                bits = np.concatenate([np.tile(np.array([1, 0], np.uint8), 40), bits])                       # modem.py:33
                dphi = np.where(bits == 1, np.pi, 0.0)                                                       # modem.py:44-48
            o = int(offsets[r]) + u.lead
            ph0, step = 0.0, 1 << 20
            for s0 in range(0, len(dphi), step):                                                            # bounded temporaries
                ph = torch.cumsum(torch.from_numpy(dphi[s0:s0 + step]).to(dev), 0) + ph0
                ph0 = float(ph[-1])
                w = (torch.sin(base[None, :] + ph[:, None]) * env[None, :]).reshape(-1)
                buf[o + s0 * sps: o + s0 * sps + w.numel()] = w.to(torch.float32)
    gen = torch.Generator(device=dev)
    for r, u in enumerate(group):
        x = buf[int(offsets[r]): int(offsets[r]) + u.n_samples]
        gen.manual_seed(770_000 + u.seed)
        sigma = 0.1 if u.kind == "noise" else float(torch.sqrt(torch.mean(x.double() ** 2) / 10 ** (u.snr / 10)))
        step = 1 << 24
        for s0 in range(0, u.n_samples, step):
            m = min(step, u.n_samples - s0)
            x[s0:s0 + m] = (x[s0:s0 + m].double() + torch.randn(m, generator=gen, device=dev, dtype=torch.float64) * sigma).float()


def _frame_dtype():
    return np.dtype([("offset", "<u8"), ("name_off", "<u8"), ("payload_off", "<u8"), ("name_len", "<u4"), ("part", "<u4"), ("total", "<u4"),
                     ("file_size", "<u4"), ("file_crc", "<u4"), ("data_len", "<u4"), ("payload_crc", "<u4"), ("reserved", "<u4")])


# ----------------------------------------------------------------------------------------- the runner
class Runner:
    def __init__(self, torch, eng, dev, cfg: int):
        self.torch, self.eng, self.dev, self.cfg = torch, eng, dev, cfg
        self.dev_ms, self.samples, self.synth_s = 0.0, 0, 0.0
        self.e2e = None
        self.want_e2e = True

    def synth_wave(self, units: List[Unit]):
        torch, dev = self.torch, self.dev
        t0 = time.perf_counter()
        groups = {}
        for u in units:
            groups.setdefault((u.kind, u.mode, u.rate, u.carrier, u.tones, u.ramp_free), []).append(u)
        wave = []
        for g in groups.values():
            offsets = np.concatenate([[0], np.cumsum([u.n_samples for u in g])]).astype(np.uint64)
            buf = torch.zeros(int(offsets[-1]) + 8, dtype=torch.float32, device=dev)
            synth_group(torch, dev, self.eng, g, buf, offsets)
            wave.append((g, buf, offsets))
        torch.cuda.synchronize()
        self.synth_s += time.perf_counter() - t0
        return wave

    def decode_wave(self, wave, timed: bool = True):
        """Device-resident decode of one wave (timed with CUDA events on the engine's stream unless it is the warm-up pass
        over the same wave: workspaces grow and the allocator fills on the first pass); returns {unit idx: result dict}."""
        import fbdsp
        from fbdsp import _lib, fsk as fskmod
        from fbdsp.decoder import mode_params
        torch, eng, dev = self.torch, self.eng, self.dev
        u64p = ctypes.POINTER(ctypes.c_uint64)
        flags = _lib.FB_SAMPLES_ON_DEVICE | _lib.FB_OUT_ON_DEVICE | _lib.FB_ASYNC
        es = torch.cuda.ExternalStream(eng.stream, device=dev)
        plan = []
        for g, buf, offsets in wave:                                # designs and output buffers: set-up, not timed
            n, u0 = len(g), g[0]
            ent = {"g": g, "buf": buf, "offsets": offsets, "n": n, "err": None}
            try:
                kind, baud = mode_params(u0.mode, u0.rate)          # decoder.py:422-434
                if kind == "fsk":
                    tones = u0.tones if u0.tones is not None else (1200.0, 2200.0)
                    d = fskmod.fsk_design(u0.baud if u0.tones is not None else baud, tones[0], tones[1], float(FS))
                    ent["fsk"] = d
                    sizes = [(int(eng.lib.fb_fsk_out_bound(ctypes.byref(d), u.n_samples)) + 19) // 16 * 16 for u in g]
                else:
                    d = fbdsp.psk_design(float(baud), u0.carrier_eff, float(FS), 1.0 if kind == "bpsk" else 1.5, kind == "bpsk")
                    ent["psk"] = d
                    sizes = [(int(eng.lib.fb_psk_out_bound(ctypes.byref(d.c_struct), u.n_samples)) + 19) // 16 * 16 for u in g]
            except ValueError as e:                                 # scipy's Butterworth ValueError, as in the reference
                ent["err"] = f"ValueError: {e}"
                plan.append(ent)
                continue
            oo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
            ent.update(oo=oo, out=torch.zeros(int(oo[-1]) + 16, dtype=torch.uint8, device=dev),
                       ol=torch.zeros(n, dtype=torch.int64, device=dev), sy=torch.full((n,), -1, dtype=torch.int64, device=dev),
                       st=torch.zeros(n, dtype=torch.int32, device=dev),
                       fr=torch.zeros(n * MAXF * ctypes.sizeof(_lib.fb_frame), dtype=torch.uint8, device=dev),
                       nf=torch.zeros(n, dtype=torch.int32, device=dev), pb=torch.zeros(n, dtype=torch.int64, device=dev))
            plan.append(ent)
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(es)
        for ent in plan:
            if ent["err"]:
                continue
            buf, oo, offs = ent["buf"], ent["oo"], ent["offsets"]
            if "psk" in ent:
                eng.psk_demod_raw(ent["psk"], buf.data_ptr(), offs, _lib.FB_F32, flags, ent["out"].data_ptr(), oo,
                                  ent["ol"].data_ptr(), ent["sy"].data_ptr(), ent["st"].data_ptr())
            else:
                rc = eng.lib.fb_fsk_demod_batch(eng.handle, ctypes.byref(ent["fsk"]), ent["n"], buf.data_ptr(), offs.ctypes.data_as(u64p), _lib.FB_F32,
                                                flags, ent["out"].data_ptr(), oo.ctypes.data_as(u64p), ent["ol"].data_ptr(), ent["sy"].data_ptr(),
                                                ent["st"].data_ptr())
                _lib.check(eng.lib, eng.handle, rc, "fb_fsk_demod_batch")
            rc = eng.lib.fb_parse_frames_batch(eng.handle, ent["n"], ent["out"].data_ptr(), oo.ctypes.data_as(u64p), ent["ol"].data_ptr(), MAXF,
                                               ent["fr"].data_ptr(), ent["nf"].data_ptr(), ent["pb"].data_ptr(), flags)
            _lib.check(eng.lib, eng.handle, rc, "fb_parse_frames_batch")
            if self.cfg == 3:                                       # per-part fec.py decode where the demodulator left the payloads
                eng.sync()                                          # the frame table comes to the host (inside the timed region)
                frs = np.frombuffer(ent["fr"].cpu().numpy().tobytes(), dtype=_frame_dtype()).reshape(ent["n"], MAXF)
                nf = ent["nf"].cpu().numpy()
                starts, lens, owner = [], [], []
                for r in range(ent["n"]):
                    for k in range(min(int(nf[r]), MAXF)):
                        starts.append(int(oo[r]) + int(frs[r, k]["payload_off"])); lens.append(int(frs[r, k]["data_len"])); owner.append((r, k))
                if starts:
                    st_a, ln_a = np.array(starts, dtype=np.uint64), np.array(lens, dtype=np.uint64)
                    fo = np.concatenate([[0], np.cumsum([(int(eng.lib.fb_rs_out_bound(int(m))) + 15) // 16 * 16 for m in lens])]).astype(np.uint64)
                    ent["fec_out"] = torch.zeros(int(fo[-1]) + 16, dtype=torch.uint8, device=dev)
                    ent["fec_len"] = torch.zeros(len(starts), dtype=torch.int64, device=dev)
                    ent["fec_ok"] = torch.zeros(len(starts), dtype=torch.int32, device=dev)
                    rc = eng.lib.fb_rs_decode_spans(eng.handle, len(starts), ent["out"].data_ptr(), st_a.ctypes.data_as(u64p), ln_a.ctypes.data_as(u64p),
                                                    ent["fec_out"].data_ptr(), fo.ctypes.data_as(u64p), ent["fec_len"].data_ptr(),
                                                    ent["fec_ok"].data_ptr(), flags)
                    _lib.check(eng.lib, eng.handle, rc, "fb_rs_decode_spans")
                    ent.update(fec_off=fo, fec_owner=owner)
        e1.record(es)
        eng.sync()
        if not timed:
            return None
        self.dev_ms += e0.elapsed_time(e1)
        self.samples += sum(u.n_samples for ent in plan for u in ent["g"])
        return self.collect(plan)

    def collect(self, plan):
        """Untimed: results to the host, one dict per unit."""
        res = {}
        for ent in plan:
            g = ent["g"]
            if ent["err"]:
                for u in g:
                    res[u.idx] = {"sha": hashlib.sha256(b"").hexdigest()[:16], "sync": -1, "status": 0, "raw_len": 0, "error": ent["err"],
                                  "frames": [], "raw": None, "bits": None}
                continue
            out = ent["out"].cpu().numpy()
            ol, sy, st = ent["ol"].cpu().numpy(), ent["sy"].cpu().numpy(), ent["st"].cpu().numpy()
            frs = np.frombuffer(ent["fr"].cpu().numpy().tobytes(), dtype=_frame_dtype()).reshape(ent["n"], MAXF)
            nf = ent["nf"].cpu().numpy()
            fec = {}
            if "fec_owner" in ent:
                fo, fl, fk, fbytes = ent["fec_off"], ent["fec_len"].cpu().numpy(), ent["fec_ok"].cpu().numpy(), ent["fec_out"].cpu().numpy()
                for j, own in enumerate(ent["fec_owner"]):
                    fec[own] = (fbytes[int(fo[j]): int(fo[j]) + int(fl[j])].tobytes(), int(fk[j]))
            for r, u in enumerate(g):
                o = int(ent["oo"][r])
                raw = out[o: o + int(ol[r])].tobytes()
                frames = []
                for k in range(min(int(nf[r]), MAXF)):
                    f = frs[r, k]
                    rec = {"name": raw[int(f["name_off"]): int(f["name_off"]) + int(f["name_len"])].decode("utf-8", "ignore"),
                           "part": int(f["part"]), "total": int(f["total"]), "file_size": int(f["file_size"]), "final_crc": int(f["file_crc"]),
                           "data": raw[int(f["payload_off"]): int(f["payload_off"]) + int(f["data_len"])]}
                    if (r, k) in fec:
                        rec["data"], rec["fec_crc_ok"] = fec[(r, k)]
                    frames.append(rec)
                keep = self.cfg == 4 or (self.cfg == 5 and u.idx % 37 == 0)
                bits = self.eng.last_bits(r) if (self.cfg == 4 and "psk" in ent) else None      # last batch on this handle = this group
                res[u.idx] = {"sha": hashlib.sha256(raw).hexdigest()[:16], "sync": int(sy[r]), "status": int(st[r]), "raw_len": len(raw),
                              "error": None, "frames": frames, "raw": raw if keep else None, "bits": bits}
        return res

    def e2e_wave(self, wave):
        """Host-buffer path through the public Python API on one wave (PCM16 in pinned memory)."""
        from fbdsp import decoder, fec as ffec
        torch = self.torch
        units, recs = [], []
        pinned = torch.empty(sum(int(off[-1]) for _, _, off in wave), dtype=torch.int16, pin_memory=True)
        host = pinned.numpy()
        pos = 0
        for g, buf, offsets in wave:
            n = int(offsets[-1])
            pinned[pos:pos + n].copy_((buf[:n] * 32767.0).round_().clamp_(-32768, 32767).to(torch.int16))
            for r, u in enumerate(g):
                recs.append(host[pos + int(offsets[r]): pos + int(offsets[r]) + u.n_samples])
                units.append(u)
            pos += n
        torch.cuda.synchronize()
        args = ([u.mode for u in units], [u.rate for u in units])
        kw = dict(engine=self.eng, carriers=[u.carrier for u in units], tones=[u.tones for u in units], pcm16=True)

        def once():
            out = decoder.decode_corpus(recs, *args, **kw)
            dec = ffec.rs_decode_batch([f["data"] for r in out for f in r.frames], self.eng) if self.cfg == 3 else None
            return out, dec
        once()
        t0 = time.perf_counter()
        out, dec = once()
        dt = time.perf_counter() - t0
        n_s = sum(u.n_samples for u in units)
        return {"seconds": dt, "samples": n_s, "h2d_bytes": int(n_s * 2), "d2h_bytes": int(sum(len(r.raw) for r in out)),
                "recordings": len(units), "frames": int(sum(len(r.frames) for r in out)),
                "payload_bytes": int(sum(len(f["data"]) for r in out for f in r.frames)),
                "fec_blocks_crc_ok": (int(sum(ok for _, ok in dec)) if dec is not None else None)}


# ----------------------------------------------------------------------------------------- audits (untimed, oracle = checker)
def _audit_unit4(a):
    """Decision match of one sweep record against the oracle's float64 chain (runs in a worker process)."""
    u, x, got_bits, got_sync = a
    from oracle import modem_v2 as o2
    st = o2.qpsk_stages(x, u.rate, u.carrier_eff)
    want = st["bits"]
    if len(want) != len(got_bits):
        return u.snr, len(want) // 2, len(want) // 2, 0, False
    bad = np.nonzero(got_bits != want)[0]
    margin = o2.qpsk_margin(st["diff"])
    sym_bad = np.unique(bad // 2)
    return u.snr, len(want) // 2, len(sym_bad), int(np.sum(margin[sym_bad] >= 1e-5)), bool(got_sync == st["sync"])


def _audit_unit5(a):
    u, x, raw = a
    from oracle import modem_v2 as o2
    from fbdsp.decoder import mode_params
    kind, baud = mode_params(u.mode, u.rate)
    if kind == "fsk":
        want = o2.fsk_demodulate(x, u.baud, u.tones[0], u.tones[1])
    elif kind == "bpsk":
        want = o2.bpsk_demodulate(x, baud, u.carrier_eff)
    else:
        want = o2.qpsk_demodulate(x, baud, u.carrier_eff)
    return raw == want


def run(args, cfg: int, rank: int, world: int, local: int, result_out, clocks_cls):
    import json
    import torch
    import fbdsp
    from fbdsp import shard
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    eng = fbdsp.Engine(local)
    scale = args.scale
    if cfg == 3:
        file_bytes = int((1 << 30) * scale)
        units = config3_units(file_bytes)
        workload = f"ofdm8_fec_{file_bytes}B_file_{len(units)}_parts"
        scaling = "strong"
    elif cfg == 4:
        units = config4_units(int(10_000_000 * scale))
        workload = f"psk8_38400_awgn_0-30dB_{len(units)}_records_{int(10_000_000 * scale)}_symbols_per_point"
        scaling = "strong"
    else:
        units = config5_units(args.units if args.units > 0 else int(1250 * scale) * world)
        workload = f"mixed_corpus_{len(units)}_recordings"
        scaling = "strong" if args.units > 0 else "weak"
    lengths = [u.n_samples for u in units]
    runner = Runner(torch, eng, dev, cfg)
    wave_cap = int(args.wave_gsamples * 1e9)
    launches0 = eng.kernel_launches
    keep_wave = {}
    clocks = clocks_cls(local)
    clocks.start()

    def decode_fn(local_units):
        out, wave_units, acc = [], [], 0
        res = {}

        def flush():
            nonlocal wave_units, acc
            if not wave_units:
                return
            wave = runner.synth_wave(wave_units)
            for _ in range(min(args.warmup, 1)):                   # same wave, untimed: workspace growth, allocator, caches
                runner.decode_wave(wave, timed=False)
            r = runner.decode_wave(wave)
            if cfg == 4:                                            # audit inputs stay until the oracle has seen them
                for g, buf, offsets in wave:
                    for k, u in enumerate(g):
                        r[u.idx]["x"] = buf[int(offsets[k]): int(offsets[k]) + u.n_samples].cpu().numpy()
            if cfg == 5:
                for g, buf, offsets in wave:
                    for k, u in enumerate(g):
                        if r[u.idx]["raw"] is not None and u.n_samples <= 30 * FS:
                            r[u.idx]["x"] = buf[int(offsets[k]): int(offsets[k]) + u.n_samples].cpu().numpy()
            if runner.want_e2e and runner.e2e is None and not args.no_e2e:
                runner.e2e = runner.e2e_wave(wave)
            res.update(r)
            del wave
            torch.cuda.empty_cache()
            wave_units, acc = [], 0
        for u in local_units:
            if acc + u.n_samples > wave_cap and wave_units:
                flush()
            wave_units.append(u)
            acc += u.n_samples
        flush()
        for u in local_units:
            out.append(res[u.idx])
        return out

    def barrier():
        if dist is not None:
            dist.barrier()
        eng.sync()
        torch.cuda.synchronize()

    # warm-up: one small wave of the first parameter sets through the whole sequence (kernels, allocations), untimed
    mine = shard.lpt_partition(lengths, world)[rank]
    if mine:
        wu = Runner(torch, eng, dev, cfg)
        seen, warm = set(), []
        for i in mine:
            k = (units[i].kind, units[i].mode, units[i].rate, units[i].carrier, units[i].tones)
            if k not in seen and units[i].n_samples < 40_000_000:
                seen.add(k); warm.append(units[i])
        for _ in range(max(1, min(args.warmup, 3))):
            wu.decode_wave(wu.synth_wave(warm))
    barrier()
    t_wall = time.perf_counter()
    # local results stay local for the audits (they carry sample arrays); the gather moves only the small fields
    local_full = {}

    def decode_and_strip(local_units):
        full = decode_fn(local_units)
        slim = []
        for u, r in zip(local_units, full):
            local_full[u.idx] = r
            s = {k: v for k, v in r.items() if k not in ("raw", "bits", "x", "frames")}
            s["frames"] = [{k: v for k, v in f.items() if k != "data"} | {"len": len(f["data"]), "crc": binascii.crc32(f["data"]) & 0xFFFFFFFF}
                           for f in r["frames"]]
            slim.append(s)
        return slim

    gathered = shard.decode_sharded([None] * len(units), lengths, decode_and_strip, rank, world, dist, load_fn=lambda i: units[i])
    barrier()
    wall_s = time.perf_counter() - t_wall
    clk = clocks.stop()

    # ---- per-rank audits against the oracle on the rank's own units (host cores), then reduced --------------------------
    audit = {}
    if cfg == 4 and not args.no_audit:
        import multiprocessing as mp
        jobs = [(units[i], local_full[i]["x"], local_full[i]["bits"], local_full[i]["sync"]) for i in mine]
        with mp.get_context("fork").Pool(max(1, min(len(jobs), (os.cpu_count() or 8) // max(1, world)))) as pool:
            rows = pool.map(_audit_unit4, jobs, chunksize=1)
        audit["rows"] = rows
    if cfg == 5 and not args.no_audit:
        import multiprocessing as mp
        jobs = [(units[i], local_full[i]["x"], local_full[i]["raw"]) for i in mine if local_full[i].get("x") is not None]
        with mp.get_context("fork").Pool(max(1, min(len(jobs) or 1, (os.cpu_count() or 8) // max(1, world)))) as pool:
            oks = pool.map(_audit_unit5, jobs, chunksize=1) if jobs else []
        audit["raw_checked"], audit["raw_equal"] = len(oks), int(sum(oks))
    # config 3: the decoded parts travel to rank 0 as uint8 tensors over NCCL (no pickle of GBs), then the join
    parts_local = {}
    if cfg == 3:
        for i in mine:
            for f in local_full[i]["frames"]:
                parts_local[(i, f["part"])] = f["data"]
    t_g = time.perf_counter()
    joined = None
    if cfg == 3:
        frames_all = []
        if dist is not None:
            blob = b"".join(parts_local[k] for k in sorted(parts_local))
            meta = [(k, len(parts_local[k])) for k in sorted(parts_local)]
            metas = [None] * world if rank == 0 else None
            dist.gather_object(meta, metas, dst=0)
            sizes = torch.tensor([len(blob)], dtype=torch.int64, device=dev)
            all_sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(all_sizes, sizes)
            mx = int(max(int(s.item()) for s in all_sizes))
            t = torch.zeros(mx, dtype=torch.uint8, device=dev)
            if blob:
                t[:len(blob)] = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).to(dev)
            bufs = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
            dist.gather(t, bufs, dst=0)
            if rank == 0:
                for rk in range(world):
                    b = bufs[rk].cpu().numpy().tobytes()
                    p = 0
                    for (i, part), ln in metas[rk]:
                        parts_local[(i, part)] = b[p:p + ln]
                        p += ln
        if rank == 0:
            for i, g in enumerate(gathered):
                for f in g["frames"]:
                    frames_all.append({"name": f["name"], "part": f["part"], "total": f["total"], "file_size": f["file_size"],
                                       "final_crc": f["final_crc"], "data": parts_local[(i, f["part"])]})
            # arrival order is irrelevant to the join: feed the parts in a seeded shuffle, one of them twice
            order = np.random.default_rng(33).permutation(len(frames_all)).tolist()
            if order:
                order.append(order[0])
            files = shard.assemble_parts([frames_all[j] for j in order], decompress=False)
            joined = [{"name": v["name"], "complete": v["complete"], "size_ok": v["size_ok"], "crc_ok": v["crc_ok"], "missing": len(v["missing"]),
                       "bytes": len(v["data"]) if v["data"] is not None else 0} for v in files.values()]
    gather_s = time.perf_counter() - t_g

    # ---- reductions ----------------------------------------------------------------------------------------------------------
    def allmax(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())
    rank_ms = [runner.dev_ms]
    if dist is not None:
        t = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(t, torch.tensor([runner.dev_ms], dtype=torch.float64, device=dev))
        rank_ms = [float(v.item()) for v in t]
    dev_ms = allmax(runner.dev_ms)
    total_samples = allsum(float(runner.samples))
    launches = allsum(float(eng.kernel_launches - launches0))
    e2e_local = runner.e2e or {"seconds": 0.0, "samples": 0, "h2d_bytes": 0, "d2h_bytes": 0, "recordings": 0, "frames": 0, "payload_bytes": 0}
    e2e_s = allmax(e2e_local["seconds"])
    e2e_samples = allsum(float(e2e_local["samples"]))
    audits = [audit]
    if dist is not None:
        audits = [None] * world if rank == 0 else None
        dist.gather_object(audit, audits, dst=0)
    if rank != 0:
        dist.destroy_process_group()
        return
    inv = hashlib.sha256()
    n_err = n_frames = 0
    payload_bytes = 0
    for g in gathered:
        inv.update(f"{g['sha']}:{g['sync']}:{g['status']}:{g['error']};".encode())
        n_err += g["error"] is not None
        n_frames += len(g["frames"])
        payload_bytes += sum(f["len"] for f in g["frames"])
    line = {"metric": "demod Msamples/s", "value": total_samples / (dev_ms * 1e-3) / 1e6, "unit": "Msamples/s", "n_gpus": world, "steps": 1,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload, "baseline_config": cfg, "units": len(units), "samples": int(total_samples), "scale": scale,
                       "wave_gsamples": args.wave_gsamples, "sharding": "fbdsp.shard.decode_sharded: LPT over sample counts, no data-path collective",
                       "l2": "waves (GBs) larger than L2"},
            "payload_MB_per_s": payload_bytes / (dev_ms * 1e-3) / 1e6, "payload_bytes_valid": payload_bytes, "frames_valid": n_frames,
            "units_with_reference_error": n_err, "shard_invariance_sha256": inv.hexdigest()[:32], "gpu_launches": int(launches), "clocks": clk,
            "device_ms_per_rank": rank_ms, "wall_s_incl_synthesis": wall_s, "synth_s_rank0": runner.synth_s, "gather_s": gather_s,
            "e2e": {"value": (e2e_samples / e2e_s / 1e6) if e2e_s else None, "unit": "Msamples/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": e2e_local["h2d_bytes"], "d2h_bytes_per_step": e2e_local["d2h_bytes"],
                    "api": "fbdsp.decoder.decode_corpus" + (" + fbdsp.fec.rs_decode_batch" if cfg == 3 else ""),
                    "sample": "first wave of every rank, PCM16 in pinned host memory", "rank0": e2e_local}}
    if cfg == 3:
        line["join"] = joined
        line["fec_crc_ok_parts"] = int(sum(f.get("fec_crc_ok", 0) == 1 for g in gathered for f in g["frames"]))
    if cfg == 4 and not args.no_audit:
        per = {}
        for a in audits:
            for snr, nsym, nbad, nbad_outside, sync_ok in a.get("rows", []):
                p = per.setdefault(int(snr), [0, 0, 0, 0, 0])
                p[0] += nsym; p[1] += nbad; p[2] += nbad_outside; p[3] += int(sync_ok); p[4] += 1
        line["decision_match"] = {str(s): {"symbols": v[0], "mismatched": v[1], "match_pct": 100.0 * (1 - v[1] / max(1, v[0])),
                                           "mismatched_with_margin_ge_1e-5": v[2], "sync_equal": f"{v[3]}/{v[4]}"} for s, v in sorted(per.items())}
        line["decision_match_min_pct"] = min((v["match_pct"] for v in line["decision_match"].values()), default=None)
        line["margin_rule_violations"] = int(sum(v["mismatched_with_margin_ge_1e-5"] for v in line["decision_match"].values()))
    if cfg == 5 and not args.no_audit:
        line["raw_bytes_audit"] = {"checked": int(sum(a.get("raw_checked", 0) for a in audits)), "equal_to_oracle": int(sum(a.get("raw_equal", 0) for a in audits)),
                                   "sample": "every 37th recording of <= 30 s, all schemes"}
    print(json.dumps(line), file=result_out, flush=True)
    if dist is not None:
        dist.destroy_process_group()
