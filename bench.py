#!/usr/bin/env python
"""bench.py -- FileBeep receive hot path on B200: DQPSK 9600 sym/s batch demodulation of 256 synthetic
3-minute 96 kHz parts per GPU (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One process per GPU (torchrun for N > 1, weak scaling: every rank demodulates its own 256 recordings, no
data-path collective).  A step = one pass of the hot path over the whole batch:
psk_edge_kernel + psk_main_kernel + sync search + byte packing + FBPC frame parse with CRC32 (all on the device).  `value` is timed with CUDA events on the engine's own stream with
the batch resident in HBM; `e2e` goes through the host-buffer C-ABI call (pinned host samples -> H2D -> kernels
-> D2H of the raw bytes) inside the timed region.  Inputs (17.7 GB per GPU) are far larger than L2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "audio-modem-radio_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

FS = 96000
BAUD, CARRIER = 9600, 9600.0          # round-tripping DQPSK pair; raw-byte parity also holds at 3000
METRIC = "demod Msamples/s (DQPSK 9600 sym/s, 3-min 96 kHz parts)"


# ----------------------------------------------------------------------------------------- synthetic batch
def make_recording_bits(k: int, n_samples: int, sps: int):
    """Seeded payload -> FBPC frame -> DQPSK phase increments (host; needs CRC32 of the payload)."""
    from oracle.frames import frame_data
    rng = np.random.default_rng(1000 + k)
    n_sym_max = n_samples // sps
    payload_len = max(16, (n_sym_max - 40) // 4 - 64)          # fill the part: 4 symbols per byte + header + preamble
    payload = rng.integers(0, 256, payload_len, dtype=np.uint8).tobytes()
    framed = frame_data(f"part{k}.bin", payload, k, 256, payload_len * 256, 0)
    bits = np.unpackbits(np.frombuffer(framed, dtype=np.uint8))
    pre = np.array([0, 0] * 30 + [1, 1] * 10, dtype=np.uint8)    # modem.py:148
    bits = np.concatenate([pre, bits])
    code = bits[0::2].astype(np.int64) * 2 + bits[1::2]
    table = np.array([0.0, np.pi / 2, -np.pi / 2, np.pi])         # modem.py:160-165
    return payload_len, table[code]


def synth_batch_device(torch, dev, n_rec: int, n_samples: int, snr_db: float, rank: int):
    """DQPSK waveforms per modem.py:138-186 synthesised on the GPU with torch (set-up only, not the product),
    zero-padded to n_samples, AWGN at snr_db over the whole record."""
    sps = FS // BAUD
    batch = torch.empty(n_rec * n_samples, dtype=torch.float32, device=dev)
    t_sym = torch.arange(sps, device=dev, dtype=torch.float64) / FS
    base = 2 * np.pi * CARRIER * t_sym
    env = torch.ones(sps, dtype=torch.float64, device=dev)
    ramp = int(sps * 0.1)
    if ramp:
        env[:ramp] = torch.linspace(0, 1, ramp, dtype=torch.float64, device=dev)
        env[-ramp:] = torch.linspace(1, 0, ramp, dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev)
    payload_bytes = 0
    for k in range(n_rec):
        plen, dphi = make_recording_bits(rank * n_rec + k, n_samples, sps)
        payload_bytes += plen
        ph = torch.cumsum(torch.from_numpy(dphi).to(dev), 0)
        w = (torch.sin(base[None, :] + ph[:, None]) * env[None, :]).reshape(-1)
        x = torch.zeros(n_samples, dtype=torch.float64, device=dev)
        m = min(n_samples, w.numel())
        x[:m] = w[:m]
        gen.manual_seed(77000 + rank * n_rec + k)
        sigma = float(torch.sqrt(torch.mean(x * x) / 10 ** (snr_db / 10)))
        x += torch.randn(n_samples, generator=gen, device=dev, dtype=torch.float64) * sigma
        batch[k * n_samples:(k + 1) * n_samples] = x.to(torch.float32)
    return batch, payload_bytes


# ----------------------------------------------------------------------------------------- clocks
def bind_rank_to_gpu_cpus(local: int, world: int):
    """Binds this rank's threads to its share of the CPUs next to its GPU (NVML affinity mask, else the PCI device's
    local_cpulist; when every GPU reports the same set the ranks split it) BEFORE any pinned buffer is allocated, so that the
    staging pages are first-touched on that socket.  Returns a description for the JSON line."""
    info = {"bound": False}
    try:
        allowed = sorted(os.sched_getaffinity(0))
        bind_rank_to_gpu_cpus.allowed = allowed          # restored before the CPU baseline leg (it uses every host core)
        cpus, numa = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1]
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            node_path = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
            if os.path.exists(node_path):
                numa = int(open(node_path).read().strip())
        except Exception:      # noqa: BLE001
            cpus = None
        cpus = [c for c in (cpus or allowed) if c in allowed] or allowed
        share = cpus[local % max(1, world)::max(1, world)] if len(cpus) >= 2 * world else cpus   # every GPU the same set: split it
        os.sched_setaffinity(0, share)
        info = {"bound": True, "gpu_cpu_affinity": f"{cpus[0]}-{cpus[-1]} ({len(cpus)} cpus)", "numa_node": numa,
                "rank_cpus": len(share)}
    except Exception as e:      # noqa: BLE001
        info = {"bound": False, "error": str(e)[:120]}
    return info


def h2d_probe(torch, dev, dist, src_pinned, chunk_bytes=1 << 30, reps=8):
    """Bare pinned-host -> device copy bandwidth with every rank copying at the same time (the host-side roof of `e2e`):
    `reps` cudaMemcpyAsync of `chunk_bytes` from different places of this rank's pinned staging buffer, CUDA events."""
    src = src_pinned.view(torch.uint8)
    chunk_bytes = min(chunk_bytes, src.numel())
    dst = torch.empty(chunk_bytes, dtype=torch.uint8, device=dev)
    n_off = max(1, src.numel() // chunk_bytes)
    dst.copy_(src[:chunk_bytes], non_blocking=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        o = (i % n_off) * chunk_bytes
        dst.copy_(src[o:o + chunk_bytes], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = reps * chunk_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    agg, lo = gbs, gbs
    if dist is not None:
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        agg = float(t.item())
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        lo = float(t.item())
    del dst
    return {"aggregate_gbs": agg, "min_rank_gbs": lo, "this_rank_gbs": gbs, "chunk_bytes": chunk_bytes, "reps": reps,
            "what": "concurrent pinned cudaMemcpyAsync H2D on all ranks, no kernels: the host-side roof of e2e"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:      # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:      # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------- CPU baseline
def _cpu_worker(args):
    seed, n = args
    from oracle import modem_v2 as o2, signals as sig
    _, _, x = sig.kat_signal(sig.qpsk_modulate, seed, n // 40 - 64, 20, baud=BAUD, carrier=CARRIER)
    x = np.concatenate([x, np.zeros(max(0, n - len(x)), np.float32)])[:n]
    pcm = np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)       # what the WAV part holds
    t = time.perf_counter()
    xw = pcm.astype(np.float64) / 32768.0                                      # soundfile.read semantics (decoder.py:381)
    raw = o2.qpsk_demodulate(xw, BAUD, CARRIER)
    return time.perf_counter() - t, len(x), len(raw)


def cpu_baseline(seconds_per_rec: int = 30, reps: int = 1):
    """Oracle port of modem.qpsk_demodulate on the host cores: one recording per worker process (the reference
    code is single-threaded), bounded sample, throughput = samples / wall time of the parallel section."""
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, 32))
    n = seconds_per_rec * FS
    jobs = [(5000 + i, n) for i in range(cores * reps)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_worker, [(1, 2 * FS)] * cores)                 # import + warm up each worker
        t = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t
    samples = sum(r[1] for r in res)
    single = float(np.mean([r[1] / r[0] for r in res])) / 1e6
    return {"value": samples / wall / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "single_core_msamples_s": single,
            "sample": f"{len(jobs)} x {seconds_per_rec}-s DQPSK recordings as PCM16 (WAV part payload) -> /32768 -> oracle/modem_v2.qpsk_demodulate "
                      f"(vectorised port of modem.py:189-266; faster than the reference's per-symbol Python loop), "
                      f"one process per core"}


# ----------------------------------------------------------------------------------------- other schemes
def scheme_kernels(eng, torch, batch, n_rec, n_samp, offsets, steps=3):
    """Kernel-only throughput of the v1 streaming demodulators (north_star's named kernels) and the v2 aliases on the
    batch already resident in HBM: whole C-ABI call timed with CUDA events on the engine stream."""
    import ctypes
    from fbdsp import _lib, modem_v1 as g
    import fbdsp
    dev = batch.device
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:      # noqa: BLE001
        pass
    u64p = ctypes.POINTER(ctypes.c_uint64)
    flags = _lib.FB_SAMPLES_ON_DEVICE | _lib.FB_OUT_ON_DEVICE | _lib.FB_ASYNC
    es = torch.cuda.ExternalStream(eng.stream, device=dev)
    ol = torch.zeros(n_rec, dtype=torch.int64, device=dev)
    sy = torch.zeros(n_rec, dtype=torch.int64, device=dev)
    st = torch.zeros(n_rec, dtype=torch.int32, device=dev)
    out = {}

    def timed(name, f):
        for _ in range(2):
            f()
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(es)
        for _ in range(steps):
            f()
        e1.record(es)
        eng.sync()
        ms = e0.elapsed_time(e1) / steps
        byt = n_rec * n_samp * 4 + int(ol.sum().item())
        out[name] = {"ms": round(ms, 3), "gsamples_per_s": round(n_rec * n_samp / ms / 1e6, 1), "GBps": round(byt / ms / 1e6, 1),
                     "frac_of_measured_hbm": round(byt / ms / 1e6 / peak, 4)}

    for name, (p, table) in {"v1_qpsk_9600": g.psk_params(g.V1_QPSK, 9600, 9600.0), "v1_bpsk_9600": g.psk_params(g.V1_BPSK, 9600, 3000.0),
                             "v1_ofdm8_9600": g.ofdm_params(9600, 8), "v1_ofdm4_4800": g.ofdm_params(4800, 4),
                             "v1_psk8_2400": g.psk_params(g.V1_PSK8, 2400, 12000.0),
                             "v1_psk8_38400": g.psk_params(g.V1_PSK8, 38400, 12000.0),
                             "v1_fsk_9600_goertzel_uart": g.fsk_params(9600, 8000.0, 16000.0, 7500, 16500, True)}.items():
        size = (int(eng.lib.fb_v1_out_bound(ctypes.byref(p), n_samp)) + 7) // 4 * 4
        oo = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(size)
        buf = torch.empty(n_rec * size + 16, dtype=torch.uint8, device=dev)

        def f(p=p, table=table, oo=oo, buf=buf):
            rc = eng.lib.fb_v1_demod_batch(eng.handle, ctypes.byref(p), table.ctypes.data, n_rec, batch.data_ptr(), offsets.ctypes.data_as(u64p),
                                           _lib.FB_F32, flags, buf.data_ptr(), oo.ctypes.data_as(u64p), ol.data_ptr(), st.data_ptr())
            _lib.check(eng.lib, eng.handle, rc, "fb_v1_demod_batch")
        timed(name, f)
        del buf
    for name, (baud, carrier, k, n0sps) in {"v2_qpsk_9600_c3000_gui_default": (9600, 3000.0, 1.5, False),
                                            "v2_bpsk_4800_c9600": (4800, 9600.0, 1.0, True)}.items():
        dd = fbdsp.psk_design(float(baud), float(carrier), float(FS), k, n0sps)
        oo = eng.out_bounds(dd, [n_samp] * n_rec)
        buf = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device=dev)
        timed(name, lambda dd=dd, oo=oo, buf=buf: eng.psk_demod_raw(dd, batch.data_ptr(), offsets, _lib.FB_F32, flags, buf.data_ptr(), oo,
                                                                     ol.data_ptr(), sy.data_ptr(), st.data_ptr()))
        del buf
    try:      # v2 FSK (modem.py:298-341) with tones whose Butterworth design is valid (SURVEY 8d config 1 variant)
        from fbdsp import fsk as fskmod
        dd = fskmod.fsk_design(9600, 12000.0, 24000.0, float(FS))
        size = (int(eng.lib.fb_fsk_out_bound(ctypes.byref(dd), n_samp)) + 7) // 4 * 4
        oo = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(size)
        buf = torch.empty(n_rec * size + 16, dtype=torch.uint8, device=dev)

        def f2():
            rc = eng.lib.fb_fsk_demod_batch(eng.handle, ctypes.byref(dd), n_rec, batch.data_ptr(), offsets.ctypes.data_as(u64p), _lib.FB_F32, flags,
                                            buf.data_ptr(), oo.ctypes.data_as(u64p), ol.data_ptr(), sy.data_ptr(), st.data_ptr())
            _lib.check(eng.lib, eng.handle, rc, "fb_fsk_demod_batch")
        timed("v2_fsk_9600_m12000_s24000", f2)
        del buf
    except Exception as e:      # noqa: BLE001  (informational leg: never fail the bench line)
        out["v2_fsk_9600_m12000_s24000"] = {"error": str(e)[:200]}
    return out


# ----------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fbdsp")
    ap.add_argument("--recordings", type=int, default=256)
    ap.add_argument("--seconds", type=int, default=180)
    ap.add_argument("--snr", type=float, default=20.0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-schemes", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-length raw-byte parity check of recording 0 against the oracle")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configs[N-1]: 2 = the headline (default); 3 / 4 / 5 run sharded through fbdsp.shard (tools/bench_configs.py)")
    ap.add_argument("--scale", type=float, default=1.0, help="configs 3-5: fraction of the stated workload size")
    ap.add_argument("--wave-gsamples", type=float, default=4.0, help="configs 3-5: samples resident per wave and GPU, in units of 1e9")
    ap.add_argument("--units", type=int, default=0, help="config 5: total recordings (default 1250 * scale per GPU: weak scaling, 10 000 on 8 GPUs)")
    ap.add_argument("--no-audit", action="store_true", help="configs 4-5: skip the oracle audit of the outputs")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner, torchrun) also write to fd 1, so keep the
    # real stdout aside for the result line and send everything else to stderr
    sys.stdout.flush()
    result_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    binding = bind_rank_to_gpu_cpus(local, world) if args.impl != "reference" and not os.environ.get("FB_BENCH_NO_BIND") else None
    workload = f"qpsk{BAUD}_c{int(CARRIER)}_{args.recordings}x{args.seconds}s_f32"
    config = {"workload": workload, "scheme": "DQPSK (modem.qpsk_demodulate)", "baud": BAUD, "carrier_hz": CARRIER,
              "fs_hz": FS, "recordings_per_gpu": args.recordings, "seconds_per_recording": args.seconds,
              "snr_db": args.snr, "sample_dtype": "float32 resident in HBM for `value`/roofline; PCM16 host buffers (WAV payload) for `e2e`", "l2": "inputs (GBs) larger than L2, no flush needed",
              "sharding": "independent recordings per rank, no collective in the data path"}

    if args.config != 2 and args.impl != "reference":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        return bench_configs.run(args, args.config, rank, world, local, result_out, ClockSampler)

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, args.steps)
        for _ in range(max(0, args.warmup)):
            cpu_baseline(5)
        vals = []
        t0 = time.perf_counter()
        for _ in range(steps):
            cb = cpu_baseline(20)
            vals.append(cb["value"])
        ms = (time.perf_counter() - t0) / steps * 1e3
        v = float(np.mean(vals))
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": dict(config, reference_sample=(f"each step demodulates {cb['cores']} x 20-s recordings of this workload (one per "
                                                                   "host core, the same DQPSK parameter set and SNR), NOT the 256 x 180-s batch: throughput "
                                                                   "is linear in the record length (SURVEY 6), the value is samples / wall time of that sample"),
                                        frame_parse="not included (the reference arm times the demodulator only; the product arm parses frames too)"),
                          "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              file=result_out, flush=True)
        return

    import torch
    import fbdsp
    from fbdsp import _lib
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    eng = fbdsp.Engine(local)
    d = fbdsp.psk_design(float(BAUD), float(CARRIER), float(FS), 1.5, False)
    n_rec, n_samp = args.recordings, args.seconds * FS
    batch, payload_bytes_in = synth_batch_device(torch, dev, n_rec, n_samp, args.snr, rank)
    torch.cuda.synchronize()
    offsets = (np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(n_samp))
    out_offsets = eng.out_bounds(d, [n_samp] * n_rec)
    out_dev = torch.empty(int(out_offsets[-1]) + 16, dtype=torch.uint8, device=dev)
    out_len = torch.zeros(n_rec, dtype=torch.int64, device=dev)
    sync_idx = torch.zeros(n_rec, dtype=torch.int64, device=dev)
    status = torch.zeros(n_rec, dtype=torch.int32, device=dev)
    flags = _lib.FB_SAMPLES_ON_DEVICE | _lib.FB_OUT_ON_DEVICE | _lib.FB_ASYNC
    es = torch.cuda.ExternalStream(eng.stream, device=dev)

    import ctypes
    MAXF = 4
    frames_dev = torch.zeros(n_rec * MAXF * ctypes.sizeof(_lib.fb_frame), dtype=torch.uint8, device=dev)
    nfr_dev = torch.zeros(n_rec, dtype=torch.int32, device=dev)
    pbytes_dev = torch.zeros(n_rec, dtype=torch.int64, device=dev)
    oo_p = out_offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))

    def step():
        # demodulate (edge + main + sync search + byte pack), then parse FBPC frames + CRC32 -- all on the device
        eng.psk_demod_raw(d, batch.data_ptr(), offsets, _lib.FB_F32, flags, out_dev.data_ptr(), out_offsets,
                          out_len.data_ptr(), sync_idx.data_ptr(), status.data_ptr())
        rc = eng.lib.fb_parse_frames_batch(eng.handle, n_rec, out_dev.data_ptr(), oo_p, out_len.data_ptr(), MAXF,
                                           frames_dev.data_ptr(), nfr_dev.data_ptr(), pbytes_dev.data_ptr(), flags)
        _lib.check(eng.lib, eng.handle, rc, "fb_parse_frames_batch")

    def barrier():
        if dist is not None:
            dist.barrier()
        eng.sync()
        torch.cuda.synchronize()

    eng.lib.fb_set_profiling(eng.handle, 1)
    clocks = ClockSampler(local)                      # nvidia-smi needs ~0.2 s to deliver its first sample: start it before
    clocks.start()                                    # the warm-up so the (short) timed region is covered
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = eng.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kernel_ms = []
    barrier()
    ev[0].record(es)
    for _ in range(args.steps):
        step()
    ev[1].record(es)
    barrier()
    total_ms = ev[0].elapsed_time(ev[1])
    launches = eng.kernel_launches - launches0
    # dominant-kernel duration, CUDA events on its own stream: re-run K profiled steps one at a time
    kernel_ms = []
    for _ in range(args.steps):
        step()
        kernel_ms.append(float(eng.lib.fb_kernel_ms(eng.handle)))
    barrier()
    t_load = time.perf_counter()
    while len(clocks.lines) < 5 and time.perf_counter() - t_load < 3.0:      # keep the GPU under the same load until the
        step()                                                               # sampler has a few readings (untimed)
        eng.sync()
    clk = clocks.stop()
    print(f"[bench] rank {rank} (cuda:{local}): {total_ms / args.steps:.3f} ms per step, interior kernel {np.mean(kernel_ms):.3f} ms", file=sys.stderr, flush=True)
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    samples_per_step = n_rec * n_samp * world
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6

    # results of the last step: raw bytes, frames (host parse + CRC32, untimed), parity spot check vs the oracle
    ol = out_len.cpu().numpy()
    raw_bytes = int(ol.sum())
    out_host = out_dev.cpu().numpy()
    from oracle.frames import parse_fbp_stream
    payload_ok = int(pbytes_dev.sum().item())            # CRC-valid payload bytes found by the device parser (timed)
    o0 = int(out_offsets[0])                              # spot check of recording 0 against the oracle parser
    want0 = sum(len(fr["data"]) for fr in parse_fbp_stream(out_host[o0:o0 + int(ol[0])].tobytes()))
    assert want0 == int(pbytes_dev[0].item()), "device frame parser disagrees with the oracle"
    # full-length raw-byte parity of recording 0 of this rank (17.28 M samples at the default size): every byte the reference's
    # qpsk_demodulate returns -- also before the sync word and after the frame -- and the sync index, against the oracle (untimed)
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import modem_v2 as o2
        t_p = time.perf_counter()
        st0 = o2.qpsk_stages(batch[:n_samp].cpu().numpy(), BAUD, CARRIER)
        got0 = out_host[o0:o0 + int(ol[0])].tobytes()
        parity = {"recording": 0, "samples": int(n_samp), "raw_bytes": len(st0["raw"]), "raw_bytes_equal": got0 == st0["raw"],
                  "sync_idx_equal": int(sync_idx[0].item()) == int(st0["sync"]), "oracle_s": round(time.perf_counter() - t_p, 2),
                  "against": "oracle/modem_v2.qpsk_stages (restatement of modem.py:189-266, pinned by tests/golden)"}
        if not parity["raw_bytes_equal"]:                 # reported, not fatal: the line still carries the measurement
            a_, b_ = np.frombuffer(got0, np.uint8), np.frombuffer(st0["raw"], np.uint8)
            m_ = min(len(a_), len(b_))
            parity["differing_bytes"] = int(np.count_nonzero(a_[:m_] != b_[:m_])) + abs(len(a_) - len(b_))
            print(f"[bench] WARNING: full-length parity of recording 0 failed: {parity}", file=sys.stderr, flush=True)
    if dist is not None:
        t = torch.tensor([raw_bytes, payload_ok], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        raw_all, payload_all = float(t[0]), float(t[1])
    else:
        raw_all, payload_all = float(raw_bytes), float(payload_ok)

    # ---- e2e: host-buffer C-ABI call, pinned samples, H2D + kernels + D2H of the raw bytes inside the timed region
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError('skipped (--no-e2e)')
        host = torch.empty(batch.numel(), dtype=torch.float32, pin_memory=True)
        host.copy_(batch)
        out_h = torch.empty(int(out_offsets[-1]) + 16, dtype=torch.uint8, pin_memory=True)
        ol_h = torch.zeros(n_rec, dtype=torch.int64, pin_memory=True)
        sy_h = torch.zeros(n_rec, dtype=torch.int64, pin_memory=True)
        st_h = torch.zeros(n_rec, dtype=torch.int32, pin_memory=True)

        fr_h = (_lib.fb_frame * (n_rec * MAXF))()
        nf_h = np.zeros(n_rec, dtype=np.int32)
        pb_h = np.zeros(n_rec, dtype=np.uint64)

        def step_e2e():
            eng.psk_demod_raw(d, host.data_ptr(), offsets, _lib.FB_F32, 0, out_h.data_ptr(), out_offsets,
                              ol_h.data_ptr(), sy_h.data_ptr(), st_h.data_ptr())
            rc = eng.lib.fb_parse_frames_batch(eng.handle, n_rec, out_h.data_ptr(), oo_p, ol_h.data_ptr(), MAXF,
                                               ctypes.addressof(fr_h), nf_h.ctypes.data, pb_h.ctypes.data, 0)
            _lib.check(eng.lib, eng.handle, rc, "fb_parse_frames_batch")
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        assert int(ol_h.sum()) == raw_bytes and int(pb_h.sum()) == payload_ok, "host-buffer path disagrees with the device-resident path"
        e2e = {"value": samples_per_step / e2e_s / 1e6, "unit": "Msamples/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(batch.numel() * 4), "d2h_bytes_per_step": int(out_offsets[-1]) + n_rec * 20,
               "steps": args.e2e_steps, "api": "fb_psk_demod_batch + fb_parse_frames_batch with host pointers (fbdsp.Engine)"}
    except RuntimeError as e:       # pinned allocation can fail on a small host
        e2e = {"value": None, "unit": "Msamples/s", "error": str(e)[:200]}

    # ---- e2e on PCM16 host buffers (what a WAV part holds; decoder.decode_wav_file path: device scales by 1/32768) ----
    e2e_pcm16 = None
    if not args.no_e2e and e2e and e2e.get("value"):
        try:
            del host
            pcm = torch.empty(batch.numel(), dtype=torch.int16, pin_memory=True)
            pcm.copy_((batch * 32767.0).clamp_(-32768, 32767).to(torch.int16))

            def step_pcm():
                eng.psk_demod_raw(d, pcm.data_ptr(), offsets, _lib.FB_S16, 0, out_h.data_ptr(), out_offsets,
                                  ol_h.data_ptr(), sy_h.data_ptr(), st_h.data_ptr())
                rc = eng.lib.fb_parse_frames_batch(eng.handle, n_rec, out_h.data_ptr(), oo_p, ol_h.data_ptr(), MAXF,
                                                   ctypes.addressof(fr_h), nf_h.ctypes.data, pb_h.ctypes.data, 0)
                _lib.check(eng.lib, eng.handle, rc, "fb_parse_frames_batch")
            probe = h2d_probe(torch, dev, dist, pcm)
            step_pcm()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                step_pcm()
            barrier()
            s_pcm = (time.perf_counter() - t0) / args.e2e_steps
            if dist is not None:
                t = torch.tensor([s_pcm], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                s_pcm = float(t.item())
            e2e_pcm16 = {"value": samples_per_step / s_pcm / 1e6, "unit": "Msamples/s", "ms_per_step": s_pcm * 1e3,
                         "h2d_bytes_per_step": int(batch.numel() * 2), "d2h_bytes_per_step": int(out_offsets[-1]) + n_rec * 20,
                         "steps": args.e2e_steps, "payload_bytes_valid": int(pb_h.sum()), "host_format": "PCM16",
                         "h2d_probe": probe, "h2d_gbs_achieved_all_ranks": batch.numel() * 2 * world / s_pcm / 1e9, "cpu_binding": binding,
                         "api": "fb_psk_demod_batch (FB_S16) + fb_parse_frames_batch with host pointers (fbdsp.Engine)",
                         "note": "host buffers hold the WAV parts' PCM16 payload (what decode_wav_file reads; the device scales by "
                                 "1/32768 exactly like soundfile); e2e_f32 is the same call on float32 host buffers"}
            del pcm
        except RuntimeError as e:
            e2e_pcm16 = {"value": None, "error": str(e)[:200]}

    # ---- other schemes' dominant kernels on the same device-resident batch (rank 0, informational) -------------------
    # every rank measures them on its own batch at the same time; rank 0 reports its own figures plus the sum over all GPUs
    schemes = None
    if not args.no_schemes:
        if dist is not None:
            dist.barrier()
        schemes = scheme_kernels(eng, torch, batch, n_rec, n_samp, offsets)
        if dist is not None:
            names = sorted(k for k, v in schemes.items() if isinstance(v, dict) and "gsamples_per_s" in v)
            t = torch.tensor([schemes[k]["gsamples_per_s"] for k in names], dtype=torch.float64, device=dev)
            ok = torch.tensor([float(len(names))], dtype=torch.float64, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == len(names):             # (a scheme that failed on some rank would change the list)
                dist.all_reduce(t)
                for k, v in zip(names, t.tolist()):
                    schemes[k]["gsamples_per_s_all_gpus"] = round(v, 1)
                    schemes[k]["n_gpus"] = world

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:      # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k_ms = float(np.mean([k for k in kernel_ms if k and k > 0])) if kernel_ms else float("nan")
    alg_bytes = n_rec * n_samp * 4 + raw_bytes
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    mma = not os.environ.get("FB_PSK_NO_MMA")
    prof = "r02_psk_mma_ncu.txt" if mma else "r01_psk_main_ncu.txt"
    roofline = {"kernel": "psk_mma_kernel<float, Sched10>" if mma else "psk_main_kernel<float>", "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "traffic_source": None, "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms / ms_per_step, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
                "note": ("tensor-pipe kernel (mma.sync HMMA on fp16 hi/lo split operands, TMA-staged samples, DESIGN.md 4.1): bound by "
                         "shared-memory wavefronts and the latency of its three warp roles, not by HBM; DRAM traffic equals the "
                         "algorithmic bytes") if mma else
                        ("fp32-FMA co-limited, not HBM-bound: 32 (FIR, FFMA2) + 8 (slow-pole features) + ~8 FMA per sample, "
                         "so >= 6.1 ms at the measured 124 FMA/clk/SM (DESIGN.md 4)")}
    try:      # dram__bytes_read + dram__bytes_write per sample of the same kernel, from the committed `ncu --set full` capture
        for ln in open(os.path.join(ROOT, "profiles", prof)):
            if ln.startswith("dram_bytes_per_sample"):
                roofline["traffic"] = float(ln.split()[1]) * n_rec * n_samp
                roofline["traffic_source"] = (f"static profile: profiles/{prof} (one ncu --set full launch on 32 recordings), bytes per "
                                              "sample x the samples of this launch; not measured in this run")
    except Exception:      # noqa: BLE001
        pass
    if getattr(bind_rank_to_gpu_cpus, "allowed", None):
        os.sched_setaffinity(0, bind_rank_to_gpu_cpus.allowed)
    cb = None if args.no_cpu else cpu_baseline(30)
    line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "dtype_note": ("float32 samples; interior symbols: fp16 hi/lo split operands (22 significant bits) on HMMA with fp32 accumulation, "
                           "slow-pole scan / slicer in fp32; record edges in float64; decisions bit-equal to the float64 reference on every "
                           "golden case and on recording 0 of this run (parity_spot_check)"),
            "data": "synthetic", "config": dict(config, frame_parse="device (parse_frames_kernel: FBPC scan + header checks + CRC32), inside the timed step"),
            "raw_MB_per_s": raw_all / (ms_per_step * 1e-3) / 1e6, "payload_MB_per_s": payload_all / (ms_per_step * 1e-3) / 1e6,
            "payload_bytes_valid": payload_all, "payload_bytes_sent_rank0": payload_bytes_in,
            "gsamples_per_s_per_gpu": value / 1e3 / world, "gpu_launches": int(launches), "clocks": clk,
            "e2e": (e2e_pcm16 if (e2e_pcm16 and e2e_pcm16.get("value")) else e2e), "roofline": roofline, "cpu_baseline": cb,
            "e2e_f32": e2e, "schemes": schemes, "parity_spot_check": parity}
    print(json.dumps(line), file=result_out, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
