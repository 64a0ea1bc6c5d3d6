#!/usr/bin/env python
"""Throughput of the batch modulators (device-resident output): python tools/bench_modulate.py [--payloads 64] [--bytes 431000]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-modem-radio_b200")]
import torch, fbdsp
from fbdsp import modulate as m
ap = argparse.ArgumentParser(); ap.add_argument("--payloads", type=int, default=64); ap.add_argument("--bytes", type=int, default=431000)
a = ap.parse_args()
dev = torch.device("cuda", 0); eng = fbdsp.Engine(0)
rng = np.random.default_rng(0)
data = torch.from_numpy(rng.integers(0, 256, a.payloads * a.bytes, dtype=np.uint8)).to(dev)
doff = np.arange(a.payloads + 1, dtype=np.uint64) * np.uint64(a.bytes)
es = torch.cuda.ExternalStream(eng.stream, device=dev)
for name, prm in (("dqpsk_9600", m.psk_mod_params(m.FB_MOD_DQPSK, 9600, 9600.0, 96000)), ("dbpsk_4800", m.psk_mod_params(m.FB_MOD_DBPSK, 4800, 9600.0, 96000)),
                  ("cpfsk_9600", m.fsk_mod_params(9600, 12000.0, 24000.0, 96000))):
    n = m.out_samples(prm[0], a.bytes, eng)
    ooff = np.arange(a.payloads + 1, dtype=np.uint64) * np.uint64(n)
    out = torch.empty(a.payloads * n, dtype=torch.float32, device=dev)
    for _ in range(2):
        m.modulate_batch_device(*prm, data.data_ptr(), doff, out.data_ptr(), ooff, eng)
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(es)
    for _ in range(3):
        m.modulate_batch_device(*prm, data.data_ptr(), doff, out.data_ptr(), ooff, eng)
    e1.record(es); eng.sync()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"modulator": name, "payloads": a.payloads, "samples": a.payloads * n, "ms": round(ms, 2), "gsamples_per_s": round(a.payloads * n / ms / 1e6, 1)}), flush=True)
