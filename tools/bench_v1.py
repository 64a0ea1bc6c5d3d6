#!/usr/bin/env python
"""Device-resident throughput of the v1 streaming kernels (north_star's named kernels) on synthetic noise:
python tools/bench_v1.py [--recordings 256] [--seconds 180].  Prints one JSON line per scheme."""
import argparse, ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-modem-radio_b200")]
import torch
import fbdsp
from fbdsp import _lib, modem_v1 as g

ap = argparse.ArgumentParser()
ap.add_argument("--recordings", type=int, default=256)
ap.add_argument("--seconds", type=int, default=180)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--only", default="")
args = ap.parse_args()
dev = torch.device("cuda", 0)
eng = fbdsp.Engine(0)
n_rec, n = args.recordings, args.seconds * 96000
tdt = {"f32": torch.float32, "s16": torch.int16}[args.dtype]
if args.dtype == "f32":
    batch = torch.randn(n_rec * n, device=dev, dtype=torch.float32) * 0.3
else:
    batch = (torch.randn(n_rec * n, device=dev, dtype=torch.float32) * 8000).to(torch.int16)
offsets = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(n)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
u64p = ctypes.POINTER(ctypes.c_uint64)
flags = _lib.FB_SAMPLES_ON_DEVICE | _lib.FB_OUT_ON_DEVICE | _lib.FB_ASYNC
esz = batch.element_size()
for name, (p, table) in {"v1_qpsk_9600": g.psk_params(g.V1_QPSK, 9600, 9600.0), "v1_bpsk_9600": g.psk_params(g.V1_BPSK, 9600, 3000.0),
                         "v1_psk8_38400": g.psk_params(g.V1_PSK8, 38400, 12000.0), "v1_qpsk_1200": g.psk_params(g.V1_QPSK, 1200, 3000.0),
                         "v1_ofdm8_9600": g.ofdm_params(9600, 8), "v1_ofdm4_4800": g.ofdm_params(4800, 4)}.items():
    if args.only and name not in args.only.split(','):
        continue
    size = (int(eng.lib.fb_v1_out_bound(ctypes.byref(p), n)) + 7) // 4 * 4
    out_offsets = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(size)
    out = torch.empty(n_rec * size + 16, dtype=torch.uint8, device=dev)
    out_len = torch.zeros(n_rec, dtype=torch.int64, device=dev)
    status = torch.zeros(n_rec, dtype=torch.int32, device=dev)
    eng.lib.fb_set_profiling(eng.handle, 1)
    def step():
        rc = eng.lib.fb_v1_demod_batch(eng.handle, ctypes.byref(p), table.ctypes.data, n_rec, batch.data_ptr(), offsets.ctypes.data_as(u64p),
                                       _lib.FB_F32 if args.dtype == "f32" else _lib.FB_S16, flags, out.data_ptr(), out_offsets.ctypes.data_as(u64p),
                                       out_len.data_ptr(), status.data_ptr())
        _lib.check(eng.lib, eng.handle, rc, "fb_v1_demod_batch")
    for _ in range(3):
        step()
    eng.sync()
    ms = []
    for _ in range(args.steps):
        step()
        ms.append(float(eng.lib.fb_kernel_ms(eng.handle)))
    eng.sync()
    k = float(np.mean(ms))
    byt = n_rec * n * esz + int(out_len.sum().item())
    print(json.dumps({"scheme": name, "dtype": args.dtype, "kernel_ms": k, "gsamples_per_s": n_rec * n / k / 1e6, "GBps": byt / k / 1e6,
                      "frac_of_measured_hbm": byt / k / 1e6 / peak, "raw_bytes": int(out_len.sum().item())}))
