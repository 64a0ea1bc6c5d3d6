#!/usr/bin/env python
"""Where a bench step's time goes: psk call alone, frame-parse call alone, both (CUDA events on the engine's stream) and
the host-side cost of issuing each call.  python tools/step_breakdown.py [--recordings 256]"""
import argparse, ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-modem-radio_b200")]
import torch, fbdsp
from fbdsp import _lib
ap = argparse.ArgumentParser(); ap.add_argument("--recordings", type=int, default=256); ap.add_argument("--dtype", default="f32"); ap.add_argument("--carrier", type=float, default=9600.0)
args = ap.parse_args()
dev = torch.device("cuda", 0); eng = fbdsp.Engine(0)
d = fbdsp.psk_design(9600.0, args.carrier, 96000.0, 1.5, False)
n_rec, n = args.recordings, 180 * 96000
if args.dtype == "f32":
    batch = torch.randn(n_rec * n, device=dev, dtype=torch.float32) * 0.3; dt = _lib.FB_F32
else:
    batch = (torch.randn(n_rec * n, device=dev, dtype=torch.float32) * 8000).to(torch.int16); dt = _lib.FB_S16
offsets = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(n)
oo = eng.out_bounds(d, [n] * n_rec)
out = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device=dev)
ol = torch.zeros(n_rec, dtype=torch.int64, device=dev); sy = torch.zeros(n_rec, dtype=torch.int64, device=dev); st = torch.zeros(n_rec, dtype=torch.int32, device=dev)
fr = torch.zeros(n_rec * 4 * ctypes.sizeof(_lib.fb_frame), dtype=torch.uint8, device=dev); nf = torch.zeros(n_rec, dtype=torch.int32, device=dev); pb = torch.zeros(n_rec, dtype=torch.int64, device=dev)
flags = _lib.FB_SAMPLES_ON_DEVICE | _lib.FB_OUT_ON_DEVICE | _lib.FB_ASYNC
oo_p = oo.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
es = torch.cuda.ExternalStream(eng.stream, device=dev)
def psk(): eng.psk_demod_raw(d, batch.data_ptr(), offsets, dt, flags, out.data_ptr(), oo, ol.data_ptr(), sy.data_ptr(), st.data_ptr())
def parse(): _lib.check(eng.lib, eng.handle, eng.lib.fb_parse_frames_batch(eng.handle, n_rec, out.data_ptr(), oo_p, ol.data_ptr(), 4, fr.data_ptr(), nf.data_ptr(), pb.data_ptr(), flags), "parse")
def both(): psk(); parse()
res = {}
for name, f in (("psk", psk), ("parse", parse), ("both", both)):
    for _ in range(3): f()
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(es); t0 = time.perf_counter()
    for _ in range(10): f()
    host = (time.perf_counter() - t0) / 10; e1.record(es); eng.sync()
    res[name] = {"gpu_ms": e0.elapsed_time(e1) / 10, "host_issue_ms": host * 1e3}
print(json.dumps(res))
