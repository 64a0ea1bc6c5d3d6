#!/usr/bin/env python
"""Phase timeline of psk_main_kernel (experiments): needs a -DPM_TRACE build of the library,
    NVCC_EXTRA=-DPM_TRACE FB_OUT=lib/libfbdsp_trace.so bash audio-modem-radio_b200/build.sh
    FBDSP_LIB=audio-modem-radio_b200/lib/libfbdsp_trace.so python tools/pm_trace.py
Prints, over the CTAs of the steady state, the median time (us) thread 0 spends in each phase, the CTA residency and
how many CTAs of an SM overlap."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-modem-radio_b200")]
import torch, fbdsp
from fbdsp import _lib
dev = torch.device("cuda", 0); eng = fbdsp.Engine(0)
d = fbdsp.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
n_rec, n = int(os.environ.get("RECS", 8)), 180 * 96000
batch = torch.randn(n_rec * n, device=dev, dtype=torch.float32) * 0.3
offsets = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(n)
oo = eng.out_bounds(d, [n] * n_rec)
out = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device=dev)
ol = torch.zeros(n_rec, dtype=torch.int64, device=dev); sy = torch.zeros(n_rec, dtype=torch.int64, device=dev); st = torch.zeros(n_rec, dtype=torch.int32, device=dev)
flags = _lib.FB_SAMPLES_ON_DEVICE | _lib.FB_OUT_ON_DEVICE | _lib.FB_ASYNC
for _ in range(3):
    eng.psk_demod_raw(d, batch.data_ptr(), offsets, _lib.FB_F32, flags, out.data_ptr(), oo, ol.data_ptr(), sy.data_ptr(), st.data_ptr())
eng.sync()
N = 16384
t = np.zeros((N, 12), dtype=np.uint64); sm = np.zeros(N, dtype=np.uint32)
got = eng.lib.fb_debug_pm_trace(t.ctypes.data, sm.ctypes.data)
assert got == N, got
n_tiles = min(N, n_rec * (n // 10 // 2016))
t = t[1000:n_tiles - 1000].astype(np.int64); sm = sm[1000:n_tiles - 1000]
names = ["stage", "boundary sums", "barrier", "slow poles", "FIR", "slicer+store"]
res = {names[i]: round(float(np.median(t[:, i + 1] - t[:, i])) / 1e3, 2) for i in range(6)}
res["residency_us"] = round(float(np.median(t[:, 6] - t[:, 0])) / 1e3, 2)
res["p90_residency_us"] = round(float(np.percentile(t[:, 6] - t[:, 0], 90)) / 1e3, 2)
span = (t[:, 6].max() - t[:, 0].min()) / 1e3
if t[:, 7].min() > 0:                                     # finer marks inside the slow-pole phase
    res["slow: features"] = round(float(np.median(t[:, 7] - t[:, 3])) / 1e3, 2)
    res["slow: fold + warp scan"] = round(float(np.median(t[:, 8] - t[:, 7])) / 1e3, 2)
    res["slow: carry (2 barriers)"] = round(float(np.median(t[:, 9] - t[:, 8])) / 1e3, 2)
    res["slow: column recursion"] = round(float(np.median(t[:, 4] - t[:, 9])) / 1e3, 2)
res["tiles"] = int(len(t)); res["span_us"] = round(float(span), 1)
res["tiles_per_sm_slot_us"] = round(float(span) * 148 * 2 / len(t), 2)
print(json.dumps(res))
