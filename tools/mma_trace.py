#!/usr/bin/env python
"""Phase timeline of psk_mma_kernel (FB_MMA_TRACE=1): median clocks per tile between the stamps thread 0 of the MMA warps and
thread 0 of the loader warps record.  usage: FB_MMA_TRACE=1 python tools/mma_trace.py [recordings]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-modem-radio_b200")]
os.environ["FB_MMA_TRACE"] = "1"
import torch, fbdsp
from fbdsp import _lib
n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 180 * 96000
eng = fbdsp.Engine(0)
d = fbdsp.psk_design(9600.0, 9600.0, 96000.0, 1.5, False)
x = torch.randn(n_rec * n, device="cuda", dtype=torch.float32) * 0.3
off = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(n)
oo = eng.out_bounds(d, [n] * n_rec)
out = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device="cuda")
ol = torch.zeros(n_rec, dtype=torch.int64, device="cuda"); sy = torch.zeros_like(ol); st = torch.zeros(n_rec, dtype=torch.int32, device="cuda")
fl = _lib.FB_SAMPLES_ON_DEVICE | _lib.FB_OUT_ON_DEVICE | _lib.FB_ASYNC
eng.lib.fb_set_profiling(eng.handle, 1)
for _ in range(4):
    eng.psk_demod_raw(d, x.data_ptr(), off, _lib.FB_F32, fl, out.data_ptr(), oo, ol.data_ptr(), sy.data_ptr(), st.data_ptr())
eng.sync()
print("kernel_ms (events around the launch):", eng.lib.fb_kernel_ms(eng.handle))
buf = np.zeros((148, 48, 16), dtype=np.int64)
eng.lib.fb_debug_mma_trace.restype = ctypes.c_int
eng.lib.fb_debug_mma_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
k = eng.lib.fb_debug_mma_trace(eng.handle, buf.ctypes.data, 148)
t = buf[:, 4:44, :].astype(np.float64)            # steady-state tiles
def med(i, j):
    return np.median(t[:, :, j] - t[:, :, i])
def period(i):
    return np.median(t[:, 1:, i] - t[:, :-1, i])
print(f"MMA warps : wait full {med(0, 1):7.0f} | MMA phase (incl. wait for the handoff buffer) {med(1, 2):7.0f} | period {period(0):7.0f} clk")
print(f"post warps: wait ufull {med(3, 4):7.0f} | scan + symbols + slicer {med(4, 5):7.0f} | period {period(3):7.0f} clk")
print(f"MMA detail: full -> handoff buffer free {med(1, 14):7.0f} | first m-tile {med(14, 15):7.0f} | second m-tile + arrive {med(15, 2):7.0f}")
print(f"post detail: tile/non-chained + features + warp scan {med(4, 6):7.0f} | barrier {med(6, 7):7.0f} | carries + symbols {med(7, 8):7.0f} | yex barrier {med(8, 9):7.0f} | slicer + store {med(9, 5):7.0f}")
print(f"loaders   : wait for the first box {med(10, 11):7.0f} | convert + store (incl. waits for the other boxes) {med(11, 12):7.0f} | period {period(10):7.0f} clk")

per = t[:, 1:, 0] - t[:, :-1, 0]
print(f"MMA period: mean {per.mean():.0f} p90 {np.percentile(per, 90):.0f} max {per.max():.0f}; first stamp spread across CTAs {(buf[:, 0, 0].max() - buf[:, 0, 0].min())} clk (different SM clocks: indicative only)")
tot = buf[:, 47, 0] - buf[:, 0, 0]
print(f"47 tiles took: median {np.median(tot):.0f} min {tot.min()} max {tot.max()} clk")

ns = buf[:, 47, 13] - buf[:, 0, 13]
print(f"47 tiles took (globaltimer): median {np.median(ns):.0f} ns -> SM clock {np.median(tot) / np.median(ns) * 1000:.0f} MHz; CTA start spread {(buf[:, 0, 13].max() - buf[:, 0, 13].min()) / 1e3:.1f} us; first tile .. tile 47 end spread {(buf[:, 47, 13].max() - buf[:, 0, 13].min()) / 1e3:.1f} us")

if os.environ.get("FB_TRACE_DUMP"):
    cta = int(os.environ["FB_TRACE_DUMP"])
    base = buf[cta, 8, 0]
    names = {0: "M.top", 1: "M.full", 13: "M.rangechk", 14: "M.uempty", 15: "M.mt0", 2: "M.done", 3: "P.top", 4: "P.ufull", 6: "P.scan", 7: "P.bar", 8: "P.sym", 9: "P.yex", 5: "P.done", 10: "L.top", 11: "L.issued", 12: "L.conv"}
    ev = []
    for it in range(8, 14):
        for slot, nm in names.items():
            ev.append((int(buf[cta, it, slot] - base), f"{nm}[{it}]"))
    for tm, nm in sorted(ev):
        print(f"{tm:8d} {nm}")
