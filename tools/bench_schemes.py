#!/usr/bin/env python
"""Device-resident throughput of every demodulator family on synthetic noise (timing only; parity lives in tests/):
python tools/bench_schemes.py [--recordings 64] [--seconds 180] [--only name,name]  -> one JSON line per scheme.
`ms` is the whole C-ABI call (all kernels of the scheme, CUDA events on the engine stream); `kernel_ms` the dominant
kernel where the library records it."""
import argparse, ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-modem-radio_b200")]
import torch
import fbdsp
from fbdsp import _lib, modem_v1 as g, fsk as fskmod


def run(args):
    dev = torch.device("cuda", 0)
    eng = fbdsp.Engine(0)
    n_rec, n = args.recordings, args.seconds * 96000
    if args.dtype == "f32":
        batch = torch.randn(n_rec * n, device=dev, dtype=torch.float32) * 0.3; dt = _lib.FB_F32
    else:
        batch = (torch.randn(n_rec * n, device=dev, dtype=torch.float32) * 8000).to(torch.int16); dt = _lib.FB_S16
    esz = batch.element_size()
    offsets = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(n)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:      # noqa: BLE001
        pass
    u64p = ctypes.POINTER(ctypes.c_uint64)
    flags = _lib.FB_SAMPLES_ON_DEVICE | _lib.FB_OUT_ON_DEVICE | _lib.FB_ASYNC
    es = torch.cuda.ExternalStream(eng.stream, device=dev)
    ol = torch.zeros(n_rec, dtype=torch.int64, device=dev); sy = torch.zeros(n_rec, dtype=torch.int64, device=dev)
    st = torch.zeros(n_rec, dtype=torch.int32, device=dev)
    schemes = {}

    def add_psk(name, baud, carrier, k, n0sps):
        d = fbdsp.psk_design(float(baud), float(carrier), 96000.0, k, n0sps)
        oo = eng.out_bounds(d, [n] * n_rec)
        out = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device=dev)
        schemes[name] = lambda: eng.psk_demod_raw(d, batch.data_ptr(), offsets, dt, flags, out.data_ptr(), oo, ol.data_ptr(), sy.data_ptr(), st.data_ptr())

    def add_v1(name, p, table):
        size = (int(eng.lib.fb_v1_out_bound(ctypes.byref(p), n)) + 7) // 4 * 4
        oo = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(size)
        out = torch.empty(n_rec * size + 16, dtype=torch.uint8, device=dev)
        def f():
            rc = eng.lib.fb_v1_demod_batch(eng.handle, ctypes.byref(p), table.ctypes.data, n_rec, batch.data_ptr(), offsets.ctypes.data_as(u64p),
                                           dt, flags, out.data_ptr(), oo.ctypes.data_as(u64p), ol.data_ptr(), st.data_ptr())
            _lib.check(eng.lib, eng.handle, rc, "fb_v1_demod_batch")
        schemes[name] = f

    def add_fsk2(name, baud, mark, space):
        d = fskmod.fsk_design(baud, mark, space, 96000.0)
        size = (int(eng.lib.fb_fsk_out_bound(ctypes.byref(d), n)) + 7) // 4 * 4
        oo = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(size)
        out = torch.empty(n_rec * size + 16, dtype=torch.uint8, device=dev)
        def f():
            rc = eng.lib.fb_fsk_demod_batch(eng.handle, ctypes.byref(d), n_rec, batch.data_ptr(), offsets.ctypes.data_as(u64p), dt, flags,
                                            out.data_ptr(), oo.ctypes.data_as(u64p), ol.data_ptr(), sy.data_ptr(), st.data_ptr())
            _lib.check(eng.lib, eng.handle, rc, "fb_fsk_demod_batch")
        schemes[name] = f

    add_psk("v2_qpsk_9600_c9600", 9600, 9600.0, 1.5, False)
    add_psk("v2_qpsk_9600_c3000", 9600, 3000.0, 1.5, False)
    add_psk("v2_bpsk_4800_c9600", 4800, 9600.0, 1.0, True)
    add_psk("v2_qpsk_1200_c3000", 1200, 3000.0, 1.5, False)
    add_psk("v2_psk8_38400_c12000", 38400, 12000.0, 1.5, False)
    add_fsk2("v2_fsk_9600_m12000_s24000", 9600, 12000.0, 24000.0)
    add_v1("v1_qpsk_9600", *g.psk_params(g.V1_QPSK, 9600, 9600.0))
    add_v1("v1_bpsk_9600", *g.psk_params(g.V1_BPSK, 9600, 3000.0))
    add_v1("v1_psk8_2400", *g.psk_params(g.V1_PSK8, 2400, 12000.0))
    add_v1("v1_psk8_38400", *g.psk_params(g.V1_PSK8, 38400, 12000.0))
    add_v1("v1_qpsk_1200", *g.psk_params(g.V1_QPSK, 1200, 3000.0))
    add_v1("v1_ofdm8_9600", *g.ofdm_params(9600, 8))
    add_v1("v1_ofdm4_4800", *g.ofdm_params(4800, 4))
    add_v1("v1_fsk_9600_goertzel_uart", *g.fsk_params(9600, 8000.0, 16000.0, 7500, 16500, True))
    add_v1("v1_fsk_1200_goertzel_uart", *g.fsk_params(1200, 1200.0, 2200.0, 700, 2700, True))
    add_v1("v1_fskhs_19200_goertzel", *g.fsk_params(19200, 12000.0, 18000.0, 8000, 22000, False))
    only = set(args.only.split(",")) if args.only else None
    eng.lib.fb_set_profiling(eng.handle, 1)
    for name, f in schemes.items():
        if only and name not in only:
            continue
        for _ in range(2):
            f()
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(es)
        for _ in range(args.steps):
            f()
        e1.record(es); eng.sync()
        ms = e0.elapsed_time(e1) / args.steps
        k = float(eng.lib.fb_kernel_ms(eng.handle))
        byt = n_rec * n * esz + int(ol.sum().item())
        print(json.dumps({"scheme": name, "dtype": args.dtype, "recordings": n_rec, "ms": round(ms, 3), "kernel_ms": round(k, 3),
                          "gsamples_per_s": round(n_rec * n / ms / 1e6, 1), "GBps": round(byt / ms / 1e6, 1),
                          "frac_of_measured_hbm": round(byt / ms / 1e6 / peak, 4), "raw_bytes": int(ol.sum().item())}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--recordings", type=int, default=64)
    ap.add_argument("--seconds", type=int, default=180)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--only", default="")
    run(ap.parse_args())
