"""Batch demodulation engine: the Python side of libfbdsp.so.

One Engine = one fb_handle = one CUDA stream + device workspace on one GPU.  Recordings that share a
parameter set are demodulated in ONE call (CSR offsets, no padding); results come back per recording
with their own status, so one bad recording never fails a batch.
"""
from __future__ import annotations

import ctypes
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .design import PskDesign, psk_design

PADLEN_MSG = "The length of the input vector x must be greater than padlen, which is %d."


@dataclass
class DemodResult:
    raw: bytes            # what the reference's X_demodulate returns
    sync_idx: int         # bit index of the first "FB" magic (modem.py:118,248,330), -1 if absent
    status: int           # _lib.FB_ST_*

    def raise_for_status(self, padlen: int = 27):
        if self.status == _lib.FB_ST_TOO_SHORT:
            raise ValueError(PADLEN_MSG % padlen)          # scipy.signal.filtfilt's message, as in the reference
        if self.status == _lib.FB_ST_UNSUPPORTED:
            raise _lib.FbdspError("recording too long for a parameter set that needs full-window evaluation")


def _as_samples(x) -> np.ndarray:
    """What the reference's numpy/scipy code would make of `samples`: any real array-like becomes float;
    float32 and float64 are passed through untouched (no copy), everything else is promoted to float64."""
    if type(x).__name__ == "_Pcm16":       # decoder.decode_wav_file: PCM16 straight from the WAV (device scales by 1/32768)
        return np.ascontiguousarray(x.pcm, dtype=np.int16)
    a = np.asarray(x)
    if a.ndim != 1:
        a = a.reshape(-1) if a.ndim == 0 else a
        if a.ndim != 1:
            raise ValueError("samples must be one-dimensional")
    if a.dtype == np.float32 or a.dtype == np.float64:
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a, dtype=np.float64)


def flat_view(recordings: Sequence[np.ndarray]) -> np.ndarray:
    """The recordings back to back as ONE array: without a copy when they already are consecutive slices of one buffer
    (e.g. the parts of a pinned staging ring), else concatenated."""
    if len(recordings) == 1:
        return np.ascontiguousarray(recordings[0])
    r0 = recordings[0]
    if all(r.flags.c_contiguous for r in recordings):
        p = r0.ctypes.data
        for r in recordings:
            if r.ctypes.data != p:
                break
            p += r.nbytes
        else:
            n = sum(len(r) for r in recordings)
            return np.ndarray((n,), dtype=r0.dtype, buffer=(ctypes.c_char * (n * r0.itemsize)).from_address(r0.ctypes.data))
    return np.ascontiguousarray(np.concatenate(recordings))


_DT = {np.dtype(np.float32): _lib.FB_F32, np.dtype(np.float64): _lib.FB_F64, np.dtype(np.int16): _lib.FB_S16}


class Engine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if self.lib.fb_abi_version() != 1:
            raise _lib.FbdspError("libfbdsp ABI mismatch")
        n = self.lib.fb_device_count()
        if n <= 0:
            raise _lib.FbdspError("no CUDA device visible: fbdsp has no CPU fallback")
        self.device = device
        self.handle = self.lib.fb_create(device)
        if not self.handle:
            raise _lib.FbdspError(f"fb_create({device}) failed (need an sm_100 device; {n} CUDA device(s) visible)")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.fb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001  (interpreter shutdown)
            pass

    @property
    def stream(self) -> int:
        return int(self.lib.fb_stream(self.handle) or 0)

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.fb_kernel_launches(self.handle))

    def sync(self):
        _lib.check(self.lib, self.handle, self.lib.fb_sync(self.handle), "fb_sync")

    # ------------------------------------------------------------------ WAV ingest
    def ingest_resample(self, frames: np.ndarray, n_out: int) -> np.ndarray:
        """Channel 0 of interleaved frames (int16 PCM / float32 / float64; shape (n,) or (n, channels)) as float64,
        resampled to n_out samples with scipy.signal.resample's FFT method on the device (decoder.py:381-387)."""
        a = np.ascontiguousarray(frames)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        if np.dtype(a.dtype) not in _DT:
            a = np.ascontiguousarray(a, dtype=np.float64)
        n, nch = a.shape
        out = np.empty(int(n_out), dtype=np.float64)
        rc = self.lib.fb_ingest_resample(self.handle, a.ctypes.data, n, nch, _DT[np.dtype(a.dtype)], int(n_out), out.ctypes.data, 0)
        _lib.check(self.lib, self.handle, rc, "fb_ingest_resample")
        return out

    # ------------------------------------------------------------------ DPSK
    def psk_demod_raw(self, d: PskDesign, samples_ptr: int, offsets: np.ndarray, dtype: int, flags: int,
                      out_ptr: int, out_offsets: np.ndarray, out_len_ptr: int, sync_ptr: int, status_ptr: int):
        """Thin call-through: pointers may be host or device according to `flags` (see include/fbdsp.h)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out_offsets = np.ascontiguousarray(out_offsets, dtype=np.uint64)
        n_rec = len(offsets) - 1
        taps = d.taps.ctypes.data if d.taps is not None else None
        sw = d.slow_w.ctypes.data if d.slow_w is not None else None
        rc = self.lib.fb_psk_demod_batch(
            self.handle, ctypes.byref(d.c_struct), taps, sw, n_rec, samples_ptr,
            offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), dtype, flags, out_ptr,
            out_offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), out_len_ptr, sync_ptr, status_ptr)
        _lib.check(self.lib, self.handle, rc, "fb_psk_demod_batch")

    def out_bounds(self, d: PskDesign, lengths: Sequence[int]) -> np.ndarray:
        """Output slot offsets (4-byte aligned) for recordings of the given lengths."""
        sizes = np.array([int(self.lib.fb_psk_out_bound(ctypes.byref(d.c_struct), int(n))) for n in lengths], dtype=np.uint64)
        sizes = (sizes + np.uint64(3)) // np.uint64(4) * np.uint64(4)
        return np.concatenate([[np.uint64(0)], np.cumsum(sizes, dtype=np.uint64)]).astype(np.uint64)

    def psk_demod_batch(self, recordings: Sequence[np.ndarray], d: PskDesign, exact_silence: bool = True) -> List[DemodResult]:
        """Host-buffer batch call: recordings (same dtype: float32, float64 or int16 PCM) -> DemodResult each.

        exact_silence: recordings that contain a long run of EXACT zeros (zero-padded WAVs without a noise floor) are
        demodulated by the float64 step-by-step evaluation of the reference's recurrences (the design's `emulate_only`
        path): inside such a run the reference's decisions (modem.py:216-241) ride on the decaying leakage of its IIR
        state, far below what the factorised float32 path resolves."""
        if len(recordings) == 0:
            return []
        if exact_silence and not d.emulate_only:
            min_run = max(64, int(d.w_bp) // 5)            # leakage still >= 1e-2 of the signal next to it: float32 is exact enough
            silent = [i for i, r in enumerate(recordings) if len(r) <= EMULATE_MAX_SAMPLES and has_zero_run(r, min_run)]
            if silent:
                de = emulating_copy(d)
                res: List[Optional[DemodResult]] = [None] * len(recordings)
                for i, v in zip(silent, self.psk_demod_batch([recordings[i] for i in silent], de, exact_silence=False)):
                    res[i] = v
                rest = [i for i in range(len(recordings)) if res[i] is None]
                for i, v in zip(rest, self.psk_demod_batch([recordings[i] for i in rest], d, exact_silence=False)):
                    res[i] = v
                return res      # type: ignore[return-value]
        dt = recordings[0].dtype
        if any(r.dtype != dt for r in recordings):
            raise ValueError("all recordings of one batch must share a dtype")
        if np.dtype(dt) not in _DT:
            raise ValueError(f"unsupported sample dtype {dt}")
        lengths = [len(r) for r in recordings]
        offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
        flat = flat_view(recordings)
        out_offsets = self.out_bounds(d, lengths)
        out = np.empty(int(out_offsets[-1]) + 4, dtype=np.uint8)
        n = len(recordings)
        out_len = np.zeros(n, dtype=np.uint64)
        sync = np.zeros(n, dtype=np.int64)
        status = np.zeros(n, dtype=np.int32)
        self.psk_demod_raw(d, flat.ctypes.data, offsets, _DT[np.dtype(dt)], 0, out.ctypes.data, out_offsets,
                           out_len.ctypes.data, sync.ctypes.data, status.ctypes.data)
        res = []
        for r in range(n):
            o = int(out_offsets[r])
            res.append(DemodResult(out[o:o + int(out_len[r])].tobytes(), int(sync[r]), int(status[r])))
        return res

    def last_bits(self, rec: int) -> np.ndarray:
        """Decided bit stream (0/1 array) of recording `rec` of the last DPSK batch (test hook)."""
        nb = ctypes.c_uint64(0)
        _lib.check(self.lib, self.handle, self.lib.fb_psk_last_bits(self.handle, rec, None, 0, ctypes.byref(nb)), "fb_psk_last_bits")
        buf = np.zeros((nb.value + 7) // 8, dtype=np.uint8)
        _lib.check(self.lib, self.handle,
                   self.lib.fb_psk_last_bits(self.handle, rec, buf.ctypes.data, len(buf), ctypes.byref(nb)), "fb_psk_last_bits")
        return np.unpackbits(buf)[: nb.value]


EMULATE_MAX_SAMPLES = 1 << 26            # whole-record float64 evaluation: csrc/psk_v2.cu EMU_MAX


def has_zero_run(x: np.ndarray, min_run: int) -> bool:
    """True when x holds at least `min_run` consecutive samples that are exactly zero.

    A run of >= min_run zeros covers at least one index that is a multiple of min_run, so only every min_run-th sample is
    inspected (N / min_run reads for a recording without such a run -- every noisy one) and the runs around the zeros found
    there are measured."""
    n = len(x)
    min_run = max(1, int(min_run))
    if n < min_run:
        return False
    hits = np.flatnonzero(x[::min_run] == 0) * min_run
    last_end = -1
    for i in hits:
        i = int(i)
        if i <= last_end:                                  # inside a run that was already measured
            continue
        lo = max(0, i - min_run + 1)
        hi = min(n, i + min_run)
        w = np.flatnonzero(x[lo:hi])                        # non-zero samples of the window around the hit
        left = w[w < i - lo]
        right = w[w > i - lo]
        start = lo + int(left[-1]) + 1 if len(left) else lo
        end = lo + int(right[0]) if len(right) else hi      # first non-zero after the hit (window end if none)
        if end - start >= min_run:
            return True
        last_end = end
    return False


def emulating_copy(d: PskDesign) -> PskDesign:
    """The same parameter set evaluated step by step in float64 on the device (no factorised interior kernel)."""
    import copy
    import dataclasses
    cs = copy.copy(d.c_struct)                             # ctypes structures copy by value
    cs.emulate_only = 1
    return dataclasses.replace(d, emulate_only=True, c_struct=cs)


_default: Optional[Engine] = None
_default_lock = threading.Lock()


def default_engine() -> Engine:
    """Process-wide engine on FBDSP_DEVICE (default 0), created once on first use.  It may be shared by threads (the
    reference calls decode_from_buffer from the Qt GUI thread and from a QThread, filebeep_advanced_v2.py:324,1112): the
    library serialises the calls of one handle (include/fbdsp.h)."""
    global _default
    with _default_lock:
        if _default is None:
            import os
            _default = Engine(int(os.environ.get("FBDSP_DEVICE", "0")))
    return _default


def demod_psk(samples, baud, carrier, samp_rate, band_k, n0_is_sps, engine: Optional[Engine] = None) -> bytes:
    d = psk_design(float(baud), float(carrier), float(samp_rate), float(band_k), bool(n0_is_sps))   # raises like scipy
    x = _as_samples(samples)
    if len(x) <= d.c_struct.pad_bp:
        raise ValueError(PADLEN_MSG % d.c_struct.pad_bp)
    res = (engine or default_engine()).psk_demod_batch([x], d)[0]
    res.raise_for_status(d.c_struct.pad_bp)
    return res.raw
