"""ctypes binding of libfbdsp.so (include/fbdsp.h).  No fallback: if the library or a B200 is missing,
the product raises instead of computing anything on the CPU."""
from __future__ import annotations

import ctypes
import os

from .design import fb_psk_design

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FBDSP_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libfbdsp.so"))

FB_F32, FB_F64, FB_S16 = 0, 1, 2
FB_SAMPLES_ON_DEVICE, FB_OUT_ON_DEVICE, FB_ASYNC = 1, 2, 4
FB_ST_OK, FB_ST_EMPTY, FB_ST_TOO_SHORT, FB_ST_UNSUPPORTED = 0, 1, 2, 3

# every symbol include/fbdsp.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "fb_abi_version", "fb_device_count", "fb_strerror", "fb_last_error", "fb_create", "fb_destroy", "fb_stream",
    "fb_sync", "fb_kernel_launches", "fb_set_profiling", "fb_kernel_ms", "fb_psk_out_bound", "fb_psk_demod_batch", "fb_psk_last_bits", "fb_debug_mma_band",
    "fb_rs_out_bound", "fb_viterbi_out_bound", "fb_rs_decode_batch", "fb_rs_decode_spans", "fb_viterbi_decode_batch", "fb_crc32_batch",
    "fb_parse_frames_batch", "fb_fsk_out_bound", "fb_fsk_demod_batch", "fb_v1_out_bound", "fb_v1_demod_batch",
    "fb_ingest_resample", "fb_mod_out_samples", "fb_modulate_batch",
]


class fb_frame(ctypes.Structure):
    """Mirror of `struct fb_frame` in include/fbdsp.h."""
    _fields_ = [("offset", ctypes.c_uint64), ("name_off", ctypes.c_uint64), ("payload_off", ctypes.c_uint64),
                ("name_len", ctypes.c_uint32), ("part", ctypes.c_uint32), ("total", ctypes.c_uint32),
                ("file_size", ctypes.c_uint32), ("file_crc", ctypes.c_uint32), ("data_len", ctypes.c_uint32),
                ("payload_crc", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class FbdspError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FbdspError(f"{LIB_PATH} not found: build it with audio-modem-radio_b200/build.sh "
                         f"(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, u64p, i64p, i32p, u8p = c.c_void_p, c.POINTER(c.c_uint64), c.POINTER(c.c_int64), c.POINTER(c.c_int32), c.c_void_p
    lib.fb_abi_version.restype = c.c_int
    lib.fb_device_count.restype = c.c_int
    lib.fb_strerror.restype = c.c_char_p
    lib.fb_strerror.argtypes = [c.c_int]
    lib.fb_last_error.restype = c.c_char_p
    lib.fb_last_error.argtypes = [vp]
    lib.fb_create.restype = vp
    lib.fb_create.argtypes = [c.c_int]
    lib.fb_destroy.argtypes = [vp]
    lib.fb_destroy.restype = None
    lib.fb_stream.restype = vp
    lib.fb_stream.argtypes = [vp]
    lib.fb_sync.argtypes = [vp]
    lib.fb_kernel_launches.restype = c.c_uint64
    lib.fb_kernel_launches.argtypes = [vp]
    lib.fb_set_profiling.argtypes = [vp, c.c_int]
    lib.fb_kernel_ms.restype = c.c_float
    lib.fb_kernel_ms.argtypes = [vp]
    lib.fb_psk_out_bound.restype = c.c_uint64
    lib.fb_psk_out_bound.argtypes = [c.POINTER(fb_psk_design), c.c_uint64]
    lib.fb_psk_demod_batch.restype = c.c_int
    lib.fb_psk_demod_batch.argtypes = [vp, c.POINTER(fb_psk_design), vp, vp, c.c_int, vp, u64p, c.c_int, c.c_int,
                                       u8p, u64p, vp, vp, vp]
    lib.fb_debug_mma_band.restype = c.c_int
    lib.fb_debug_mma_band.argtypes = [c.POINTER(fb_psk_design), vp, vp]
    lib.fb_psk_last_bits.restype = c.c_int
    lib.fb_psk_last_bits.argtypes = [vp, c.c_int, vp, c.c_uint64, u64p]
    lib.fb_rs_out_bound.restype = c.c_uint64
    lib.fb_rs_out_bound.argtypes = [c.c_uint64]
    lib.fb_viterbi_out_bound.restype = c.c_uint64
    lib.fb_viterbi_out_bound.argtypes = [c.c_uint64]
    lib.fb_rs_decode_batch.restype = c.c_int
    lib.fb_rs_decode_batch.argtypes = [vp, c.c_int, vp, u64p, vp, u64p, vp, vp, c.c_int]
    lib.fb_rs_decode_spans.restype = c.c_int
    lib.fb_rs_decode_spans.argtypes = [vp, c.c_int, vp, u64p, u64p, vp, u64p, vp, vp, c.c_int]
    lib.fb_viterbi_decode_batch.restype = c.c_int
    lib.fb_viterbi_decode_batch.argtypes = [vp, c.c_int, vp, u64p, vp, u64p, vp, c.c_int]
    lib.fb_crc32_batch.restype = c.c_int
    lib.fb_crc32_batch.argtypes = [vp, c.c_int, vp, u64p, vp, c.c_int]
    lib.fb_parse_frames_batch.restype = c.c_int
    lib.fb_parse_frames_batch.argtypes = [vp, c.c_int, vp, u64p, vp, c.c_int, vp, vp, vp, c.c_int]
    lib.fb_fsk_out_bound.restype = c.c_uint64
    lib.fb_fsk_out_bound.argtypes = [vp, c.c_uint64]
    lib.fb_fsk_demod_batch.restype = c.c_int
    lib.fb_fsk_demod_batch.argtypes = [vp, vp, c.c_int, vp, u64p, c.c_int, c.c_int, u8p, u64p, vp, vp, vp]
    lib.fb_v1_out_bound.restype = c.c_uint64
    lib.fb_v1_out_bound.argtypes = [vp, c.c_uint64]
    lib.fb_v1_demod_batch.restype = c.c_int
    lib.fb_v1_demod_batch.argtypes = [vp, vp, vp, c.c_int, vp, u64p, c.c_int, c.c_int, u8p, u64p, vp, vp]
    lib.fb_ingest_resample.restype = c.c_int
    lib.fb_ingest_resample.argtypes = [vp, vp, c.c_uint64, c.c_int, c.c_int, c.c_uint64, vp, c.c_int]
    lib.fb_mod_out_samples.restype = c.c_uint64
    lib.fb_mod_out_samples.argtypes = [vp, c.c_uint64]
    lib.fb_modulate_batch.restype = c.c_int
    lib.fb_modulate_batch.argtypes = [vp, vp, vp, vp, c.c_int, vp, u64p, vp, u64p, c.c_int]
    _lib = lib
    return lib


def check(lib, handle, rc: int, what: str):
    if rc != 0:
        detail = lib.fb_last_error(handle).decode() if handle else ""
        raise FbdspError(f"{what}: {lib.fb_strerror(rc).decode()} {detail}".strip())
