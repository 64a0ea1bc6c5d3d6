"""Drop-in for the reference's receive driver (decoder.py:380-389, 417-464) plus the batch driver the engine adds.

    decode_wav_file(path, mode, symbol_rate) -> list[str]      same signature / return value as the reference
    decode_from_buffer(data, mode, symbol_rate) -> list[str]
    decode_batch(recordings, mode, symbol_rate) -> list[BatchResult]      many recordings, one launch per parameter set

Mode dispatch, defaults and error behaviour are the reference's (decoder.py:422-434, 460-464): every exception is
caught, a traceback is printed and [] is returned -- so FSK modes with the product's default tones yield [] exactly
as they do there (scipy's Butterworth design raises ValueError).  Compression (utils/compression.py) is out of scope
and untouched: the reference's own intelligent_decompress is used when importable, else a stdlib restatement.
"""
from __future__ import annotations

import os
import time
import traceback
import wave
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import frames as _frames
from . import modem as _modem
from .design import psk_design
from .engine import DemodResult, Engine, default_engine

SAMPLE_RATE = 96000                    # modem.py:11
RECV_DIR = "recv"                      # decoder.py:17
VERBOSE = bool(int(os.environ.get("FBDSP_VERBOSE", "0")))


def _log(msg: str):
    if VERBOSE:
        print(msg)


def _decompress(data: bytes) -> bytes:
    """utils/compression.py:103-123 (untouched subsystem): use the reference's function when it is importable."""
    try:
        from utils.compression import intelligent_decompress          # the reference tree, when on sys.path
        return intelligent_decompress(data)
    except ImportError:
        pass
    import lzma
    import zlib
    try:
        if data.startswith(b"LZMA"):
            return lzma.decompress(data[4:])
        if data.startswith(b"DLZM"):
            d = np.frombuffer(lzma.decompress(data[4:]), dtype=np.uint8)
            return np.cumsum(d, dtype=np.uint64).astype(np.uint8).tobytes() if len(d) else b""     # delta_decompress :260-273
        if data.startswith(b"ZLIB"):
            return zlib.decompress(data[4:])
        if data.startswith(b"RAW"):
            return data[4:]                                            # the reference strips 4 (compression.py:113-114)
        try:
            return zlib.decompress(data)
        except Exception:      # noqa: BLE001
            return data
    except Exception:          # noqa: BLE001
        return data


def mode_params(mode: str, symbol_rate):
    """decoder.py:422-434 -> ('bpsk'|'qpsk'|'fsk', baud)."""
    if mode == "BPSK":
        return "bpsk", symbol_rate
    if mode == "QPSK" or mode == "8PSK":
        return "qpsk", symbol_rate
    if mode.startswith("FSK"):
        baud = 1200
        if "9600" in mode:
            baud = 9600
        elif "19200" in mode:
            baud = 19200
        return "fsk", baud
    return "qpsk", symbol_rate                                         # OFDM4/8, SSTV, APSK16, DSSS, MSK, ...


def _demodulate(data, mode: str, symbol_rate) -> bytes:
    kind, baud = mode_params(mode, symbol_rate)
    if kind == "bpsk":
        return _modem.bpsk_demodulate(data, baud=baud)
    if kind == "fsk":
        return _modem.fsk_demodulate(data, baud=baud)
    return _modem.qpsk_demodulate(data, baud=baud)


def decode_from_buffer(data, mode: str, symbol_rate) -> list:
    """decoder.py:417-464."""
    _log(f"Demodulando {len(data)} amostras em modo {mode}...")
    try:
        raw_bytes = _demodulate(data, mode, symbol_rate)
        _log(f"Bytes brutos demodulados: {len(raw_bytes)}")
        frames = _frames.parse_fbp_stream_enhanced(raw_bytes)
        saved = []
        os.makedirs(RECV_DIR, exist_ok=True)
        for frame in frames:
            try:
                final_data = _decompress(frame["data"])
                ts = int(time.time())
                path = os.path.join(RECV_DIR, f"{ts}_{os.path.basename(frame['name'])}")
                with open(path, "wb") as f:
                    f.write(final_data)
                saved.append(path)
            except Exception as e:      # noqa: BLE001
                print(f"Erro salvando arquivo: {e}")
        return saved
    except Exception as e:              # noqa: BLE001  (decoder.py:460-464)
        print(f"Erro crítico na demodulação: {e}")
        traceback.print_exc()
        return []


def read_wav(path: str):
    """The PCM16 frames soundfile.read would scale for decoder.decode_wav_file (decoder.py:381): int16 array of shape
    (frames,) or (frames, channels), and the sample rate.  Channel selection and the 1/32768 scaling happen on the device."""
    with wave.open(path, "rb") as w:
        sr, nch, sw = w.getframerate(), w.getnchannels(), w.getsampwidth()
        raw = w.readframes(w.getnframes())
    if sw != 2:
        raise ValueError("only PCM16 WAV is supported")
    pcm = np.frombuffer(raw, dtype="<i2")
    return (pcm.reshape(-1, nch) if nch > 1 else pcm), sr


def _wav_samples(path: str):
    pcm, sr = read_wav(path)
    if sr != SAMPLE_RATE:               # decoder.py:385-387: FFT resampler, on the device (csrc/resample.cu)
        n = pcm.shape[0]
        return default_engine().ingest_resample(pcm, int(round(n * float(SAMPLE_RATE) / sr)))
    if pcm.ndim > 1:
        pcm = np.ascontiguousarray(pcm[:, 0])                                                    # decoder.py:382
    return pcm


def decode_wav_file(path: str, mode: str, symbol_rate) -> list:
    """decoder.py:380-389.  96 kHz PCM16 goes to the GPU as int16 (half the bytes of float32, same values)."""
    data = _wav_samples(path)
    if data.dtype == np.int16:
        data = _Pcm16(data)
    return decode_from_buffer(data, mode, symbol_rate)


class _Pcm16:
    """Marks an int16 array as PCM (value/32768) rather than integer-valued samples."""
    def __init__(self, pcm):
        self.pcm = pcm

    def __len__(self):
        return len(self.pcm)


# ------------------------------------------------------------------------------------------ batch driver
@dataclass
class BatchResult:
    raw: bytes
    sync_idx: int
    status: int
    frames: list = field(default_factory=list)       # [{'name','data','final_crc', 'part','total','file_size',...}]
    error: Optional[str] = None                      # exception text when the reference would have returned []


def decode_batch(recordings: Sequence[np.ndarray], mode: str, symbol_rate, engine: Optional[Engine] = None,
                 carrier: float = 3000.0) -> List[BatchResult]:
    """All recordings (same mode / symbol rate, any lengths; float32, float64 or int16 PCM) in one launch
    sequence: demodulate -> frame parse + CRC32 on the device.  One failing recording never fails the batch."""
    eng = engine or default_engine()
    kind, baud = mode_params(mode, symbol_rate)
    n = len(recordings)
    if kind == "fsk":
        out = []
        for x in recordings:
            try:
                raw = _modem.fsk_demodulate(x, baud=baud)
                out.append(BatchResult(raw, -1, 0, _frames.parse_batch([raw], eng, full=True)[0]))
            except Exception as e:      # noqa: BLE001
                out.append(BatchResult(b"", -1, 0, [], f"{type(e).__name__}: {e}"))
        return out
    d = psk_design(float(baud), float(carrier), float(SAMPLE_RATE), 1.0 if kind == "bpsk" else 1.5, kind == "bpsk")
    res: List[DemodResult] = eng.psk_demod_batch(list(recordings), d) if n else []
    parsed = _frames.parse_batch([r.raw for r in res], eng, full=True) if n else []
    return [BatchResult(r.raw, r.sync_idx, r.status, fr) for r, fr in zip(res, parsed)]
