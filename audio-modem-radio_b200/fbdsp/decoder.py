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
    """What soundfile.read hands decoder.decode_wav_file (decoder.py:381), and the sample rate: PCM16 stays an int16 array of
    shape (frames,) or (frames, channels) -- channel selection and the 1/32768 scaling happen on the device --, every other
    WAV encoding soundfile accepts (8-bit unsigned, 24- and 32-bit PCM, 32- and 64-bit IEEE float, plain or
    WAVE_FORMAT_EXTENSIBLE) becomes the float64 array soundfile would return (integers scaled by 2^-(bits-1))."""
    import struct
    with open(path, "rb") as f:
        blob = f.read()
    if len(blob) < 12 or blob[:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(blob):
        cid, size = blob[pos:pos + 4], struct.unpack("<I", blob[pos + 4:pos + 8])[0]
        body = blob[pos + 8: pos + 8 + size]
        if cid == b"fmt " and fmt is None:
            fmt = body
        elif cid == b"data" and data is None:
            data = body
        pos += 8 + size + (size & 1)
    if fmt is None or data is None or len(fmt) < 16:
        raise ValueError(f"{path}: missing fmt / data chunk")
    tag, nch, sr, _, block, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == 0xFFFE and len(fmt) >= 26:                                  # WAVE_FORMAT_EXTENSIBLE: the sub-format GUID's first word
        tag = struct.unpack("<H", fmt[24:26])[0]
    nch = max(1, nch)
    frame = nch * (bits // 8)
    data = data[: len(data) // frame * frame] if frame else b""
    if tag == 1 and bits == 16:
        pcm = np.frombuffer(data, dtype="<i2")
    elif tag == 1 and bits == 8:
        pcm = (np.frombuffer(data, dtype=np.uint8).astype(np.float64) - 128.0) / 128.0
    elif tag == 1 and bits == 24:
        b3 = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b3[:, 0] | (b3[:, 1] << 8) | (b3[:, 2] << 16)
        pcm = (v - ((v & 0x800000) << 1)).astype(np.float64) / 8388608.0
    elif tag == 1 and bits == 32:
        pcm = np.frombuffer(data, dtype="<i4").astype(np.float64) / 2147483648.0
    elif tag == 3 and bits in (32, 64):
        pcm = np.frombuffer(data, dtype="<f4" if bits == 32 else "<f8").astype(np.float64)
    else:
        raise ValueError(f"{path}: unsupported WAV encoding (format tag {tag}, {bits} bits)")
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


def decode_wav_files(paths: Sequence[str], modes, symbol_rates, engine: Optional[Engine] = None) -> list:
    """Batch form of decode_wav_file: every file read and brought to 96 kHz like decoder.py:381-387, then ONE
    decode_corpus call (recordings grouped by parameter set).  Returns a CorpusResult per file; files the reference would
    fail on (unreadable, FSK default tones, ...) carry their error text instead of raising."""
    n = len(paths)
    modes = [modes] * n if isinstance(modes, str) else list(modes)
    rates = [symbol_rates] * n if np.isscalar(symbol_rates) else list(symbol_rates)
    recs, errs = [], {}
    for i, pth in enumerate(paths):
        try:
            recs.append(_wav_samples(pth))
        except Exception as e:      # noqa: BLE001  (decode_wav_file would raise before decode_from_buffer's try)
            recs.append(np.zeros(0, np.float32))
            errs[i] = f"{type(e).__name__}: {e}"
    out = decode_corpus(recs, modes, rates, engine=engine, pcm16=True)
    for i, e in errs.items():
        out[i] = BatchResult(b"", -1, 0, [], e)
    return out


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
    n = len(recordings)
    return decode_corpus(recordings, [mode] * n, [symbol_rate] * n, engine=engine, carriers=[carrier] * n, pcm16=True)


# ------------------------------------------------------------------------------------------ corpus driver
def corpus_groups(modes: Sequence[str], symbol_rates: Sequence, carriers=None, tones=None, dtypes=None):
    """decoder.py:422-434 applied to every recording of a corpus, then grouped: {(kind, baud, carrier | (mark, space),
    dtype): [recording indices]} in first-appearance order.  carriers / tones are per-recording overrides of what
    decode_from_buffer hard-wires (carrier 3000.0 via the demodulators' defaults; FSK tones 1200.0 / 2200.0, for which the
    reference's Butterworth design raises) -- None entries keep the reference's values; a tones entry may carry a third
    element, the bit rate, for FSK rates the three mode strings cannot name."""
    groups = {}
    for i, (mode, sr) in enumerate(zip(modes, symbol_rates)):
        kind, baud = mode_params(mode, sr)
        if kind == "fsk":
            par = tuple(float(v) for v in tones[i][:2]) if tones is not None and tones[i] is not None else (1200.0, 2200.0)
            if tones is not None and tones[i] is not None and len(tones[i]) > 2:
                baud = tones[i][2]                                     # (mark, space, baud): bit rates the mode strings cannot name
        else:
            par = float(carriers[i]) if carriers is not None and carriers[i] is not None else 3000.0
        dt = np.dtype(dtypes[i]).str if dtypes is not None else ""
        groups.setdefault((kind, baud, par, dt), []).append(i)
    return groups


def _status_error(st: int, pad: int) -> Optional[str]:
    from . import _lib
    from .engine import PADLEN_MSG
    if st == _lib.FB_ST_TOO_SHORT:
        return "ValueError: " + PADLEN_MSG % pad                      # scipy.signal.filtfilt, as in the reference
    if st == _lib.FB_ST_UNSUPPORTED:
        return "FbdspError: recording too long for a parameter set that needs full-window evaluation"
    return None


def decode_corpus(recordings: Sequence, modes: Sequence[str], symbol_rates: Sequence, engine: Optional[Engine] = None,
                  carriers=None, tones=None, pcm16: bool = False, parse: bool = True) -> List[BatchResult]:
    """A corpus of recordings with per-recording (mode, symbol rate): the reference would loop decode_from_buffer over
    them (decoder.py:417-464); here they are grouped by parameter set (corpus_groups) and every group -- DPSK and FSK
    alike -- is demodulated by ONE batched launch sequence, then all raw streams are parsed in one device call.
    Results come back in the caller's order.  A recording the reference fails on (scipy's ValueError for the FSK default
    tones or for N <= padlen) gets `.error` and no frames, exactly where decode_from_buffer returns []; it never fails
    the corpus.  int16 arrays are PCM16 (value / 32768) when pcm16 is set, integer-valued samples otherwise."""
    from . import fsk as _fsk
    from .engine import _as_samples
    eng = engine or default_engine()
    n = len(recordings)
    if not (len(modes) == len(symbol_rates) == n):
        raise ValueError("recordings, modes and symbol_rates must have the same length")
    xs = []
    for x in recordings:
        if isinstance(x, _Pcm16):
            xs.append(np.ascontiguousarray(x.pcm, dtype=np.int16))
        elif pcm16 and isinstance(x, np.ndarray) and x.dtype == np.int16:
            xs.append(np.ascontiguousarray(x))
        else:
            xs.append(_as_samples(x))
    out: List[Optional[BatchResult]] = [None] * n
    groups = corpus_groups(modes, symbol_rates, carriers, tones, [x.dtype for x in xs])
    for (kind, baud, par, _dt), idx in groups.items():
        try:
            if kind == "fsk":
                d = _fsk.fsk_design(baud, par[0], par[1], float(SAMPLE_RATE))          # raises like scipy (modem.py:307)
                res = _fsk.fsk_demod_batch([xs[i] for i in idx], d, eng)
                pad = d.pad
            else:
                d = psk_design(float(baud), par, float(SAMPLE_RATE), 1.0 if kind == "bpsk" else 1.5, kind == "bpsk")
                res = eng.psk_demod_batch([xs[i] for i in idx], d)
                pad = d.c_struct.pad_bp
            for i, r in zip(idx, res):
                out[i] = BatchResult(r.raw, r.sync_idx, r.status, [], _status_error(r.status, pad))
        except (ValueError, ZeroDivisionError, OverflowError) as e:     # the design itself is invalid: every recording of the group fails
            for i in idx:
                out[i] = BatchResult(b"", -1, 0, [], f"{type(e).__name__}: {e}")
    if parse and n:
        ok = [i for i in range(n) if out[i].error is None]
        for i, fr in zip(ok, _frames.parse_batch([out[i].raw for i in ok], eng, full=True)):
            out[i].frames = fr
    return out
