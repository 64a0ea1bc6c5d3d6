"""Oracle: vectorised numpy/scipy restatement of the reference's executable modem.py ("v2").

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the reference
lines it restates.  The numerically sensitive arithmetic is the same third-party code the
reference calls: scipy.signal.{butter,filtfilt,hilbert} and numpy (oracle pinned on
numpy 2.3.5 / scipy 1.18.1 -- the reference pins no versions).  What is restated here is
the reference's own Python: the per-symbol / per-bit loops are vectorised and the
str.join/str.find/int(..., 2) byte packer is replaced by exact integer equivalents.

Pinned against /root/reference by tools/make_golden.py -> tests/golden/*.npz
(byte-for-byte equality of every demodulator output, including raised exception types).
"""
from __future__ import annotations

import numpy as np
from scipy import signal

SAMPLE_RATE = 96000                      # modem.py:11
MAGIC_BITS = np.array([0, 1, 0, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.uint8)  # modem.py:116,247,329


# ----------------------------------------------------------------------------- back end
def find_sync(bits: np.ndarray) -> int:
    """First occurrence of the 16-bit magic in the decided bit stream, or -1.

    Restates ``bit_str.find("0100011001000010")`` (modem.py:118, 248, 330): bits are the
    characters '0'/'1', so a byte-string search over the 0/1 array is the same search.
    """
    return np.ascontiguousarray(bits, dtype=np.uint8).tobytes().find(MAGIC_BITS.tobytes())


def pack_from(bits: np.ndarray, start: int) -> bytes:
    """MSB-first byte packing from ``start``: modem.py:121-133, 251-264, 333-339.

    ``for i in range(0, len(valid) - 7, 8): int(valid[i:i+8], 2)`` -> floor(len/8) whole bytes.
    """
    bits = np.ascontiguousarray(bits, dtype=np.uint8)[start:]
    nbytes = len(bits) // 8
    if nbytes <= 0:
        return b""
    return np.packbits(bits[: nbytes * 8]).tobytes()


def sync_and_pack(bits: np.ndarray):
    """(raw bytes, sync_idx): sync found -> pack from it, else pack from bit 0."""
    idx = find_sync(bits)
    return pack_from(bits, idx if idx != -1 else 0), idx


# ----------------------------------------------------------------------------- PSK front end
def psk_baseband(samples, baud, carrier, samp_rate, band_k):
    """Common DPSK front end (modem.py:73-88 with band_k=1, modem.py:194-204 with band_k=1.5).

    Zero-phase Butterworth-4 band-pass -> complex mix with a continuous LO ->
    zero-phase Butterworth-4 low-pass.  Raises exactly what scipy raises in the reference.
    """
    nyquist = samp_rate / 2
    low = (carrier - baud * band_k) / nyquist
    high = (carrier + baud * band_k) / nyquist
    b, a = signal.butter(4, [max(0.01, low), min(0.99, high)], btype="band")
    filtered = signal.filtfilt(b, a, samples)
    t = np.arange(len(filtered)) / samp_rate
    baseband = filtered * np.exp(-1j * 2 * np.pi * carrier * t)
    b_lp, a_lp = signal.butter(4, baud / nyquist, btype="low")
    return signal.filtfilt(b_lp, a_lp, baseband)


def qpsk_slice(diff: np.ndarray) -> np.ndarray:
    """Dibit decisions, modem.py:216-241 (same float comparisons, vectorised)."""
    ang = np.angle(diff)
    ang = np.where(ang < 0, ang + 2 * np.pi, ang)
    c00 = (ang < np.pi / 4) | (ang > 7 * np.pi / 4)
    c01 = (~c00) & (np.pi / 4 <= ang) & (ang < 3 * np.pi / 4)
    c11 = (~c00) & (~c01) & (3 * np.pi / 4 <= ang) & (ang < 5 * np.pi / 4)
    hi = np.where(c00 | c01, 0, 1).astype(np.uint8)          # 00,01 -> 0 ; 11,10 -> 1
    lo = np.where(c01 | c11, 1, 0).astype(np.uint8)          # 01,11 -> 1 ; 00,10 -> 0
    bits = np.empty(2 * len(diff), dtype=np.uint8)
    bits[0::2] = hi
    bits[1::2] = lo
    return bits


def qpsk_margin(diff: np.ndarray) -> np.ndarray:
    """Soft-metric margin per differential symbol: angular distance to the nearest
    sector edge divided by pi/4 (0 = on an edge, 1 = sector centre).  SURVEY 8d config 4."""
    ang = np.mod(np.angle(diff), 2 * np.pi)
    r = np.mod(ang - np.pi / 4, np.pi / 2)
    return np.minimum(r, np.pi / 2 - r) / (np.pi / 4)


def qpsk_stages(samples, baud=1200, carrier=3000.0, samp_rate=96000):
    """All intermediate stages of qpsk_demodulate (modem.py:189-266) for the parity tests."""
    sps = int(samp_rate / baud)
    bb = psk_baseband(samples, baud, carrier, samp_rate, 1.5)
    symbols = bb[sps // 2 :: sps]                             # modem.py:209
    if len(symbols) < 2:                                      # modem.py:211
        return dict(symbols=symbols, diff=None, bits=None, sync=-1, raw=b"")
    diff = symbols[1:] * np.conj(symbols[:-1])                # modem.py:214
    bits = qpsk_slice(diff)
    raw, idx = sync_and_pack(bits)
    return dict(symbols=symbols, diff=diff, bits=bits, sync=idx, raw=raw)


def qpsk_demodulate(samples, baud=1200, carrier=3000.0, samp_rate=96000) -> bytes:
    """modem.py:189-266."""
    return qpsk_stages(samples, baud, carrier, samp_rate)["raw"]


def bpsk_stages(samples, baud=1200, carrier=3000.0, samp_rate=96000):
    """All intermediate stages of bpsk_demodulate (modem.py:68-135)."""
    sps = int(samp_rate / baud)
    bb = psk_baseband(samples, baud, carrier, samp_rate, 1.0)
    symbols = bb[sps::sps]                                    # modem.py:92-93 (edge sampled)
    if len(symbols) < 2:                                      # modem.py:95-96
        return dict(symbols=symbols, diff=None, bits=None, sync=-1, raw=b"")
    diff = symbols[1:] * np.conj(symbols[:-1])                # modem.py:100
    bits = (np.real(diff) < 0).astype(np.uint8)               # modem.py:105
    raw, idx = sync_and_pack(bits)
    return dict(symbols=symbols, diff=diff, bits=bits, sync=idx, raw=raw)


def bpsk_margin(diff: np.ndarray) -> np.ndarray:
    """|Re d| / |d|."""
    mag = np.abs(diff)
    return np.where(mag > 0, np.abs(diff.real) / np.where(mag > 0, mag, 1), 0.0)


def bpsk_demodulate(samples, baud=1200, carrier=3000.0, samp_rate=96000) -> bytes:
    """modem.py:68-135."""
    return bpsk_stages(samples, baud, carrier, samp_rate)["raw"]


# ----------------------------------------------------------------------------- FSK
def fsk_stages(samples, baud=1200, mark_freq=1200.0, space_freq=2200.0, samp_rate=96000):
    """All intermediate stages of fsk_demodulate (modem.py:298-341)."""
    samples = np.asarray(samples)
    spb = int(samp_rate / baud)                               # modem.py:301
    nyq = samp_rate / 2

    def get_envelope(freq):                                   # modem.py:306-309
        b, a = signal.butter(3, [(freq - baud) / nyq, (freq + baud) / nyq], btype="band")
        filt = signal.filtfilt(b, a, samples)
        return np.abs(signal.hilbert(filt))

    mark_env = get_envelope(mark_freq)
    space_env = get_envelope(space_freq)
    cmp_ = (mark_env > space_env).astype(np.int64)            # modem.py:315
    n = len(cmp_)
    q = spb // 4
    centres = np.arange(spb // 2, n, spb)                     # modem.py:320
    if q == 0 or len(centres) == 0:                           # empty chunk -> no bit appended (:322)
        decided = np.zeros(0, dtype=np.uint8)
    else:
        lo = centres - q                                      # >= 0 because spb//2 >= spb//4
        hi = np.minimum(centres + q, n)                       # python slicing truncates at n
        csum = np.concatenate(([0], np.cumsum(cmp_)))
        ones = csum[hi] - csum[lo]
        length = hi - lo
        keep = length > 0
        # np.mean(chunk) > 0.5  <=>  2*ones > length (exact for these small integers)
        decided = (2 * ones[keep] > length[keep]).astype(np.uint8)
    idx = find_sync(decided)                                  # modem.py:329-330
    raw = pack_from(decided, idx if idx != -1 else 0)         # modem.py:333-339
    return dict(mark_env=mark_env, space_env=space_env, bits=decided, sync=idx, raw=raw)


def fsk_demodulate(samples, baud=1200, mark_freq=1200.0, space_freq=2200.0, samp_rate=96000) -> bytes:
    """modem.py:298-341.  Raises ValueError for every product default (f - baud <= 0)."""
    return fsk_stages(samples, baud, mark_freq, space_freq, samp_rate)["raw"]


# ----------------------------------------------------------------------------- aliases
def psk8_demodulate(s, b=1200, c=3000.0, s_r=96000):
    """modem.py:348 -- 8PSK is an alias of QPSK; parameter names are part of the contract."""
    return qpsk_demodulate(s, b, c, s_r)


def fsk_high_speed_demodulate(s, baud=19200, s_r=96000):
    """modem.py:355-356."""
    return fsk_demodulate(s, baud, 8000, 16000, s_r)


def ofdm_demodulate_simple(s, baud, carrier, num_subcarriers, samp_rate=96000):
    """modem.py:375-376 -- OFDM is an alias of QPSK; num_subcarriers is ignored."""
    return qpsk_demodulate(s, baud, carrier, samp_rate)


def ft8_demodulate(s, b, c, sr=96000):
    """modem.py:391."""
    return fsk_demodulate(s, 50, c, c + 50, sr)


def psk31_demodulate(s, b, c, sr=96000):
    """modem.py:397."""
    return bpsk_demodulate(s, 31.25, c, sr)
