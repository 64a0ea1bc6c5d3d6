"""Seeded synthetic-signal generators.  TEST INFRASTRUCTURE ONLY.

Vectorised restatements of the reference *modulators* (TX side, out of product scope;
used only to make inputs): bpsk_modulate (modem.py:28-65), qpsk_modulate (modem.py:138-186),
fsk_modulate (modem.py:270-295).  tools/make_golden.py checks them bit-for-bit against the
reference modulators on short inputs (the reference loops cost ~10 us/symbol, these ~20 ns).

`ramp_free=True` is NOT reference behaviour: the reference modulators crash when
int(sps * 0.1) == 0 (modem.py:59-61,181-183); the flag skips the edge ramp so that
high-baud (sps < 10) demodulator inputs can still be synthesised.
"""
from __future__ import annotations

import math

import numpy as np

from .frames import frame_data


def _envelope(sps: int, ramp_free: bool) -> np.ndarray:
    env = np.ones(sps, dtype=np.float64)
    ramp = int(sps * 0.1)
    if ramp == 0:
        if not ramp_free:
            raise ValueError("could not broadcast input array from shape (0,) into shape (%d,)" % sps)
        return env
    env[:ramp] = np.linspace(0, 1, ramp)
    env[-ramp:] = np.linspace(1, 0, ramp)
    return env


def _psk_wave(phases: np.ndarray, sps: int, carrier: float, samp_rate: int, ramp_free: bool) -> np.ndarray:
    t_symbol = np.arange(sps) / samp_rate
    base = 2 * np.pi * carrier * t_symbol                     # modem.py:54,178 (restarts every symbol)
    env = _envelope(sps, ramp_free)
    out = np.empty(len(phases) * sps, dtype=np.float32)
    step = 1 << 16
    for i in range(0, len(phases), step):                      # bounded temporaries
        ph = phases[i: i + step]
        out[i * sps: (i + len(ph)) * sps] = (np.sin(base[None, :] + ph[:, None]) * env[None, :]).reshape(-1)
    return out


def bpsk_modulate(data_bytes: bytes, baud=1200, carrier=3000.0, samp_rate=96000, ramp_free=False) -> np.ndarray:
    """modem.py:28-65 (DBPSK; preamble [1,0]*40; 1 -> +pi)."""
    bits = np.unpackbits(np.frombuffer(data_bytes, dtype=np.uint8))
    bits = np.concatenate([np.tile(np.array([1, 0], np.uint8), 40), bits])
    sps = int(samp_rate / baud)
    phases = np.cumsum(np.where(bits == 1, np.pi, 0.0))        # sequential float adds, as :44-48
    return _psk_wave(phases, sps, carrier, samp_rate, ramp_free)


def qpsk_modulate(data_bytes: bytes, baud=1200, carrier=3000.0, samp_rate=96000, ramp_free=False) -> np.ndarray:
    """modem.py:138-186 (DQPSK; preamble [0,0]*30+[1,1]*10; 00->0, 01->+pi/2, 11->pi, 10->-pi/2)."""
    bits = np.unpackbits(np.frombuffer(data_bytes, dtype=np.uint8))
    pre = np.array([0, 0] * 30 + [1, 1] * 10, dtype=np.uint8)
    bits = np.concatenate([pre, bits])
    di = bits.reshape(-1, 2)
    code = di[:, 0].astype(np.int64) * 2 + di[:, 1]
    table = np.array([0.0, np.pi / 2, -np.pi / 2, np.pi])      # index = 2*b0 + b1
    sps = int(samp_rate / baud)
    phases = np.cumsum(table[code])                            # modem.py:170-174
    return _psk_wave(phases, sps, carrier, samp_rate, ramp_free)


def fsk_modulate(data_bytes: bytes, baud=1200, mark_freq=1200.0, space_freq=2200.0, samp_rate=96000) -> np.ndarray:
    """modem.py:270-295 (CPFSK, preamble AA AA AA AA, x0.9)."""
    spb = int(round(samp_rate * (1.0 / baud)))
    t = np.arange(spb) / samp_rate
    bits = np.unpackbits(np.frombuffer(b"\xAA\xAA\xAA\xAA" + data_bytes, dtype=np.uint8))
    freqs = np.where(bits == 1, mark_freq, space_freq)
    # the phase carry is sequential with a modulo (modem.py:292-293): scalar loop, exact
    phases = np.empty(len(bits), dtype=np.float64)
    phase = 0
    two_pi = 2 * np.pi
    ratio = spb / samp_rate
    for i, f in enumerate(freqs.tolist()):
        phases[i] = phase
        phase += two_pi * f * ratio
        phase %= two_pi
    out = np.empty(len(bits) * spb, dtype=np.float32)
    step = 1 << 15
    for i in range(0, len(bits), step):
        fr = freqs[i: i + step]
        ph = phases[i: i + step]
        arg = (2 * np.pi * fr)[:, None] * t[None, :] + ph[:, None]
        out[i * spb: (i + len(fr)) * spb] = (np.sin(arg).astype(np.float32) * np.float32(0.9)).reshape(-1)
    return out


def add_awgn(x: np.ndarray, snr_db: float, rng: np.random.Generator) -> np.ndarray:
    """SURVEY App. C.2 noise recipe: float64 add, noise power from mean(x^2) over the record."""
    x64 = x.astype(np.float64)
    x64 = x64 + rng.standard_normal(len(x64)) * math.sqrt(float(np.mean(x64 ** 2)) / 10 ** (snr_db / 10))
    return x64.astype(np.float32)


def kat_signal(mod, seed: int, nbytes: int, snr_db: float, name: str = "kat.bin", **mod_kw):
    """(payload, framed bytes, noisy float32 signal)."""
    rng = np.random.default_rng(seed)
    payload = rng.integers(0, 256, nbytes, dtype=np.uint8).tobytes()
    framed = frame_data(name, payload, 0, 1, nbytes, 0)
    x = add_awgn(mod(framed, **mod_kw), snr_db, rng)
    return payload, framed, x
