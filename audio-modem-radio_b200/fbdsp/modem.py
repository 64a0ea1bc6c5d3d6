"""Drop-in for the reference's modem.py receive functions (same names, arguments, defaults, return
types and exceptions), backed by libfbdsp.so on a B200.  Swap in with e.g.

    import decoder, fbdsp.modem as fb
    decoder.qpsk_demodulate = fb.qpsk_demodulate          # decoder.py:12-14 binds these names at import

Signatures kept exactly: modem.py:68, 189, 298, 348, 355, 375, 391, 397.  There is no CPU path here.
"""
from __future__ import annotations

from . import engine as _engine

SAMPLE_RATE = 96000          # modem.py:11


def bpsk_demodulate(samples, baud=1200, carrier=3000.0, samp_rate=96000) -> bytes:
    """DBPSK, modem.py:68-135: band = carrier +- baud, symbols at bb[sps::sps], bit = Re(d) < 0."""
    return _engine.demod_psk(samples, baud, carrier, samp_rate, 1.0, True)


def qpsk_demodulate(samples, baud=1200, carrier=3000.0, samp_rate=96000) -> bytes:
    """DQPSK, modem.py:189-266: band = carrier +- 1.5 baud, symbols at bb[sps//2::sps], 4-sector slicer."""
    return _engine.demod_psk(samples, baud, carrier, samp_rate, 1.5, False)


def fsk_demodulate(samples, baud=1200, mark_freq=1200.0, space_freq=2200.0, samp_rate=96000) -> bytes:
    """modem.py:298-341.  Every product default raises ValueError from the Butterworth design, exactly as
    in the reference (mark - baud <= 0); valid tone sets run on the device."""
    from . import fsk as _fsk
    return _fsk.demod_fsk(samples, baud, mark_freq, space_freq, samp_rate)


def psk8_demodulate(s, b=1200, c=3000.0, s_r=96000):
    """modem.py:348: 8PSK is an alias of DQPSK; the parameter NAMES are part of the contract
    (decoder.py:334 calls it with baud=/carrier= and gets a TypeError)."""
    return qpsk_demodulate(s, b, c, s_r)


def fsk_high_speed_demodulate(s, baud=19200, s_r=96000):
    """modem.py:355-356."""
    return fsk_demodulate(s, baud, 8000, 16000, s_r)


def ofdm_demodulate_simple(s, baud, carrier, num_subcarriers, samp_rate=96000):
    """modem.py:375-376: OFDM4/8 are aliases of DQPSK; num_subcarriers is ignored."""
    return qpsk_demodulate(s, baud, carrier, samp_rate)


def ft8_demodulate(s, b, c, sr=96000):
    """modem.py:391."""
    return fsk_demodulate(s, 50, c, c + 50, sr)


def psk31_demodulate(s, b, c, sr=96000):
    """modem.py:397."""
    return bpsk_demodulate(s, 31.25, c, sr)


# ---- TX side: same names / defaults as modem.py:28,138,270,344,351,371,379 -----------------------------
def bpsk_modulate(data_bytes: bytes, baud=1200, carrier=3000.0, samp_rate=96000):
    """DBPSK, modem.py:28-65 (preamble [1,0]*40; 1 -> phase += pi; 10 % linear edge ramps)."""
    from . import modulate as _m
    return _m.modulate_batch([data_bytes], *_m.psk_mod_params(_m.FB_MOD_DBPSK, baud, carrier, samp_rate))[0]


def qpsk_modulate(data_bytes: bytes, baud=1200, carrier=3000.0, samp_rate=96000):
    """DQPSK, modem.py:138-186 (preamble [0,0]*30+[1,1]*10; 00 -> 0, 01 -> +pi/2, 11 -> pi, 10 -> -pi/2)."""
    from . import modulate as _m
    return _m.modulate_batch([data_bytes], *_m.psk_mod_params(_m.FB_MOD_DQPSK, baud, carrier, samp_rate))[0]


def fsk_modulate(data_bytes: bytes, baud=1200, mark_freq=1200.0, space_freq=2200.0, samp_rate=96000):
    """CPFSK, modem.py:270-295 (preamble AA AA AA AA, phase carried across bits modulo 2 pi, x 0.9)."""
    from . import modulate as _m
    return _m.modulate_batch([data_bytes], *_m.fsk_mod_params(baud, mark_freq, space_freq, samp_rate))[0]


def psk8_modulate(d, b=1200, c=3000.0, s=96000):
    """modem.py:344."""
    return qpsk_modulate(d, b, c, s)


def fsk_high_speed_modulate(d, baud=19200, s=96000):
    """modem.py:351-352."""
    return fsk_modulate(d, baud, 8000, 16000, s)


def ofdm_modulate_simple(d, baud, carrier, num_subcarriers, samp_rate=96000):
    """modem.py:371-372."""
    return qpsk_modulate(d, baud, carrier, samp_rate)


def apsk16_modulate(d, b, c, s=96000):
    """modem.py:379."""
    return qpsk_modulate(d, b, c, s)
