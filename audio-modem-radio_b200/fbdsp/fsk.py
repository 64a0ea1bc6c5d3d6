"""v2 FSK host side (modem.py:298-341): Butterworth design with the reference's exact error behaviour.
The device path for valid tone sets is built in csrc/fsk_v2.cu."""
from __future__ import annotations

from scipy import signal


def fsk_design_check(baud, mark_freq, space_freq, samp_rate):
    """Runs the same scipy design calls as get_envelope (modem.py:306-307): raises the reference's
    ValueError for every product default (f - baud <= 0)."""
    nyq = samp_rate / 2
    out = []
    for freq in (mark_freq, space_freq):
        out.append(signal.butter(3, [(freq - baud) / nyq, (freq + baud) / nyq], btype="band"))
    return out


def demod_fsk(samples, baud, mark_freq, space_freq, samp_rate):
    fsk_design_check(baud, mark_freq, space_freq, samp_rate)
    from ._lib import FbdspError
    raise FbdspError("v2 FSK device kernel not built yet for valid tone sets (no CPU fallback by design)")
