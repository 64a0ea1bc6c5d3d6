"""CPU: the factorised v2 FSK evaluation (tests/model_fsk_fast.py) against the oracle's fsk_demodulate stages:
local analytic FIR + edge term + circular-Hilbert far field reproduce every vote of the reference chain."""
import numpy as np
import pytest

from oracle import modem_v2 as o2, signals as sig

import model_fsk_fast as mf


@pytest.mark.parametrize("baud,mark,space,nbytes,snr,seed", [(9600, 12000.0, 24000.0, 1200, 20, 1), (4800, 8000.0, 16000.0, 500, 12, 2)])
def test_factorised_fsk_matches_oracle_votes(baud, mark, space, nbytes, snr, seed):
    rng = np.random.default_rng(seed)
    x = sig.add_awgn(sig.fsk_modulate(bytes(rng.integers(0, 256, nbytes, dtype=np.uint8)), baud=baud, mark_freq=mark, space_freq=space), snr, rng)
    x = np.asarray(x, dtype=np.float64)
    if len(x) % 2:
        x = x[:-1]
    st = o2.fsk_stages(x, baud, mark, space)
    bits, env2 = mf.fsk_fast_bits(x, baud, mark, space)
    # the oracle's decided bit stream: rebuild it from its stages the way fsk_stages does
    from scipy import signal
    nyq = 48000.0
    envs = []
    for f in (mark, space):
        b, a = signal.butter(3, [(f - baud) / nyq, (f + baud) / nyq], btype="band")
        envs.append(np.abs(signal.hilbert(signal.filtfilt(b, a, x))))
    spb = 96000 // baud
    q = spb // 4
    centres = np.arange(spb // 2, len(x), spb)
    want = np.array([1 if np.mean((envs[0] > envs[1])[c - q: c + q]) > 0.5 else 0 for c in centres], dtype=np.uint8)
    assert np.array_equal(bits, want)
    # and the envelopes themselves, at every vote-window sample, to 1e-9 of the peak
    n = np.concatenate([np.arange(c - q, min(c + q, len(x))) for c in centres])
    peak = max(envs[0].max(), envs[1].max()) ** 2
    assert np.abs(env2[0] - envs[0][n] ** 2).max() < 1e-9 * peak
    assert np.abs(env2[1] - envs[1][n] ** 2).max() < 1e-9 * peak
    assert len(st["raw"]) > 0
