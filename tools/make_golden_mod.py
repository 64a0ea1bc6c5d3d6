#!/usr/bin/env python
"""tests/golden/modulators.npz: float32 waveforms of the UNMODIFIED reference modulators (modem.py:28-65, 138-186,
270-295 and the aliases at :344,351,371) on seeded payloads, plus the exception they raise for sps < 10.
Build container only:  python tools/make_golden_mod.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import GOLD, import_reference  # noqa: E402


def main():
    modem, _, _, _ = import_reference()
    from oracle import signals as sig
    rng = np.random.default_rng(404)
    arrs, cases = {}, []

    def add(name, fn, nbytes, args):
        data = rng.integers(0, 256, nbytes, dtype=np.uint8).tobytes()
        y = getattr(modem, fn)(data, *args)
        assert y.dtype == np.float32
        o = {"bpsk_modulate": sig.bpsk_modulate, "qpsk_modulate": sig.qpsk_modulate, "fsk_modulate": sig.fsk_modulate}.get(fn)
        if o is not None:
            assert np.array_equal(o(data, *args), y), name          # the oracle generators stay pinned too
        arrs[name + "__data"] = np.frombuffer(data, np.uint8)
        arrs[name + "__y"] = y
        cases.append({"name": name, "fn": fn, "args": list(args)})

    add("qpsk_9600_c9600", "qpsk_modulate", 61, (9600, 9600.0))
    add("qpsk_9600_c3000", "qpsk_modulate", 61, (9600, 3000.0))
    add("qpsk_1200_c2400", "qpsk_modulate", 17, (1200, 2400.0))
    add("qpsk_empty", "qpsk_modulate", 0, (4800, 9600.0))
    add("bpsk_4800_c9600", "bpsk_modulate", 33, (4800, 9600.0))
    add("bpsk_9600_c3000", "bpsk_modulate", 61, (9600, 3000.0))
    add("bpsk_1200_default", "bpsk_modulate", 5, ())
    add("fsk_9600_m12000_s24000", "fsk_modulate", 61, (9600, 12000.0, 24000.0))
    add("fsk_1200_default", "fsk_modulate", 9, ())
    add("fsk_300_m1200_s2200", "fsk_modulate", 2, (300, 1200.0, 2200.0))
    add("fsk_hs_19200", "fsk_high_speed_modulate", 61, ())
    add("psk8_9600_c12000", "psk8_modulate", 61, (9600, 12000.0))
    add("ofdm_4800_c12000", "ofdm_modulate_simple", 29, (4800, 12000.0, 8))
    errors = []
    for fn, args in (("qpsk_modulate", (38400, 12000.0)), ("bpsk_modulate", (19200, 3000.0)), ("psk8_modulate", (12000, 12000.0))):
        try:
            getattr(modem, fn)(b"abc", *args)
            raise AssertionError("expected an exception")
        except ValueError as e:
            errors.append({"fn": fn, "args": list(args), "type": "ValueError", "msg": str(e)})
    np.savez_compressed(os.path.join(GOLD, "modulators.npz"), **arrs)
    with open(os.path.join(GOLD, "modulators.json"), "w") as f:
        json.dump({"numpy": np.__version__, "cases": cases, "errors": errors}, f, indent=1)
    print("wrote", len(cases), "modulator cases,", sum(a.nbytes for a in arrs.values()), "bytes;", errors)


if __name__ == "__main__":
    main()
