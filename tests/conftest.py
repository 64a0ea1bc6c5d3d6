import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "audio-modem-radio_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    meta = json.load(open(os.path.join(GOLD, "demod_cases.json")))
    arrs = np.load(os.path.join(GOLD, "demod_cases.npz"))
    return meta, arrs


@pytest.fixture(scope="session")
def engine():
    import fbdsp
    return fbdsp.Engine(0)
