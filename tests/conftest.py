import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_bin(path):
    with open(path, "rb") as f:
        r = int.from_bytes(f.read(4), "little"); c = int.from_bytes(f.read(4), "little")
        return np.fromfile(f, dtype=np.float64, count=r * c).reshape((r, c), order="F")


@pytest.fixture(scope="session")
def simple1_q():
    return load_bin(os.path.join(GOLD, "simple1_Q.bin"))


@pytest.fixture(scope="session")
def simple2_q():
    return np.load(os.path.join(GOLD, "simple2_Q_ref.npz"))["Q"]


@pytest.fixture(scope="session")
def simple2_obs():
    return np.load(os.path.join(GOLD, "simple2_obs.npz"))


@pytest.fixture(scope="session")
def gpu_handle_factory():
    from xm_code_b200 import capi
    handles = []

    def make(**kw):
        h = capi.Handle(device=0, **kw)
        handles.append(h)
        return h
    yield make
    for h in handles:
        h.close()


def anchored_gram(R, s):
    """Gauge-invariant summary of a point: X = (sR)(sR)^T restricted to the first camera's block row."""
    sR = R * np.repeat(s, 3)[:, None]
    return sR @ sR[:3].T
