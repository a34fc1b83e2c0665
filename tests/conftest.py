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


# GPU tests written after the round's GPU budget ran out (they have passed only with the oracle standing in for the device
# calls): run them LAST, so that under `-x` a surprise there cannot mask the suite that HAS been validated on a B200.
_NOT_YET_RUN_ON_A_GPU = ("test_certificate_solver.py", "test_gpu_errors.py", "test_xm2.py")


def pytest_collection_modifyitems(config, items):
    late = [it for it in items if it.get_closest_marker("gpu") and os.path.basename(str(it.fspath)) in _NOT_YET_RUN_ON_A_GPU]
    if late:
        ids = {id(it) for it in late}
        items[:] = [it for it in items if id(it) not in ids] + late


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
