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


# The loop-back multi-GPU tests (two cooperative kernels co-resident on one GPU) and the tests that page in cuSOLVER / cuBLAS run
# LAST: under `-x` a surprise there cannot mask the single-GPU parity suite.
_RUN_LATE = ("test_gpu_multi.py", "test_reference_scripts.py")


def pytest_collection_modifyitems(config, items):
    late = [it for it in items if it.get_closest_marker("gpu") and os.path.basename(str(it.fspath)) in _RUN_LATE]
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
