import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # DS_TUNE="knob=value,..." runs the suite under non-default launch knobs (results must not change)
    if os.environ.get("DS_TUNE"):
        from dynamicscaler_b200 import lib
        for item in os.environ["DS_TUNE"].split(","):
            knob, value = item.split("=")
            lib.check(lib.ds_tune(int(knob), int(value)), "ds_tune")


@pytest.fixture(scope="session")
def small():
    return np.load(os.path.join(GOLDEN, "small.npz"))


@pytest.fixture(scope="session")
def configs():
    return np.load(os.path.join(GOLDEN, "configs.npz"))


def randn(seed, *shape):
    """Input generator shared with tests/golden/make_golden.py."""
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def none_if_neg(case):
    return [None if v < 0 else int(v) for v in case]
