import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import warp_cpu
    warp_cpu.build()
    return warp_cpu


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests fail (not skip) when it is missing."""
    from rnntransducer_b200 import _lib
    return _lib.load()
