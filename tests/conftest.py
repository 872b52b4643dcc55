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
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def bits_to_f32(bits):
    """bf16 bit patterns (uint16) -> exact float32 values."""
    return (bits.astype(np.uint32) << 16).view(np.float32)


@pytest.fixture(scope="session")
def golden():
    return load_golden
