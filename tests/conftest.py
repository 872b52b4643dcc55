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


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device and the built library; on a box without them they are skipped (the CPU suite
    is selected with -m "not gpu"; a plain `pytest tests` must not fail there)."""
    try:
        import torch
        have = torch.cuda.is_available() and os.path.exists(os.path.join(ROOT, "image-text-retrieval_b200", "libitr_b200.so"))
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and libitr_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def bits_to_f32(bits):
    """bf16 bit patterns (uint16) -> exact float32 values."""
    return (bits.astype(np.uint32) << 16).view(np.float32)


@pytest.fixture(scope="session")
def golden():
    return load_golden
