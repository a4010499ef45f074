import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUTS = os.path.join(GOLDEN, "inputs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def romeo():
    return np.fromfile(os.path.join(INPUTS, "romeo.txt"), dtype=np.uint8)


@pytest.fixture(scope="session")
def jpeg():
    return np.fromfile(os.path.join(INPUTS, "pexels.jpg"), dtype=np.uint8)


@pytest.fixture(scope="session")
def codec():
    """the CUDA codec; on a GPU box a missing library is an error, never a skip"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from huffman_b200 import Codec
    c = Codec(0)
    yield c
    c.close()
