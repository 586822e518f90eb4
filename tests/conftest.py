import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def big_fq_bytes():
    import gzip
    import numpy as np
    with gzip.open(os.path.join(GOLDEN, "big.fq.gz")) as f:
        return np.frombuffer(f.read(), dtype=np.uint8)


@pytest.fixture(scope="session")
def big_fq_path():
    return os.path.join(GOLDEN, "big.fq.gz")
