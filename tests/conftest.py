import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_dir() -> str:
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_cases():
    import glob

    import numpy as np

    out = {}
    for path in sorted(glob.glob(os.path.join(golden_dir(), "*.npz"))):
        out[os.path.splitext(os.path.basename(path))[0]] = dict(np.load(path))
    assert out, "no golden fixtures found"
    return out
