import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _sm100_available() -> bool:
    try:
        import torch

        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a host without an sm_100 device, so that a plain `pytest` run tells a
    missing GPU apart from a broken build."""
    if _sm100_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA sm_100a device (B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
