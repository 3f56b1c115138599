"""pytest configuration: markers, paths and shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running full-size case")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
