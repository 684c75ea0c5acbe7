import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "unet_sampler_golden.npz"))


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build (or reuse) the in-tree sm_100a library once per session; nvcc cross-compiles without a GPU."""
    from ivid_b200 import build
    build.build()
