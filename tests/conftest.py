import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of libvamp_b200.so, building it with nvcc if it is not there yet."""
    from vamp_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.LIB_PATH


@pytest.fixture(scope="session")
def ctx(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vamp_b200 import engine
    return engine.default_context()


@pytest.fixture(scope="session")
def golden_profiles():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "profiles.npz"))
