import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """Builds (if stale) and loads the C-ABI library.  No GPU is needed to build or load it."""
    from mmearth_train_b200 import build
    build.build()
    import mmearth_train_b200._native as nat
    return nat
