import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """libglfer_b200.so, (re)built in-tree when sources are newer (nvcc cross-compiles on CPU)."""
    from glfer_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def api(built_lib):
    from glfer_b200 import api as _api
    return _api


@pytest.fixture(scope="session")
def gpu_api(api):
    if api.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests need a B200 (no CPU fallback exists)")
    return api
