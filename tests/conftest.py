import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


# GPU tests written after this round's GPU budget was spent: they have compiled but never run on a device.  They are
# skipped unless asked for, so a first-run failure cannot mask the verified suite behind `pytest -x`.
pending_first_gpu_run = pytest.mark.skipif(not os.environ.get("B200_RUN_UNVERIFIED"),
                                           reason="first GPU run pending: set B200_RUN_UNVERIFIED=1")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library (builds it if the .so is absent -- nvcc cross-compiles without a GPU)."""
    from lmdeploy_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()
