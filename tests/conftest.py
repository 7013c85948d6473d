import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def lib():
    """The C-ABI shared library (built on demand with nvcc; cross-compiles without a GPU)."""
    from radialog_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    from radialog_b200 import _lib
    if not _lib.load().rd_device_ok(0):
        pytest.fail("CUDA device 0 is not sm_100 (B200): " + _lib.load().rd_last_error().decode())
    return torch.device("cuda:0")
