import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu through gpurun)")


@pytest.fixture(scope="session")
def cuda_lib():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rcot_b200 import _lib

    lib = _lib.lib()  # raises loudly when the extension was not built
    _lib.check(lib.rcot_check_device(), "check_device")
    return lib
