import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle on oracle/liboracle.so (built on demand with the committed Makefile)."""
    import helpers
    return helpers.load_oracle()


@pytest.fixture(scope="session")
def gpu():
    import torch  # noqa: F401  (only to give a clear skip reason)
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import naf_b200
    ctx = naf_b200.NafGpu(0)
    yield ctx
    ctx.close()
