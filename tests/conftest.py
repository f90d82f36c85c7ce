import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """Build (if stale) and load the C-ABI library; works without a GPU (nvcc cross-compiles)."""
    import __graft_entry__ as ge
    ge.build()
    from psnerf_b200 import _binding
    return _binding.load()
