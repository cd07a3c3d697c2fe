import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle (and the CUDA library, if nvcc is around) are built."""
    import __graft_entry__ as g
    g.build_oracle()
    if not os.path.exists(os.path.join(ROOT, "cubez_b200", "lib", "libcubezcuda.so")):
        g.build_cuda()
    yield
