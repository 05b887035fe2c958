import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The reference's C templates compiled by oracle/Makefile (oracle/_ref)."""
    import refdsp
    return refdsp.RefDSP()


@pytest.fixture(scope="session")
def cuda():
    """DSP tables filled by libdav1d_cuda.so (the product). Fails loudly when absent."""
    import cudadsp
    return cudadsp.CudaDSP()
