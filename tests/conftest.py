import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The C restatement (oracle/chad_oracle.c), built on demand. Test infrastructure only."""
    from oracle import bindings
    bindings.build("oracle")
    return bindings


@pytest.fixture(scope="session")
def chad_lib():
    """The product library; built on demand (nvcc cross-compiles without a GPU)."""
    from chad_tsdf_b200 import build, capi
    build.build_library()
    return capi.load()
