import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the shared libraries once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    yield


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle()


def has_gpu():
    from mpc_ros_b200 import capi
    return capi.lib().mpc_b200_device_count() > 0
