import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import portapi
    portapi.build()
    return portapi


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref/libref.so), if it was built."""
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref/libref.so not built")
    return refapi


@pytest.fixture(scope="session")
def ctx():
    import longfellow_zk_b200 as lf
    return lf.Context(0)
