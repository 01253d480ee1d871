import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled headless (oracle/_ref). Skips if it was not built."""
    from oracle import ref_binding as rb
    if not rb.ref_available():
        pytest.skip("oracle/_ref/libcutesdr_ref.so not built (needs /root/reference)")
    return rb


@pytest.fixture(scope="session")
def refbig():
    from oracle import ref_binding as rb
    if not rb.ref_available(big=True):
        pytest.skip("oracle/_ref/libcutesdr_ref_big.so not built")
    return rb


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle_binding as ob
    ob.load()
    return ob
