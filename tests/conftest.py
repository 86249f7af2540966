import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import OracleLib, build
    build()
    return OracleLib()


@pytest.fixture(scope="session")
def reflib():
    from oracle.pyoracle import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref/libpuffinn_ref.so not built (needs /root/reference at build time)")
    return RefLib()


@pytest.fixture(scope="session")
def lib():
    """The product library; fails loudly when missing (no CPU fallback)."""
    from clann_b200 import _lib
    return _lib.load()
