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
    """The plain-C restatement (oracle/cloth_oracle.c), built on demand."""
    from oracle import oracle_py
    return oracle_py.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference behind oracle/ref_shim.cpp; None when it has not been built."""
    from oracle import oracle_py
    lib = oracle_py.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libclothref.so not built (needs /root/reference)")
    return lib
