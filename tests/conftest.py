import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand from oracle/hq_oracle.c."""
    from oracle import hq_oracle

    hq_oracle.build()
    hq_oracle.load()
    return hq_oracle


@pytest.fixture(scope="session")
def hqlib():
    """The product's C ABI.  Built on demand with nvcc (cross-compiles without a GPU)."""
    from hybridquantization_b200 import _lib, build

    build.build_library()
    return _lib.load()


@pytest.fixture(scope="session")
def backend(hqlib):
    """One GPU context for the whole session (gpu tests only)."""
    from hybridquantization_b200 import ImageManipulation

    be = ImageManipulation("CIE76", False, True, 0)
    yield be
    be.close()
