import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The self-contained restatement (oracle/liboracle.so), built on demand."""
    from oracle.oracle import Oracle, build
    build("orc")
    return Oracle("orc")


@pytest.fixture(scope="session")
def asref():
    """The reference's own headers (oracle/_ref/libasref.so).  Rebuilt where /root/reference
    exists, else the prebuilt library that travelled with the snapshot; skipped if neither."""
    from oracle.oracle import Oracle, available, build
    try:
        build("asref")
    except Exception:
        pass
    if not available("asref"):
        pytest.skip("oracle/_ref/libasref.so not available")
    return Oracle("asref")
