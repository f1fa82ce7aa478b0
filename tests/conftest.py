import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Make sure the oracle (gcc) and libbhr.so (nvcc, cross-compiles without a GPU) exist."""
    import oracle
    oracle.build()
    from black_hole_renderer_b200 import build as b
    b.build()            # returns at once unless a source is newer than the binary: never test a stale library
    yield
