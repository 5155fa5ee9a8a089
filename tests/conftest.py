import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    """libgq.so must exist (built in-tree by __graft_entry__.build()); GPU tests never build it."""
    from gramtools_b200 import lib_path
    if not os.path.exists(lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    return lib_path()
