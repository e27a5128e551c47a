import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Make sure the in-tree CUDA library exists (nvcc cross-compiles on CPU)."""
    from splashy_b200 import _native
    if not _native.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    return _native.load()
