import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_everything():
    """Build what the tests load if it is not there yet (a fresh clone has no .so files): the CUDA library with
    nvcc (cross-compiles without a GPU) and the C oracle with gcc."""
    import __graft_entry__ as entry

    entry.build()
