import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native library and the CPU checker once per session (no GPU needed)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def has_gpu():
    import rfsi_b200
    return rfsi_b200.device_count() > 0


def pytest_collection_modifyitems(config, items):
    # -m gpu tests must fail loudly, not skip, when no device is there: the driver only
    # selects them on a GPU box.  Nothing to do here; kept for clarity.
    return
