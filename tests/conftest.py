import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _guard_zones():
    """With DDB_GUARD=1 every device allocation of the library carries guard zones: verify them after each test
    (the pool's compute-sanitizer is closed; this catches out-of-bounds writes of the kernels)."""
    yield
    if os.environ.get("DDB_GUARD", "0") not in ("", "0"):
        from ddb200 import _lib

        bad = _lib.load().ddb_debug_check_guards()
        assert bad == 0, _lib.load().ddb_last_error().decode()
