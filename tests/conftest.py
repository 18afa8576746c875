import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Every native piece is built in-tree before any test runs (no-op when up to date)."""
    import __graft_entry__ as g
    g.build_kyd()
    g.build_host()
    g.build_cli()
    g.build_oracle()
    return True


@pytest.fixture(scope="session")
def device(built):
    import ky_b200 as ky
    dev = ky.Device(0)
    yield dev
    dev.close()
