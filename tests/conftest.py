import pathlib as pl
import sys

import pytest

ROOT = pl.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def drb():
    import scenarios

    return scenarios.load_drb()


@pytest.fixture(scope="session")
def drbx(drb):
    import scenarios

    p, f, start = drb
    px, fx = scenarios.drbx(p, f)
    return px, fx, start


@pytest.fixture(scope="session")
def sagehen():
    import scenarios

    return scenarios.load_sagehen()


@pytest.fixture(scope="session")
def hru1():
    import scenarios

    return scenarios.load_hru1()
