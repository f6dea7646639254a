import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref) -- the checker, never the product."""
    from oracle import refdrv
    if not refdrv.available():
        pytest.skip("oracle/_ref not built (make -C oracle ref needs /root/reference)")
    return refdrv.get(False)


@pytest.fixture(scope="session")
def reff():
    from oracle import refdrv
    if not refdrv.available(True):
        pytest.skip("oracle/_ref (float) not built")
    return refdrv.get(True)
