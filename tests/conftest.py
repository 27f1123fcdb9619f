import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def nh():
    """The product's host mirror of module nhydro with a clean parameter state."""
    from mgroms_b200 import nhydro
    nhydro.load()
    nhydro.reset_params()
    yield nhydro
    try:
        nhydro.nhydro_clean()
    except Exception:
        pass
    nhydro.reset_params()
