import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


@pytest.fixture(scope='session')
def built():
    """Make sure the C-ABI library and the kernel cache exist."""
    import __graft_entry__ as g

    g.build_runtime()
    return g
