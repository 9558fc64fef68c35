import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (oracle) and make sure the product library exists; neither needs a GPU."""
    from oracle import oracle as orc
    orc.build()
    import psi_b200
    if not os.path.exists(psi_b200.lib_path()):
        psi_b200.build_library()
    yield
