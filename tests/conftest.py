import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box")


@pytest.fixture(scope="session")
def tdr_lib():
    """Build (if needed) and load libtdr_sm100.so."""
    from textualdegremoval_b200.csrc.build import build
    from textualdegremoval_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        build()
    return lib.load()
