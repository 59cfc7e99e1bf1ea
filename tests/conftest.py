import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pth in (ROOT, os.path.join(ROOT, "tests")):
    if pth not in sys.path:
        sys.path.insert(0, pth)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def engine():
    from sparselm_b200.engine import get_engine

    return get_engine()
