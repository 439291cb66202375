import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def wfx():
    import wave_fenics_b200
    return wave_fenics_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    return oracle
