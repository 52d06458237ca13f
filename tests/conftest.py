import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tb():
    import tacotron_b200
    return tacotron_b200


@pytest.fixture(scope="session")
def hp5(tb):
    return tb.hparams.override(reduction_factor=5)
