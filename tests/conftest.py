import glob
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = sorted(glob.glob(os.path.join(REPO, "tests", "golden", "*.pbfx")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def golden_ids():
    return [os.path.basename(f)[:-5] for f in GOLDEN]


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()
