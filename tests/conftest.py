import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def checkpoints():
    """Shipped Upper/Lower state dicts (Resource/Pretrained_model layout) on CPU."""
    import torch
    base = os.path.join(ROOT, "Resource", "Pretrained_model")
    up = torch.load(os.path.join(base, "Upper_Net", "epoch451_batch20frame20lr3e-05.pth"), map_location="cpu",
                    weights_only=True)
    lo = torch.load(os.path.join(base, "Lower_Net", "epoch161_batch20frame20lr0.0003.pth"), map_location="cpu",
                    weights_only=True)
    return up, lo
