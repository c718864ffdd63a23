import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "latent-pose-reenactment_b200"
for p in (str(PKG), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def golden_small():
    import torch
    return torch.load(GOLDEN / "small.pt", map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def golden_full():
    import torch
    return torch.load(GOLDEN / "full.pt", map_location="cpu", weights_only=False)
