import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "u-llava_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100a); run with -m gpu")
    config.addinivalue_line("markers", "slow: CPU test that takes more than ~10 s")


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a CPU box: they are deselected by -m "not gpu" and
    # fail loudly (no skip) if someone runs them without a device.
    pass


@pytest.fixture(scope="session")
def ctx():
    import torch
    import native
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no CPU fallback exists)"
    return native.Context.get(0)
