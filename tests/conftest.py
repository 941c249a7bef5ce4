import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_modules():
    import torch
    return torch.load(os.path.join(GOLDEN, "modules_tiny.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_step():
    import torch
    return torch.load(os.path.join(GOLDEN, "step_tiny.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_cascade():
    import torch
    return torch.load(os.path.join(GOLDEN, "cascade_tiny.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_cas_step():
    import torch
    return torch.load(os.path.join(GOLDEN, "cas_step_tiny.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_zoo():
    import torch
    return torch.load(os.path.join(GOLDEN, "zoo_tiny.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_step_gray():
    import torch
    return torch.load(os.path.join(GOLDEN, "step_gray_tiny.pt"), weights_only=False)
