import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
# The parity tests run on deterministic random-init weights of the named architectures (no checkpoints, no network):
# the explicit opt-in the product requires (saber_b200/pretrained_weights.py). test_cabi checks the default refusal.
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu under gpurun")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when collected without a device and without an explicit -m gpu.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
