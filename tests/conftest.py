import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instant-ngp-pp_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _cuda_ready():
    """torch.cuda.is_available(), retried for a few seconds when a GPU device node exists: a freshly
    leased box can refuse the very first driver initialisation."""
    import time
    import torch
    if torch.cuda.is_available():
        return True
    if not os.path.exists("/dev/nvidia0"):
        return False
    for _ in range(6):
        time.sleep(3)
        try:
            torch.cuda.init()
            return True
        except Exception:
            pass
    return torch.cuda.is_available()


def pytest_collection_modifyitems(config, items):
    if _cuda_ready():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
