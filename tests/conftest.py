import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests/` on a machine without a device skips the gpu-marked tests instead of
    failing them (the product itself has no CPU fallback: wc_create returns WC_ERR_NO_DEVICE)."""
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked tests run on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding

    binding.lib()
    return binding
