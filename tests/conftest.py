import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _clear_device_error(request):
    """tcgen05 kernels flag a timed-out mbarrier wait in a device word instead of hanging; every GPU
    test must end with that word clear."""
    yield
    if "gpu" in request.keywords:
        import torch

        if torch.cuda.is_available():
            from weathermodel_b200 import ops

            code = ops.device_error()
            assert code == 0, f"device-side mbarrier timeout, wait site code {code}"
