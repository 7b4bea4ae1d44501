import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "experimental: kernels written without GPU time left to verify them; "
                                       "they are not on any default path and their tests run only with "
                                       "PN2_EXPERIMENTAL=1 (first GPU call of the next round)")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("PN2_EXPERIMENTAL") == "1":
        return
    skip = pytest.mark.skip(reason="experimental kernel: set PN2_EXPERIMENTAL=1 to run")
    for item in items:
        if "experimental" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
