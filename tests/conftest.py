import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist before any test touches the package (never built implicitly by the product)."""
    from flash_attention_dlrs_b200 import _lib

    if not _lib.LIB_PATH.exists():
        _lib.build()
    if not os.environ.get("FA_B200_LIB"):
        _lib.build_torch_binding()   # the C++ autograd node over the C ABI (no-op when up to date)
    return _lib.LIB_PATH
