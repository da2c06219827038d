import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a CUDA device skips the GPU tests instead of failing them."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree before any test touches it (nvcc cross-compiles without
    a GPU).  Where nvcc is missing an existing build is used as it is; the pure-oracle tests do not need the
    library at all, the ones that do fail loudly on their own (c2m_b200._lib has no fallback)."""
    from c2m_b200 import _build
    try:
        _build.build()
    except RuntimeError as e:
        if "nvcc not found" not in str(e):
            raise
