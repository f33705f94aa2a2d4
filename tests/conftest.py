import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain ``pytest tests`` on a machine without a CUDA device skips the ``gpu`` tests instead of erroring
    (set FCVM_REQUIRE_GPU=1 to turn the skip back into a failure, e.g. on a box that must have one)."""
    if os.environ.get("FCVM_REQUIRE_GPU") or _have_gpu():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (the product path has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import fcvm_oracle
    fcvm_oracle.build()
    return fcvm_oracle
