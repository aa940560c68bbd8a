import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, "oracle")):
    if path not in sys.path:
        sys.path.insert(0, path)

import lscpm_b200  # noqa: E402

lscpm_b200.install()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native():
    from lscpm_b200 import _native

    return _native


@pytest.fixture(scope="session")
def built_library():
    """The CUDA library must exist in-tree (built by __graft_entry__.build())."""
    import __graft_entry__ as entry

    entry.build()
    from lscpm_b200 import _native

    return _native.load()
