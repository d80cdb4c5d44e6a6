import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) and, if missing, the CUDA library (nvcc cross-compiles without a GPU)."""
    from oracle import pt3d_oracle
    pt3d_oracle.build()
    lib = os.path.join(ROOT, "acfm_video_3d_reconstruction_b200", "libacfm_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
