import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests are skipped automatically where there is no GPU (this container)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def golden_dir():
    import pathlib
    return pathlib.Path(GOLDEN)



@pytest.fixture(autouse=True)
def _zero_draw_offset():
    """The device-side draw offset (noise.set_draw_offset) is process-wide state: every test starts from offset 0."""
    yield
    try:
        from qbn_b200 import noise
        for t in noise._draw_base.values():
            t.zero_()
    except Exception:
        pass
