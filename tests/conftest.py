import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the B200 box")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a machine without a CUDA device; on a GPU box nothing is skipped and
    a missing CUDA extension fails loudly inside the tests."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA (sm_100a) device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "kat_small.npz"))


@pytest.fixture(scope="session")
def golden_2b():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "kat_2b.npz"))


@pytest.fixture(scope="session")
def golden_9b():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "kat_9b.npz"))
