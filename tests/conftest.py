import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the B200 box")


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
