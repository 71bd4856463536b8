import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_cases(name):
    """tests/golden/<name>.npz stored as 'case/key' -> {case: {key: array}}."""
    import numpy as np
    flat = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    out = {}
    for k in flat.files:
        case, key = k.split("/", 1)
        out.setdefault(case, {})[key] = flat[k]
    return out


@pytest.fixture(scope="session")
def loss_cases():
    return load_cases("loss_cases.npz")
