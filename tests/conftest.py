import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_small.npz"))


@pytest.fixture(scope="session")
def golden_hashes():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_hashes.json")) as f:
        return json.load(f)
