import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def golden():
    """Reference-generated vectors (tests/golden/make_golden.py ran the unmodified reference classes)."""
    path = os.path.join(ROOT, "tests", "golden", "qdiff_golden.npz")
    z = np.load(path)
    cases = {}
    for key in z.files:
        name, field = key.rsplit("/", 1)
        cases.setdefault(name, {})[field] = z[key]
    return cases


@pytest.fixture(scope="session")
def golden_static():
    """Reference-generated vectors of the STATIC activation-scale path (tests/golden/make_golden_static.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "static_act_golden.npz"))
    cases = {}
    for key in z.files:
        name, field = key.rsplit("/", 1)
        cases.setdefault(name, {})[field] = z[key]
    return cases
