import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def pkg():
    """The product package (its name starts with a digit, so it is imported by string)."""
    return importlib.import_module("2dto3dreconstruction_b200")


def sub(name):
    return importlib.import_module("2dto3dreconstruction_b200." + name)


@pytest.fixture(scope="session")
def golden():
    return load_golden
