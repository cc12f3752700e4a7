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


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) when selected with -m gpu on a box without a GPU or
    # without the built extension; on the CPU-only container they are simply deselected by -m "not gpu".
    pass


@pytest.fixture(scope="session")
def mitten_scene():
    from oracle import datasets as ds
    return ds.scene_from_snapshot(np.load(os.path.join(GOLDEN, "mitten_init.npz")))


def rel_rmse(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2)))
