import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def assets():
    return load_golden("assets.npz")


@pytest.fixture(scope="session")
def g_ops():
    return load_golden("ops.npz")


@pytest.fixture(scope="session")
def g_noise():
    return load_golden("noise.npz")


@pytest.fixture(scope="session")
def g_den():
    return load_golden("denoiser.npz")


@pytest.fixture(scope="session")
def g_loops():
    return load_golden("loops.npz")


@pytest.fixture(scope="session")
def g_tv():
    return load_golden("tv.npz")


@pytest.fixture(scope="session")
def g_long():
    p = os.path.join(GOLDEN, "long.npz")
    if not os.path.exists(p):
        pytest.skip("long.npz not generated")
    return np.load(p, allow_pickle=False)


def weights_path(arch):
    return os.path.join(GOLDEN, "weights", arch + ".pdsw")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
