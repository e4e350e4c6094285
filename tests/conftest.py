import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def random_boxes(rng, b, k, spread=2.0, degenerate=False):
    """(b,k,8,3) corner boxes in the reference's corner order (model.py:108-110): random size / yaw / centre."""
    l = rng.uniform(0.3, 2.0, (b, k))
    w = rng.uniform(0.3, 2.0, (b, k))
    h = rng.uniform(0.3, 2.0, (b, k))
    yaw = rng.uniform(0, 2 * np.pi, (b, k))
    if degenerate:  # axis-aligned, shared sizes, shared edges: stresses parallel / on-boundary cases
        yaw = rng.integers(0, 4, (b, k)) * (np.pi / 2)
        l = rng.choice([0.5, 1.0], (b, k)); w = rng.choice([0.5, 1.0], (b, k)); h = rng.choice([0.5, 1.0], (b, k))
    c = rng.uniform(-spread, spread, (b, k, 3))
    if degenerate:
        c = np.round(c * 2) / 2
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1]); sy = np.array([1, 1, 1, 1, -1, -1, -1, -1]); sz = np.array([1, -1, -1, 1, 1, -1, -1, 1])
    x = sx * l[..., None] / 2; y = sy * h[..., None] / 2; z = sz * w[..., None] / 2
    cs, sn = np.cos(yaw)[..., None], np.sin(yaw)[..., None]
    X = cs * x + sn * z + c[..., 0:1]
    Y = y + c[..., 1:2]
    Z = -sn * x + cs * z + c[..., 2:3]
    return np.stack([X, Y, Z], -1).astype(np.float32)
