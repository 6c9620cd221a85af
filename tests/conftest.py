import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ws():
    import wowsr_b200
    return wowsr_b200


@pytest.fixture(scope="session")
def handle(ws):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return ws.Handle(0)


def image_like(h, w, seed=3):
    """Seeded image-like RGB test input (blurred noise, +40 on G so the green mask fires)."""
    import cv2
    import numpy as np
    r = np.random.default_rng(seed)
    a = r.integers(0, 256, (h, w, 3)).astype(np.float32)
    a = cv2.GaussianBlur(a, (0, 0), 3)
    a[..., 1] += 40
    a = (a - a.min()) / (a.max() - a.min()) * 255
    return a.astype(np.uint8)
