import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "superpixels-segmentation-gui-opencv_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def spx():
    """the product package (ctypes binding of libspx_b200.so)"""
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def synthmod():
    return importlib.import_module(PKG + ".synth")


@pytest.fixture(scope="session")
def O():
    """the CPU oracle (test infrastructure)"""
    from oracle import oracle
    oracle.lib()
    oracle.set_threads(min(16, os.cpu_count() or 1))
    return oracle


@pytest.fixture(scope="session")
def example_bgr():
    import cv2
    return cv2.imread(os.path.join(GOLDEN, "logo-absurdephoton-segmentation-image.png"), cv2.IMREAD_COLOR)


@pytest.fixture(scope="session")
def example_lab(example_bgr, O):
    return O.bgr2lab8(example_bgr)


def random_label_map(rng, h, w, cell=12, noise=0.03):
    """blocky label map with stray pixels: stresses small-component merging and contour chains"""
    gy, gx = (h + cell - 1) // cell, (w + cell - 1) // cell
    base = rng.integers(0, max(2, gy * gx // 3), (gy, gx)).astype(np.int32)
    lab = np.kron(base, np.ones((cell, cell), np.int32))[:h, :w].copy()
    m = rng.random((h, w)) < noise
    lab[m] = rng.integers(0, base.max() + 1, int(m.sum()))
    return np.ascontiguousarray(lab)
