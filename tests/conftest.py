import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def lsf():
    """The product package with the CUDA library built; fails loudly if that is impossible."""
    import levelsetfortran_b200 as pkg
    from levelsetfortran_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    _lib.lib()
    return pkg


def load_mesh(name):
    z = np.load(os.path.join(GOLDEN, f"{name}_mesh.npz"))
    return np.asfortranarray(z["surfX"].astype(np.float64)), np.asfortranarray(z["surfElem"].astype(np.int32))


def synth_field(shape, seed=0, noise=0.01, dx=0.05):
    """A smeared-sign-like field of an off-centre sphere plus noise: exercises both Godunov branches."""
    rng = np.random.default_rng(seed)
    nxp, nyp, nzp = shape
    x, y, z = np.meshgrid(np.arange(nxp), np.arange(nyp), np.arange(nzp), indexing="ij")
    r = np.sqrt((x - nxp / 2.1) ** 2 + (y - nyp / 1.9) ** 2 + (z - nzp / 2.2) ** 2) * dx - 0.3 * min(shape) * dx
    s = r / np.sqrt(r * r + dx * dx)
    return np.asfortranarray(s + noise * rng.standard_normal(shape))


def dist_field(shape, seed=0, noise=0.002, dx=0.05):
    """A noisy signed-distance-like field of an off-centre sphere: wide narrow band for min/max tests."""
    rng = np.random.default_rng(seed)
    nxp, nyp, nzp = shape
    x, y, z = np.meshgrid(np.arange(nxp), np.arange(nyp), np.arange(nzp), indexing="ij")
    r = np.sqrt((x - nxp / 2.1) ** 2 + (y - nyp / 1.9) ** 2 + (z - nzp / 2.2) ** 2) * dx - 0.3 * min(shape) * dx
    # keep the band off the grid boundary (the reference would read out of bounds there)
    r = np.maximum(r, -10.0)
    edge = np.ones(shape, dtype=bool)
    edge[1:-1, 1:-1, 1:-1] = False
    f = r + noise * rng.standard_normal(shape)
    f[edge & (np.abs(f) < 5 * dx)] = 5 * dx
    return np.asfortranarray(f)
