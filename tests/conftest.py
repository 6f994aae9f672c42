import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    import galax_b200

    galax_b200.build()
    from galax_b200 import _lib

    return _lib.lib()


def synthetic_ics(opot, N, seed, rmin=4.0, rmax=20.0):
    """r ~ U(rmin,rmax), isotropic direction, |v| = v_c(r) U(0.6,1), isotropic direction."""
    from oracle import potentials as op

    rng = np.random.default_rng(seed)
    r = rng.uniform(rmin, rmax, N)

    def iso(n):
        v = rng.normal(size=(n, 3))
        return v / np.linalg.norm(v, axis=1, keepdims=True)

    q = iso(N) * r[:, None]
    vc = op.circular_velocity(opot, r)
    p = iso(N) * (vc * rng.uniform(0.6, 1.0, N))[:, None]
    return q, p
