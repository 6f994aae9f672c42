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


def rel_dev(a, b):
    """max(|dq|/|q|, |dp|/|p|) per particle (3-vector norms), over all saves: a, b = (q, p) with shape (N, T, 3)."""
    (qa, pa), (qb, pb) = a, b
    eq = np.linalg.norm(qa - qb, axis=-1) / np.linalg.norm(qb, axis=-1)
    ep = np.linalg.norm(pa - pb, axis=-1) / np.linalg.norm(pb, axis=-1)
    return np.maximum(eq, ep).max(axis=-1)


def one_ulp_sensitivity(pot, q0, p0, t0, t1, dt0, saveat=None, draws=4, seed=7):
    """How far the REFERENCE-ORDER result (GX_SCHEME_STRICT: bit-identical to oracle/galax_oracle.c, see
    tests/test_gpu_strict.py) moves when every component of a particle's initial condition is changed by one ulp
    (max over `draws` random sign patterns).  This is the floor under the deviation between any two implementations
    that round differently, per particle; the fast kernels are held to 100 x it (sqrt(10^4 steps) uncorrelated
    rounding-level perturbations) or 1e-12, whichever is larger.  Returns (sens[N], strict solution)."""
    import galax_b200.dynamics as gd

    strict = gd.OrbitSolver(solver=gd.SemiImplicitEuler(strict=True), stepsize_controller=gd.ConstantStepSize(),
                            max_steps=None)  # fmt: skip
    base = strict.solve(pot, (q0, p0), t0, t1, saveat=saveat, dt0=dt0)
    rng = np.random.default_rng(seed)

    def nudge(x):
        up = rng.integers(0, 2, size=x.shape).astype(bool)
        return np.where(up, np.nextafter(x, np.inf), np.nextafter(x, -np.inf))

    sens = np.zeros(len(q0))
    for _ in range(draws):
        sj = strict.solve(pot, (nudge(q0), nudge(p0)), t0, t1, saveat=saveat, dt0=dt0)
        sens = np.maximum(sens, rel_dev(sj.ys, base.ys))
    return sens, base
