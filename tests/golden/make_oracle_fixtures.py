"""Write tests/golden/oracle_*.npz from the CPU oracle on seeded inputs.

    python tests/golden/make_oracle_fixtures.py

The fixtures pin the oracle against silent drift (tests/test_oracle_fixtures.py, CPU) and give the GPU tests a
reference that needs no compiler on the GPU box (tests/test_gpu_fixtures.py).  They are outputs of THIS repo's
oracle, not of the reference (which cannot be run here); the oracle itself is pinned on the reference's KATs.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

from conftest import synthetic_ics  # noqa: E402
from oracle import cref  # noqa: E402
from oracle import potentials as op  # noqa: E402

MODELS = {"MilkyWayPotential": op.milky_way_potential, "MilkyWayPotential2022": op.milky_way_potential_2022,
          "BovyMWPotential2014": op.bovy_mw_potential_2014}  # fmt: skip


def main():
    out = {}
    rng = np.random.default_rng(2026)
    r = 10 ** rng.uniform(-1, 2, 256)
    d = rng.normal(size=(256, 3))
    xyz = d / np.linalg.norm(d, axis=1, keepdims=True) * r[:, None]
    out["pot_xyz"] = xyz
    for name, f in MODELS.items():
        pot = f()
        out[f"pot_{name}_phi"] = op.potential(pot, xyz)
        out[f"pot_{name}_grad"] = op.gradient(pot, xyz)
        out[f"pot_{name}_hess"] = op.hessian(pot, xyz)
        q0, p0 = synthetic_ics(pot, 64, seed=11, rmin=6.0)
        out[f"sie_{name}_q0"], out[f"sie_{name}_p0"] = q0, p0
        ts = np.array([0.0, 333.3, 1000.0])
        q, p, st, n = cref.integrate_fixed(pot, q0, p0, 0.0, 1000.0, 0.1, ts)
        assert (st == 0).all() and (n == 10000).all()
        out[f"sie_{name}_ts"], out[f"sie_{name}_q"], out[f"sie_{name}_p"] = ts, q, p
    pot = MODELS["MilkyWayPotential2022"]()
    q0, p0 = synthetic_ics(pot, 32, seed=12, rmin=6.0)
    ts = np.linspace(0.0, 200.0, 9)
    q, p, st, na, nt = cref.integrate_dopri8(pot, q0, p0, 0.0, 200.0, ts, rtol=1e-10, atol=1e-10, dt0=1.0)
    out.update(dp8_q0=q0, dp8_p0=p0, dp8_ts=ts, dp8_q=q, dp8_p=p, dp8_nacc=na, dp8_ntot=nt)
    pot = MODELS["MilkyWayPotential"]()
    M = 128
    xq, xp = synthetic_ics(pot, M, seed=13)
    normals = rng.standard_normal((4, M))
    posvel = rng.multivariate_normal([1.6, -30, 0, 1, 20, 0], np.diag([0.1225, 529, 144, 0, 400, 484.0]), size=M, method="svd")
    out.update(rel_xq=xq, rel_xp=xp, rel_normals=normals, rel_posvel=posvel)
    for nm, fn, dr in (("fardal", cref.release_fardal, normals), ("chen", cref.release_chen, posvel)):
        ql, pl, qt, pt = fn(pot, xq, xp, 1e4, dr)
        out.update({f"rel_{nm}_ql": ql, f"rel_{nm}_pl": pl, f"rel_{nm}_qt": qt, f"rel_{nm}_pt": pt})
    np.savez_compressed(HERE / "oracle_fixtures.npz", **out)
    print("wrote", HERE / "oracle_fixtures.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
