"""CPU: the committed oracle fixtures are reproduced by the oracle as built here (drift guard)."""
from pathlib import Path

import numpy as np

from oracle import cref
from oracle import potentials as op

FX = np.load(Path(__file__).parent / "golden" / "oracle_fixtures.npz")
MODELS = {"MilkyWayPotential": op.milky_way_potential, "MilkyWayPotential2022": op.milky_way_potential_2022,
          "BovyMWPotential2014": op.bovy_mw_potential_2014}  # fmt: skip


def test_potential_fixtures():
    for name, f in MODELS.items():
        pot = f()
        assert np.allclose(op.gradient(pot, FX["pot_xyz"]), FX[f"pot_{name}_grad"], rtol=1e-14, atol=0)
        assert np.allclose(op.potential(pot, FX["pot_xyz"]), FX[f"pot_{name}_phi"], rtol=1e-14, atol=0)
        o = cref.potential_eval(pot, FX["pot_xyz"], ("hess",))
        assert np.allclose(o["hess"], FX[f"pot_{name}_hess"], rtol=1e-11, atol=1e-22)


def test_integrator_fixtures():
    pot = MODELS["MilkyWayPotential"]()
    q, p, st, n = cref.integrate_fixed(pot, FX["sie_MilkyWayPotential_q0"][:8], FX["sie_MilkyWayPotential_p0"][:8], 0.0,
                                       1000.0, 0.1, FX["sie_MilkyWayPotential_ts"])
    assert np.allclose(q, FX["sie_MilkyWayPotential_q"][:8], rtol=1e-12, atol=1e-12)
    pot = MODELS["MilkyWayPotential2022"]()
    q, p, st, na, nt = cref.integrate_dopri8(pot, FX["dp8_q0"], FX["dp8_p0"], 0.0, 200.0, FX["dp8_ts"], rtol=1e-10,
                                             atol=1e-10, dt0=1.0)
    assert np.mean(na == FX["dp8_nacc"]) > 0.9
    assert np.median(np.abs(q - FX["dp8_q"]).max(axis=(1, 2))) < 1e-10
