"""GPU: the CUDA path against the committed oracle fixtures (no oracle call, no compiler needed at run time)."""
from pathlib import Path

import numpy as np
import pytest

import galax_b200.dynamics as gd
import galax_b200.potential as gp
from conftest import one_ulp_sensitivity, rel_dev

pytestmark = pytest.mark.gpu
FX = np.load(Path(__file__).parent / "golden" / "oracle_fixtures.npz")
MODELS = {"MilkyWayPotential": gp.MilkyWayPotential, "MilkyWayPotential2022": gp.MilkyWayPotential2022,
          "BovyMWPotential2014": gp.BovyMWPotential2014}  # fmt: skip


@pytest.mark.parametrize("name", list(MODELS))
def test_potential_against_fixture(name):
    pot = MODELS[name]()
    xyz = FX["pot_xyz"]
    g = FX[f"pot_{name}_grad"]
    assert (np.abs(pot.gradient(xyz) - g) / np.linalg.norm(g, axis=1, keepdims=True)).max() < 3e-15
    assert np.abs(pot.potential(xyz) / FX[f"pot_{name}_phi"] - 1).max() < 3e-15
    H = FX[f"pot_{name}_hess"]
    assert (np.abs(pot.hessian(xyz) - H) / np.abs(H).max(axis=(1, 2), keepdims=True)).max() < 2e-13


@pytest.mark.parametrize("name", list(MODELS))
def test_fixed_step_against_fixture(name):
    """64 orbits, 10^4 SemiImplicitEuler steps, saves at 0 / 333.3 / 1000 Myr.  The reference-order kernel reproduces
    the committed oracle output BIT FOR BIT; the fast kernel is within 1e-12, or 100 x the orbit's own one-ulp
    sensitivity where that is larger (tests/test_gpu_strict.py explains the bound)."""
    pot = MODELS[name]()
    q0, p0, ts = FX[f"sie_{name}_q0"], FX[f"sie_{name}_p0"], FX[f"sie_{name}_ts"]
    ref = (FX[f"sie_{name}_q"], FX[f"sie_{name}_p"])
    sens, strict = one_ulp_sensitivity(pot, q0, p0, 0.0, 1000.0, 0.1, saveat=ts)
    assert np.array_equal(strict.ys[0], ref[0]) and np.array_equal(strict.ys[1], ref[1])
    solver = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)
    sol = solver.solve(pot, (q0, p0), 0.0, 1000.0, saveat=ts, dt0=0.1)
    e = rel_dev(sol.ys, ref)
    assert (e <= np.maximum(1e-12, 100.0 * sens)).all() and np.median(e) <= 1e-13
    assert np.array_equal(sol.ys[0][:, 0], ref[0][:, 0])


def test_dopri8_against_fixture():
    """The reference-order kernel reproduces the committed oracle output bit for bit (saves and step counts); the fast
    kernel's typical particle is inside the north_star bar and it takes statistically the same number of steps."""
    pot = gp.MilkyWayPotential2022()
    ctl = gd.PIDController(rtol=1e-10, atol=1e-10)
    strict = gd.OrbitSolver(solver=gd.Dopri8(strict=True), stepsize_controller=ctl).solve(
        pot, (FX["dp8_q0"], FX["dp8_p0"]), 0.0, 200.0, saveat=FX["dp8_ts"], dt0=1.0)
    assert np.array_equal(strict.ys[0], FX["dp8_q"]) and np.array_equal(strict.ys[1], FX["dp8_p"])
    assert np.array_equal(np.asarray(strict.stats["num_steps"]), FX["dp8_ntot"])
    sol = gd.OrbitSolver(stepsize_controller=ctl).solve(pot, (FX["dp8_q0"], FX["dp8_p0"]), 0.0, 200.0, saveat=FX["dp8_ts"], dt0=1.0)
    d = np.abs(sol.ys[0] - FX["dp8_q"]) / (1e-10 + 1e-10 * np.abs(FX["dp8_q"]))
    assert np.median(d.max(axis=(1, 2))) <= 10.0
    assert abs(int(sol.stats["num_steps"].sum()) / int(FX["dp8_ntot"].sum()) - 1) < 0.02


@pytest.mark.parametrize("df", ["fardal", "chen"])
def test_release_against_fixture(df):
    pot = gp.MilkyWayPotential()
    orbit = gd.Orbit(FX["rel_xq"], FX["rel_xp"], np.zeros(len(FX["rel_xq"])))
    if df == "fardal":
        out = gd.FardalStreamDF().sample(FX["rel_normals"], pot, orbit, 1e4)
    else:
        out = gd.ChenStreamDF().sample(FX["rel_posvel"], pot, orbit, 1e4)
    for got, key in ((out["lead"].q, "ql"), (out["lead"].p, "pl"), (out["trail"].q, "qt"), (out["trail"].p, "pt")):
        ref = FX[f"rel_{df}_{key}"]
        assert (np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)).max() < 1e-12
