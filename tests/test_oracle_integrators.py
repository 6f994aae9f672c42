"""CPU: the integrator / stream oracle against the reference's doctest KATs and independent truths."""
import json
from pathlib import Path

import numpy as np
import pytest
from scipy.integrate import solve_ivp

from oracle import cref
from oracle import dopri8_tableau as tab
from oracle import potentials as op

from conftest import synthetic_ics

KATS = json.loads((Path(__file__).parent / "golden" / "orbit_kats.json").read_text())
KMS = KATS["kms"]


def _model(m):
    return op.single(op.KIND_HERNQUIST, *m["params"])


@pytest.mark.parametrize("case", KATS["cases"], ids=lambda c: c["name"])
def test_reference_orbit_doctests(case):
    pot = _model(case["model"])
    q0 = [case["q0"]]
    p0 = [[v * KMS for v in case["p0_kms"]]]
    if case["solver"] == "SemiImplicitEuler":
        q, p, st, n = cref.integrate_fixed(pot, q0, p0, case["t0"], case["t1"], case["dt0"], case["ts"])
    else:
        q, p, st, na, nt = cref.integrate_dopri8(pot, q0, p0, case["t0"], case["t1"], case["ts"],
                                                 rtol=case["rtol"], atol=case["atol_solver"])
    assert st[0] == cref.OK
    assert np.allclose(q[0], case["q"], atol=case["atol"], rtol=0)
    assert np.allclose(p[0], case["p"], atol=case["atol"], rtol=0)


def test_tableau_order_conditions():
    r = tab.verify()
    assert r["order8_sol"] < 5e-15 and r["order7_hat"] < 5e-15 and r["order8_hat"] > 1e-6


def test_dense_output_defining_properties():
    """b(0)=0, b(1)=b_sol, b'(0)=e_1, b'(1)=e_14, order <= 5 conditions at several theta."""
    B = tab.dense_b()
    assert np.abs(tab.dense_weights(1.0) - tab.b_sol()).max() < 1e-11
    d1 = B @ np.arange(1, 7)  # derivative at theta=1
    e14 = np.zeros(14)
    e14[13] = 1
    assert np.abs(d1 - e14).max() < 1e-10
    assert B[0, 0] == 1.0 and np.all(B[1:, 0] == 0.0)
    for th in (0.1, 0.37, 0.5, 0.9):
        res = tab.order_residuals(tab.dense_weights(th), 5, theta=th)
        assert max(abs(r) for _, r in res) < 2e-12
        res6 = [r for p, r in tab.order_residuals(tab.dense_weights(th), 6, theta=th) if p == 6]
        # all order-6 conditions but one (the tall tree) hold
        assert sorted(abs(r) for r in res6)[-2] < 2e-11


def test_dense_output_accuracy_matches_step_accuracy():
    kep = op.single(op.KIND_HERNQUIST, 1e12, 5.0)
    y0 = [8.0, 0, 0, 0, 220 * KMS, 0]
    ts = np.linspace(0, 1000, 201)

    def f(t, y):
        return np.concatenate([y[3:], -op.gradient(kep, y[:3])])

    truth = solve_ivp(f, (0, 1000), y0, method="DOP853", rtol=1e-13, atol=1e-13, t_eval=ts).y.T
    q, p, st, na, nt = cref.integrate_dopri8(kep, [y0[:3]], [y0[3:]], 0.0, 1000.0, ts, rtol=1e-8, atol=1e-8)
    err = np.abs(np.concatenate([q[0], p[0]], axis=1) - truth).max(axis=1)
    assert err[0] == 0.0  # ts[0] == t0 returns y0
    assert err.max() < 5e-6  # interpolated points are as good as step end points (~2e-6)


def test_dopri8_converges_to_dop853_truth():
    pot = op.milky_way_potential_2022()
    q0, p0 = synthetic_ics(pot, 8, seed=2)
    ts = np.linspace(0, 500, 6)
    for tol, bound in ((1e-8, 3e-5), (1e-11, 1e-7)):
        q, p, st, na, nt = cref.integrate_dopri8(pot, q0, p0, 0.0, 500.0, ts, rtol=tol, atol=tol)
        assert (st == 0).all()
        for i in range(8):
            def f(t, y):
                return np.concatenate([y[3:], -op.gradient(pot, y[:3])])
            s = solve_ivp(f, (0, 500), np.concatenate([q0[i], p0[i]]), method="DOP853", rtol=1e-13, atol=1e-13, t_eval=ts)
            assert np.abs(q[i] - s.y[:3].T).max() < bound


def test_sie_first_order_and_time_accumulation():
    pot = op.milky_way_potential()
    q0, p0 = synthetic_ics(pot, 4, seed=1)
    ref, *_ = cref.integrate_dopri8(pot, q0, p0, 0.0, 50.0, [50.0], rtol=1e-12, atol=1e-12)
    errs = []
    for dt in (0.1, 0.05, 0.025):
        q, p, st, n = cref.integrate_fixed(pot, q0, p0, 0.0, 50.0, dt, [50.0])
        errs.append(np.abs(q - ref).max())
    assert 1.8 < errs[0] / errs[1] < 2.2 and 1.8 < errs[1] / errs[2] < 2.2
    # C1: dt0 = 0.1 over 1 Gyr takes exactly 10 000 steps (accumulated time, last step clipped to t1)
    q, p, st, n = cref.integrate_fixed(pot, q0[:1], p0[:1], 0.0, 1000.0, 0.1, [1000.0])
    assert n[0] == 10000 and st[0] == 0


def test_sie_matches_plain_numpy_restatement():
    pot = op.milky_way_potential()
    q0, p0 = synthetic_ics(pot, 16, seed=3)
    q, p = q0.copy(), p0.copy()
    t, t1, dt = 0.0, 7.0, 0.1
    tn = t + dt
    while t < t1:
        h = tn - t
        q = q + p * h
        p = p + op.acceleration(pot, q) * h
        t = tn
        tn = t + dt
        if tn > t1 - 1e-10:
            tn = t1
    qc, pc, st, n = cref.integrate_fixed(pot, q0, p0, 0.0, 7.0, 0.1, [7.0])
    assert np.allclose(qc[:, 0], q, rtol=1e-14, atol=0) and np.allclose(pc[:, 0], p, rtol=1e-13, atol=0)


def test_sie_saves_interpolate_linearly_and_backward_integration():
    pot = op.milky_way_potential()
    q0, p0 = synthetic_ics(pot, 4, seed=4)
    ts = np.array([0.0, 0.25, 1.0, 3.333, 10.0])
    q, p, st, n = cref.integrate_fixed(pot, q0, p0, 0.0, 10.0, 0.5, ts)
    assert np.array_equal(q[:, 0], q0)
    qa, *_ = cref.integrate_fixed(pot, q0, p0, 0.0, 10.0, 0.5, [0.5])
    assert np.allclose(q[:, 1], q0 + 0.5 * (qa[:, 0] - q0), rtol=1e-15)
    # backward: integrate to -10 and save there
    qb, pb, st, n = cref.integrate_fixed(pot, q0, p0, 0.0, -10.0, -0.5, [-10.0])
    assert (st == 0).all() and n[0] == 20 and np.isfinite(qb).all()


def test_max_steps_status():
    pot = op.milky_way_potential()
    q0, p0 = synthetic_ics(pot, 2, seed=5)
    q, p, st, n = cref.integrate_fixed(pot, q0, p0, 0.0, 10.0, 0.1, [10.0], max_steps=5)
    assert (st == cref.MAX_STEPS).all() and np.isnan(q).all()
    q, p, st, na, nt = cref.integrate_dopri8(pot, q0, p0, 0.0, 5000.0, [5000.0], rtol=1e-10, atol=1e-10, max_steps=7)
    assert (st == cref.MAX_STEPS).all() and (nt == 7).all()


def test_energy_conservation_dopri8():
    pot = op.bovy_mw_potential_2014()
    q0, p0 = synthetic_ics(pot, 16, seed=6)
    q, p, st, na, nt = cref.integrate_dopri8(pot, q0, p0, 0.0, 2000.0, [2000.0], rtol=1e-10, atol=1e-10)
    E0 = 0.5 * (p0**2).sum(-1) + op.potential(pot, q0)
    E1 = 0.5 * (p[:, 0] ** 2).sum(-1) + op.potential(pot, q[:, 0])
    assert np.abs(E1 / E0 - 1).max() < 1e-8


def test_fardal_release_algebra():
    """df/fardal15.py:49-94 restated in numpy, against the C oracle."""
    pot = op.milky_way_potential()
    rng = np.random.default_rng(0)
    M = 64
    x = rng.normal(size=(M, 3)) * 10
    v = rng.normal(size=(M, 3)) * 0.1
    n = rng.standard_normal((4, M))
    mass = np.full(M, 1e4)
    ql, pl, qt, pt = cref.release_fardal(pot, x, v, mass, n)
    r = np.linalg.norm(x, axis=1, keepdims=True)
    L = np.cross(x, v)
    om = np.linalg.norm(L / r**2, axis=1, keepdims=True)
    rt = np.cbrt(pot.G * mass[:, None] / (om**2 - op.d2potential_dr2(pot, x)[:, None]))
    rh, zh = x / r, L / np.linalg.norm(L, axis=1, keepdims=True)
    ph = v - (v * rh).sum(1, keepdims=True) * rh
    ph /= np.linalg.norm(ph, axis=1, keepdims=True)
    kr = 2.0 + n[0][:, None] * 0.5
    kvphi = kr * (0.3 + n[1][:, None] * 0.5)
    kz, kvz = n[2][:, None] * 0.5, n[3][:, None] * 0.5
    assert np.allclose(qt, x + rt * (kr * rh + kz * zh), rtol=1e-13)
    assert np.allclose(pt, v + om * rt * (kvphi * ph + kvz * zh), rtol=1e-12, atol=1e-16)
    assert np.allclose(ql, x - rt * (kr * rh - kz * zh), rtol=1e-13)
    assert np.allclose(pl, v - om * rt * (kvphi * ph - kvz * zh), rtol=1e-12, atol=1e-16)


def test_mockstream_shapes_and_finiteness_like_reference():
    """tests/unit/dynamics/mockstream/test_mockstreamgenerator.py:18-117: NFW host, 10 stripping times."""
    pot = op.single(op.KIND_NFW, 1.0e12, 15.0)
    ts = np.linspace(0.0, 4000.0, 10)
    draws = np.random.default_rng(12).standard_normal((4, 10))
    out = cref.mockstream(pot, [[30.0, 10, 20]], [[10 * KMS, -150 * KMS, -20 * KMS]], ts, 1e4, draws)
    for k in ("lead_q", "lead_p", "trail_q", "trail_p"):
        assert out[k].shape == (10, 3) and np.isfinite(out[k]).all()
    assert out["prog_q"].shape == (10, 3)


def test_dopri5_tableau_and_oracle():
    """Dormand-Prince 5(4): order conditions, Shampine mid-point, and the oracle against a DOP853 truth."""
    from oracle import dopri5_tableau as t5

    r = t5.verify()
    assert r["order5_sol"] < 5e-15 and r["order4_hat"] < 5e-15
    pot = op.milky_way_potential()
    q0, p0 = synthetic_ics(pot, 4, seed=9)
    ts = np.linspace(0, 300, 7)
    for tol, bound in ((1e-7, 2e-3), (1e-10, 2e-6)):
        q, p, st, na, nt = cref.integrate_dopri8(pot, q0, p0, 0.0, 300.0, ts, rtol=tol, atol=tol, solver="dopri5")
        assert (st == 0).all()
        for i in range(4):
            def f(t, y):
                return np.concatenate([y[3:], -op.gradient(pot, y[:3])])
            s = solve_ivp(f, (0, 300), np.concatenate([q0[i], p0[i]]), method="DOP853", rtol=1e-13, atol=1e-13, t_eval=ts)
            assert np.abs(q[i] - s.y[:3].T).max() < bound
    # dtmin with force_dtmin (the experimental API's controller): never steps below dtmin, always finishes
    q, p, st, na, nt = cref.integrate_dopri8(pot, q0, p0, 0.0, 300.0, [300.0], rtol=1e-12, atol=1e-12, dtmin=0.3,
                                             solver="dopri5")
    assert (st == 0).all() and (nt <= 1001).all()


def test_reference_experimental_integrate_orbit_doctest():
    """experimental/integrate.py:159-247: 8-decimal values incl. a save inside the first step (dense output)."""
    for case in KATS["experimental"]:
        pot = op.single(op.KIND_NFW, *case["model"]["params"])
        ts = np.linspace(case["t0"], case["t1"], case["n_saves"])
        q, p, st, na, nt = cref.integrate_dopri8(pot, case["q0"], case["p0"], case["t0"], case["t1"], ts, rtol=case["rtol"],
                                                 atol=case["atol"] if False else case["rtol"], dtmin=case["dtmin"],
                                                 max_steps=case["max_steps"])
        for row, ref in case["rows"].items():
            assert np.allclose(q[:, int(row)], ref["q"], atol=case["atol"], rtol=0)
            assert np.allclose(p[:, int(row)], ref["p"], atol=case["atol"], rtol=0)


def test_reference_integrate_field_doctest_given_its_first_step():
    """dynamics/_src/solver.py:341-365 (Kepler, dtmin = 0.05): see the note in orbit_kats.json."""
    for case in KATS["integrate_field"]:
        pot = op.single(op.KIND_HERNQUIST, case["model"]["params"][0], 0.0)
        ts = np.linspace(case["t0"], case["t1"], case["n_saves"])
        q, p, st, na, nt = cref.integrate_dopri8(pot, [case["q0"]], [case["p0"]], case["t0"], case["t1"], ts,
                                                 rtol=case["rtol"], atol=case["atol_solver"], dtmin=case["dtmin"],
                                                 max_steps=case["max_steps"], dt0=case["dt0_observed"])
        for row, ref in case["rows"].items():
            assert np.allclose(q[0, int(row)], ref["q"], atol=ref["atol"], rtol=0)
            assert np.allclose(p[0, int(row)], ref["p"], atol=ref["atol"], rtol=0)


def test_reference_linear_parameter_doctest_and_time_dependent_oracles_agree():
    """potential/_src/params/core.py:52-74: a Kepler potential losing mass linearly; cylindrical radius at 10 saves
    printed to 3 decimals.  Also: numpy and C oracles agree on time-dependent composites, frozen-time evaluation and
    a fixed-step run with a growing halo."""
    case = KATS["time_dependent"][0]
    pot = op.Potential((op.Component(op.KIND_HERNQUIST, (case["point_value"] - case["slope_msun_per_myr"] * case["point_time"], 0.0),
                                     rates=(case["slope_msun_per_myr"], 0.0)),))
    ts = np.linspace(case["t0"], case["t1"], case["n_saves"])
    q, p, st, na, nt = cref.integrate_dopri8(pot, [case["q0"]], [np.array(case["p0_kms"]) * KATS["kms"]], case["t0"], case["t1"], ts,
                                             rtol=case["rtol"], atol=case["atol_solver"])
    assert st[0] == 0 and np.allclose(np.hypot(q[0, :, 0], q[0, :, 1]), case["rho"], atol=case["atol"], rtol=0)
    # frozen-time evaluation: numpy == C, and equals the static potential with the parameters of that time
    mix = op.Potential((op.Component(op.KIND_MN, (6.8e10, 3.0, 0.28), rates=(1e7, 1e-4, 0.0)),
                        op.Component(op.KIND_NFW, (5.4e11, 15.62), rates=(2e8, 1e-3)),
                        op.Component(op.KIND_HERNQUIST, (5e9, 1.0))))
    x = np.random.default_rng(0).normal(size=(50, 3)) * 8
    t = 730.0
    static = op.Potential((op.Component(op.KIND_MN, (6.8e10 + 1e7 * t, 3.0 + 1e-4 * t, 0.28)),
                           op.Component(op.KIND_NFW, (5.4e11 + 2e8 * t, 15.62 + 1e-3 * t)), op.Component(op.KIND_HERNQUIST, (5e9, 1.0))))
    c = cref.potential_eval(mix, x, ("phi", "grad", "hess"), t=t)
    assert np.allclose(c["grad"], op.gradient(mix, x, t), rtol=1e-14) and np.allclose(c["grad"], op.gradient(static, x), rtol=1e-14)
    assert np.allclose(c["hess"], op.hessian(static, x), rtol=1e-12, atol=1e-18) and np.allclose(c["phi"], op.potential(static, x), rtol=1e-14)
    # a growing halo changes the orbit; zero rates reproduce the static run bit for bit
    q0, p0 = np.array([[8.0, 0.0, 1.0]]), np.array([[0.0, 0.2, 0.02]])
    qa, pa, *_ = cref.integrate_fixed(mix, q0, p0, 0.0, 500.0, 0.1, [500.0])
    frozen0 = op.Potential(tuple(op.Component(cc.kind, cc.params) for cc in mix.components))
    qb, pb, *_ = cref.integrate_fixed(frozen0, q0, p0, 0.0, 500.0, 0.1, [500.0])
    zero = op.Potential(tuple(op.Component(cc.kind, cc.params, rates=(0.0,) * len(cc.params)) for cc in mix.components))
    qc, pc, *_ = cref.integrate_fixed(zero, q0, p0, 0.0, 500.0, 0.1, [500.0])
    assert np.array_equal(qb, qc) and np.abs(qa - qb).max() > 1e-3


TIDAL_KATS = [  # (potential, x [kpc], v [kpc/Myr], mass, r_t [kpc]) -- dynamics/_src/cluster/api.py:54-64,70-76,180-198
    ("MilkyWayPotential", [8.0, 0.0, 0.0], [0.0, 220.0 * 0.001022712165045695, 0.0], 1e4, 0.02929074),
    ("MilkyWayPotential", [8.0, 0.0, 0.0], [0.0, 220.0, 0.0], 1e4, 0.00039036),
    ("NFW(1e12, 20)", [8.0, 0.0, 0.0], [8.0, 1e-7, 0.0], 1e4, 0.06362008),  # radial orbit; a 1e-7 tangential part fixes the plane
]


def test_reference_tidal_radius_and_lagrange_point_doctests():
    """King (1962) tidal radius r_t = cbrt(G m / (Omega^2 - d2Phi/dr2)) (cluster/radius.py:198-215), read off the release
    kernel with all Fardal draws zero (k_r = 2: x_lead = x - 2 r_t r_hat): the reference's lagrange_points / tidal_radius
    doctest values to all printed digits."""
    for name, x, v, mass, rt in TIDAL_KATS:
        pot = op.milky_way_potential() if name == "MilkyWayPotential" else op.single(op.KIND_NFW, 1e12, 20.0)
        ql, pl, qt, pt = cref.release_fardal(pot, np.array([x]), np.array([v]), mass, np.zeros((4, 1)))
        assert abs((x[0] - ql[0, 0]) / 2 - rt) < 6e-9 and abs((qt[0, 0] - x[0]) / 2 - rt) < 6e-9, name
