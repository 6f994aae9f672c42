"""GPU: the ``galax.dynamics.experimental`` mirror (integrate_orbit, Fardal2015DF, StreamSimulator) and Dopri5."""
import numpy as np
import pytest

import galax_b200.dynamics as gd
import galax_b200.experimental as ge
import galax_b200.potential as gp
from oracle import cref
from oracle import potentials as op

from conftest import synthetic_ics

pytestmark = pytest.mark.gpu
KMS = gp.KMS


def test_dopri5_matches_oracle_and_truth():
    pot, opot = gp.MilkyWayPotential(), op.milky_way_potential()
    q0, p0 = synthetic_ics(opot, 128, seed=31)
    ts = np.linspace(0.0, 300.0, 13)
    kw = dict(solver=gd.Dopri5(), controller=gd.PIDController(1e-8, 1e-8), dt0=None, max_steps=None)
    q, p, st, stats = gd._integrate(pot, q0, p0, 0.0, 300.0, ts, **kw)
    qr, pr, sr, na, nt = cref.integrate_dopri8(opot, q0, p0, 0.0, 300.0, ts, rtol=1e-8, atol=1e-8, solver="dopri5")
    qt, *_ = cref.integrate_dopri8(opot, q0, p0, 0.0, 300.0, ts, rtol=1e-13, atol=1e-13)
    d = (np.abs(q - qr) / (1e-8 + 1e-8 * np.abs(qr))).max(axis=(1, 2))
    assert np.median(d) <= 10.0 and np.mean(d <= 10.0) >= 0.6
    eg, eo = np.abs(q - qt).max(axis=(1, 2)), np.abs(qr - qt).max(axis=(1, 2))
    assert 0.5 <= np.median(eg) / np.median(eo) <= 2.0
    assert abs(int(stats["num_steps"].sum()) / int(nt.sum()) - 1) < 0.02
    # Dopri5 needs several times more steps than Dopri8 at the same tolerance
    _, _, _, s8 = gd._integrate(pot, q0, p0, 0.0, 300.0, ts, solver=gd.Dopri8(), controller=gd.PIDController(1e-8, 1e-8),
                                dt0=None, max_steps=None)
    assert int(stats["num_steps"].sum()) > 1.5 * int(s8["num_steps"].sum())


def test_dopri5_short_horizon_step_end_parity_and_single_orbit_record_path():
    pot, opot = gp.MilkyWayPotential2022(), op.milky_way_potential_2022()
    q0, p0 = synthetic_ics(opot, 64, seed=32)
    kw = dict(solver=gd.Dopri5(), controller=gd.PIDController(1e-9, 1e-9), dt0=0.5, max_steps=None)
    q, p, st, stats = gd._integrate(pot, q0, p0, 0.0, 20.0, [20.0], **kw)
    qr, pr, sr, na, nt = cref.integrate_dopri8(opot, q0, p0, 0.0, 20.0, [20.0], rtol=1e-9, atol=1e-9, dt0=0.5, solver="dopri5")
    assert np.mean(stats["num_steps"].cpu().numpy() == nt) > 0.8
    assert np.median(np.abs(q - qr).max(axis=(1, 2))) < 1e-10
    ts = np.linspace(0.0, 500.0, 200)
    a = gd._integrate(pot, q0[:1], p0[:1], 0.0, 500.0, ts, solver=gd.Dopri5(), controller=gd.PIDController(1e-8, 1e-8),
                      dt0=None, max_steps=None)
    b = gd._integrate(pot, q0[:2], p0[:2], 0.0, 500.0, ts, solver=gd.Dopri5(), controller=gd.PIDController(1e-8, 1e-8),
                      dt0=None, max_steps=None)
    assert np.allclose(a[0][0], b[0][0], rtol=1e-13, atol=1e-13)


def test_integrate_orbit_reference_benchmark_shapes():
    """tests/benchmark/test_experimental.py:80-96: Hernquist, one orbit, 1000 saves over 100 Myr; scalar / batched."""
    pot = gp.HernquistPotential(1e12, 10.0)
    opot = op.single(op.KIND_HERNQUIST, 1e12, 10.0)
    qp0 = (np.array([15.0, 0.0, 0.0]), np.array([0.0, 0.225, 0.0]))
    saveat = np.linspace(0.0, 100.0, 1000)
    sol = ge.integrate_orbit(pot, qp0, saveat=saveat)
    assert sol.ys[0].shape == (1000, 3) and sol.t0 == 0.0 and sol.t1 == 100.0
    qr, pr, *_ = cref.integrate_dopri8(opot, [qp0[0]], [qp0[1]], 0.0, 100.0, saveat, rtol=1e-7, atol=1e-7, dtmin=0.3,
                                       max_steps=10_000)
    assert np.abs(sol.ys[0] - qr[0]).max() < 1e-5
    end = ge.integrate_orbit(ge.VMap, pot, qp0, 0.0, 100.0)
    assert end.ys[0].shape == (3,) and np.abs(end.ys[0] - qr[0, -1]).max() < 1e-5
    one = ge.integrate_orbit(pot, qp0, 0.0, saveat=50.0)
    assert one.ys[0].shape == (3,)
    B = 64
    q0 = np.repeat(qp0[0][None], B, 0) * np.linspace(0.8, 1.2, B)[:, None]
    p0 = np.repeat(qp0[1][None], B, 0)
    for strat in (ge.Scan, ge.VMap, ge.NoLoop):
        sb = ge.integrate_orbit(strat, pot, (q0, p0), 0.0, 100.0, saveat=saveat)
        assert sb.ys[0].shape == (B, 1000, 3)
    t0s = np.linspace(-50.0, 0.0, B)
    sb = ge.integrate_orbit(ge.VMap, pot, (q0, p0), t0s, 100.0)
    assert sb.ys[0].shape == (B, 3) and np.isfinite(sb.ys[0]).all()
    with pytest.raises(ValueError):
        ge.integrate_orbit(pot, qp0, None, 1.0)


def test_fardal2015df_parameters_and_stream_simulator():
    """experimental/stream.py doctest setup: Hernquist(1e12, 10), 2000 release times over [-4000, -150] Myr."""
    pot = gp.HernquistPotential(1e12, 10.0)
    opot = op.single(op.KIND_HERNQUIST, 1e12, 10.0)
    rng = np.random.default_rng(0)
    x = np.array([15.0, 0.0, 0.0])
    v = np.array([0.0, 220.0 * KMS, 0.0])
    n = rng.standard_normal((4, 1))
    xl, vl, xt, vt = ge.Fardal2015DF().sample(n, pot, 0.0, x, v, 1e5)
    ref = cref.release_fardal(opot, x[None], v[None], 1e5, n)
    assert xl.shape == (3,) and np.allclose(xl, ref[0][0], rtol=1e-13) and np.allclose(vt, ref[3][0], rtol=1e-12)
    df = ge.Fardal2015DF(kr_bar=1.5, sigma_kr=0.25, kvphi_bar=0.4, sigma_kz=0.1)
    xl2, _, xt2, _ = df.sample(n, pot, 0.0, x, v, 1e5)
    # radial offset scales with k_r = 1.5 + 0.25 n instead of 2 + 0.5 n
    rt = np.linalg.norm(ref[2][0] - x) / np.hypot(2 + 0.5 * n[0, 0], 0.5 * n[2, 0])
    assert np.isclose(np.linalg.norm(xt2 - x), rt * np.hypot(1.5 + 0.25 * n[0, 0], 0.1 * n[2, 0]), rtol=1e-10)

    # the reference's doctest of Fardal2015DF.sample(jr.key(0), NFW(1e12, 15), ...) (experimental/df.py:110-123)
    nfw = gp.NFWPotential(1e12, 15.0)
    xl3, vl3, xt3, vt3 = ge.Fardal2015DF().sample(0, nfw, 0.0, np.array([15.0, 0, 0]), np.array([0, 220.0, 0]), 1e5)
    assert np.allclose(xl3, [1.49962403e01, 0.0, 2.49694925e-04], rtol=2e-9, atol=1e-12)
    assert np.allclose(vt3, [0.0, 2.19977919e02, 6.34205795e-03], rtol=2e-9, atol=1e-12)

    M = 2000
    release = np.linspace(-4000.0, -150.0, M)
    sim = ge.StreamSimulator()
    ics = sim.init(pot, (x, np.array([0.0, 0.225, 0.0])), 0.0, release_times=release, Msat=1e5, key=0)
    assert ics.qp_lead[0].shape == (M, 3) and np.isfinite(ics.qp_trail[1]).all() and ics.prog_mass.shape == (M,)
    # the reference's doctest of StreamSimulator.init(..., key=jr.key(0)) and .run (experimental/stream.py:79-118): all
    # printed digits (the GPU's step sequence can differ from the reference's in the last bits: 1e-6 margin)
    assert np.allclose(ics.qp_lead[0][0], [-10.76187104, -7.35400639, 0.0674116], rtol=0, atol=1e-6)
    assert np.allclose(ics.qp_lead[0][-1], [-4.72896837, 14.03657666, -0.09171104], rtol=0, atol=1e-6)
    assert np.allclose(ics.qp_trail[0][0], [-11.00416221, -7.5195734, 0.0674116], rtol=0, atol=1e-6)
    assert np.allclose(ics.qp_lead[1][0], [4.77386246e-02, -2.74264308e-01, -4.68601912e-04], rtol=0, atol=1e-7)
    assert np.allclose(ics.qp_trail[1][-1], [-2.10223491e-01, -8.18272245e-02, -1.58559419e-04], rtol=0, atol=1e-7)
    lead, trail = sim.run(pot, ics, t1=0.0)
    assert lead[0].shape == (M, 3) and trail[1].shape == (M, 3) and np.isfinite(lead[0]).all()
    assert np.allclose(lead[0][0], [-4.99685677e00, 5.65910858e00, 3.63136282e-02], rtol=0, atol=2e-5)
    assert np.allclose(lead[0][-1], [1.48125263e01, 3.73149460e-01, 4.11255117e-02], rtol=0, atol=2e-6)
    assert np.allclose(lead[1][-1], [-1.39058722e-02, 2.24719748e-01, -1.28802309e-03], rtol=0, atol=2e-7)
    # oracle: same ICs through the Dopri5 oracle with dtmin = 0.3
    qr, pr, st, na, nt = cref.integrate_dopri8(opot, ics.qp_lead[0][:200], ics.qp_lead[1][:200], release[:200], 0.0, [0.0],
                                               rtol=1e-7, atol=1e-7, dtmin=0.3, max_steps=10_000, solver="dopri5")
    assert np.median(np.abs(lead[0][:200] - qr[:, 0]).max(axis=1)) < 1e-3


def test_stream_simulator_key_chain_draws_on_device():
    """gx_jax_fardal_chain (host key chain + device normals) against the numpy restatement, and a 2e5-release init."""
    import time

    from galax_b200 import jaxrandom as jr

    k = jr.key(3)
    d = ge._fardal_chain_normals(k, 3000).cpu().numpy()
    ref = jr.fardal_draws_per_key(jr.split_chain(k, 3000))
    assert np.abs(d - ref).max() < 1e-14 * np.abs(ref).max()
    pot = gp.HernquistPotential(1e12, 10.0)
    M = 200_000
    t0 = time.perf_counter()
    ics = ge.StreamSimulator().init(pot, (np.array([15.0, 0.0, 0.0]), np.array([0.0, 0.225, 0.0])), 0.0,
                                    release_times=np.linspace(-4000.0, -150.0, M), Msat=1e5, key=0)
    assert time.perf_counter() - t0 < 5.0 and np.isfinite(ics.qp_lead[0]).all() and ics.qp_lead[0].shape == (M, 3)


def test_reference_experimental_integrate_orbit_doctest_on_gpu():
    """experimental/integrate.py:159-247 through the GPU path: 8 decimals, incl. a dense-output value."""
    import json
    from pathlib import Path

    kats = json.loads((Path(__file__).parent / "golden" / "orbit_kats.json").read_text())
    for case in kats["experimental"]:
        pot = gp.NFWPotential(*case["model"]["params"])
        ts = np.linspace(case["t0"], case["t1"], case["n_saves"])
        sol = ge.integrate_orbit(pot, (np.array(case["q0"]), np.array(case["p0"])), t0=case["t0"], t1=case["t1"], saveat=ts)
        for row, ref in case["rows"].items():
            assert np.allclose(sol.ys[0][:, int(row)], ref["q"], atol=case["atol"], rtol=0)
            assert np.allclose(sol.ys[1][:, int(row)], ref["p"], atol=case["atol"], rtol=0)
