"""CPU: the jax.random restatement (threefry2x32, split, normal) against published vectors and the reference's
own doctests (the only places where the reference prints values that depend on its PRNG stream)."""
import numpy as np

from galax_b200 import jaxrandom as jr
from oracle import cref
from oracle import potentials as op


def test_threefry_known_answers():
    """Random123 known-answer vectors (Salmon et al. 2011), also used by jax's own test-suite."""
    kat = [((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]  # fmt: skip
    for (k0, k1), (c0, c1), (e0, e1) in kat:
        r0, r1 = jr.threefry2x32(k0, k1, [c0], [c1])
        assert (int(r0[0]), int(r1[0])) == (e0, e1)
        assert jr._threefry_int(k0, k1, c0, c1) == (e0, e1)


def test_reference_doctest_fardal2015df_sample_key0():
    """experimental/df.py:110-123: Fardal2015DF().sample(jr.key(0), NFW(1e12, 15), t=0, x=[15,0,0], v=[0,220,0],
    Msat=1e5) -- pins key(), split(), normal() AND the release algebra / tidal radius on a reference-made value."""
    pot = op.single(op.KIND_NFW, 1e12, 15.0)
    n = jr.fardal_draws(0, 1)
    ql, pl, qt, pt = cref.release_fardal(pot, [[15.0, 0, 0]], [[0, 220.0, 0]], 1e5, n)
    assert np.allclose(ql[0], [1.49962403e01, 0.0, 2.49694925e-04], rtol=2e-9, atol=1e-12)
    assert np.allclose(pl[0], [0.0, 2.20022081e02, 6.34205795e-03], rtol=2e-9, atol=1e-12)
    assert np.allclose(qt[0], [1.50037597e01, 0.0, 2.49694925e-04], rtol=2e-9, atol=1e-12)
    assert np.allclose(pt[0], [0.0, 2.19977919e02, 6.34205795e-03], rtol=2e-9, atol=1e-12)


def test_reference_doctest_stream_simulator_init_and_run_key0():
    """experimental/stream.py:79-118: StreamSimulator().init(Hernquist(1e12, 10), qp0, 0, release_times=
    linspace(-4000, -150, 2000), Msat=1e5, key=jr.key(0)) and .run(pot, ics, t1=0): every printed number (positions,
    velocities, both arms, first and last particle) to the last digit.  One chain pins Dopri5 (diffrax's embedded
    weights: error = 2/3 of the textbook Dormand-Prince estimate), PIDController with forced dtmin = 0.3, the Dopri5
    dense output at the 2000 release times, the backward first leg, the jax PRNG key chain, Fardal2015DF and the tidal
    radius, and the per-particle start times of the final integration."""
    pot = op.single(op.KIND_HERNQUIST, 1e12, 10.0)
    rel = np.linspace(-4000.0, -150.0, 2000)
    kw = dict(rtol=1e-7, atol=1e-7, dtmin=0.3, max_steps=10_000, solver="dopri5")
    q, p, st, _, _ = cref.integrate_dopri8(pot, [[15.0, 0, 0]], [[0, 0.225, 0]], 0.0, rel[0], [rel[0]], **kw)
    Q, P, st2, _, _ = cref.integrate_dopri8(pot, q[0], p[0], rel[0], rel[-1], rel, **kw)
    assert st[0] == 0 and st2[0] == 0
    draws = jr.fardal_draws_per_key(jr.split_chain(jr.key(0), 2000))
    ql, pl, qt, pt = cref.release_fardal(pot, Q[0], P[0], 1e5, draws)
    A = dict(rtol=0, atol=6e-9)
    assert np.allclose(ql[0], [-10.76187104, -7.35400639, 0.0674116], **A)
    assert np.allclose(ql[-1], [-4.72896837, 14.03657666, -0.09171104], **A)
    assert np.allclose(qt[0], [-11.00416221, -7.5195734, 0.0674116], **A)
    assert np.allclose(qt[-1], [-4.83974586, 14.36538765, -0.09171104], **A)
    V = dict(rtol=0, atol=6e-10)
    assert np.allclose(pl[0], [4.77386246e-02, -2.74264308e-01, -4.68601912e-04], **V)
    assert np.allclose(pl[-1], [-2.09972781e-01, -8.17427593e-02, -1.58559419e-04], **V)
    assert np.allclose(pt[0], [5.07429914e-02, -2.78660906e-01, -4.68601912e-04], **V)
    assert np.allclose(pt[-1], [-2.10223491e-01, -8.18272245e-02, -1.58559419e-04], **V)
    # .run: every particle from its release time to t1 = 0 (the oldest one over 4 Gyr: 5e-8)
    qr, pr, st3, _, _ = cref.integrate_dopri8(pot, ql, pl, rel, 0.0, [0.0], **kw)
    assert (st3 == 0).all()
    assert np.allclose(qr[0, 0], [-4.99685677e00, 5.65910858e00, 3.63136282e-02], rtol=0, atol=1e-7)
    assert np.allclose(qr[-1, 0], [1.48125263e01, 3.73149460e-01, 4.11255117e-02], rtol=0, atol=1e-7)
    assert np.allclose(pr[0, 0], [-3.87842191e-01, -2.21692094e-01, 2.45336141e-03], rtol=0, atol=1e-8)
    assert np.allclose(pr[-1, 0], [-1.39058722e-02, 2.24719748e-01, -1.28802309e-03], rtol=0, atol=1e-8)


def test_split_chain_and_vectorised_draws_are_consistent():
    k = jr.key(12)
    sub = jr.split_chain(k, 5)
    a, b = k, None
    for i in range(5):
        both = jr.split(a, 2)
        a, b = both[0], both[1]
        assert np.array_equal(sub[i], b)
    one = jr.fardal_draws_per_key(sub[:1])
    assert np.allclose(one[:, 0], jr.fardal_draws(sub[0], 1)[:, 0])
    d = jr.fardal_draws(7, 100_000)
    assert abs(d.mean()) < 0.01 and abs(d.std() - 1) < 0.01 and np.isfinite(d).all()


def test_chen_multivariate_normal_svd_draws():
    """df/chen24.py:89-91: mean + (u sqrt(s)) z with threefry normals; the zero-variance row (Dv/v_esc == 1) must come
    out exactly, and the sample moments must reproduce the 6-D Gaussian (incl. the Dr-alpha covariance -4.9)."""
    from galax_b200 import jaxrandom as jr
    from galax_b200.dynamics import ChenStreamDF

    x = jr.chen_draws(7, 200_000)
    assert x.shape == (200_000, 6)
    assert np.allclose(x[:, 3], 1.0, atol=1e-12)
    assert np.allclose(x.mean(0), ChenStreamDF.mean, atol=5 * np.sqrt(np.diag(ChenStreamDF.cov) / 200_000) + 1e-12)
    C = np.cov(x.T)
    assert np.allclose(C, ChenStreamDF.cov, atol=0.02 * np.sqrt(np.outer(np.diag(ChenStreamDF.cov), np.diag(ChenStreamDF.cov))) + 1e-9)
    assert np.array_equal(x, jr.chen_draws(7, 200_000)) and not np.array_equal(x[:10], jr.chen_draws(8, 10))
    # the underlying normals are jax's: the first row is factor @ normal(key(7), (M, 6))[0]
    u, s, _ = np.linalg.svd(ChenStreamDF.cov)
    z = jr.normal(jr.key(7), (200_000, 6))
    assert np.allclose(x[0], ChenStreamDF.mean + (u * np.sqrt(s)) @ z[0], atol=1e-12)
