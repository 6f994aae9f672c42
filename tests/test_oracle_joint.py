"""The restated Dopri8 + PID loop on the reference's JOINT (shared-step) batch path, pinned on its doctests.

`oracle/joint_dopri8.py` is the generic first-order form in numpy; the C oracle and the kernels use the Nystrom
form with per-particle control.  Three things are pinned here: the tableau (one 10 Myr step, 8 digits), the
controller + initial step on two 12-dimensional solves (8 digits), and C oracle == numpy form for one particle.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import cref, joint_dopri8 as jd, potentials as op

KATS = json.loads((Path(__file__).parent / "golden" / "orbit_kats.json").read_text())
KMS = KATS["kms"]


@pytest.mark.parametrize("case", KATS["joint"], ids=lambda c: c["name"][:32])
def test_reference_joint_doctests(case):
    pot = op.single(op.KIND_HERNQUIST, *case["model"]["params"])
    q0, p0 = np.array(case["q0"], float), np.array(case["p0_kms"], float) * KMS
    if case["kind"] == "step":
        n = len(q0)
        f = lambda t, y: np.concatenate([y[n:], -op.gradient(pot, y[:n])], axis=0)
        y1, _, _ = jd.step(f, case["t0"], np.concatenate([q0, p0]), case["t1"] - case["t0"])
        q, p = y1[:n], y1[n:]
    else:
        qs, ps, stats = jd.solve(pot, q0, p0, case["t0"], case["t1"], rtol=case["rtol"], atol=case["atol_solver"])
        q, p = qs[0], ps[0]
        assert stats["num_steps"] < 4096
    assert np.allclose(q, case["q"], atol=case["atol"], rtol=0)
    assert np.allclose(p, case["p"], atol=case["atol"], rtol=0)


def test_c_oracle_nystrom_form_equals_numpy_generic_form():
    """Same algorithm, two independent codings: identical step counts, values within a fraction of the tolerance
    (step sizes come from a rounding-sensitive error estimate)."""
    pot = op.milky_way_potential()
    q0, p0 = np.array([[8.0, 0.5, 1.0]]), np.array([[0.02, 0.21, 0.05]])
    ts = np.linspace(0.0, 300.0, 7)
    qn, pn, stats = jd.solve(pot, q0, p0, 0.0, 300.0, ts, rtol=1e-9, atol=1e-9, dt0=0.5)
    qc, pc, st, na, nt = cref.integrate_dopri8(pot, q0, p0, 0.0, 300.0, ts, rtol=1e-9, atol=1e-9, dt0=0.5)
    assert (int(na[0]), int(nt[0])) == (stats["num_accepted_steps"], stats["num_steps"])
    assert np.abs(qc[0] - qn[:, 0]).max() < 2e-9 and np.abs(pc[0] - pn[:, 0]).max() < 2e-10
    # and with the initial-step heuristic (no dt0 given)
    qn, pn, stats = jd.solve(pot, q0, p0, 0.0, 300.0, ts, rtol=1e-9, atol=1e-9)
    qc, pc, st, na, nt = cref.integrate_dopri8(pot, q0, p0, 0.0, 300.0, ts, rtol=1e-9, atol=1e-9)
    assert abs(int(nt[0]) - stats["num_steps"]) <= 2
    assert np.abs(qc[0] - qn[:, 0]).max() < 5e-9
