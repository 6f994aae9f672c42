"""CPU: the generated Nystrom-form Dopri8 tables agree with the (independently typed) oracle tableau."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galax_b200" / "csrc"))


def test_generated_tables_match_oracle_tableau():
    import gen_tables as g

    from oracle import dopri8_tableau as tab

    t = g.tables()
    A = np.array([[float(v) for v in row] for row in t["A"]])
    assert np.array_equal(A, tab.a_matrix())
    assert np.array_equal(np.array([float(v) for v in t["B"]]), tab.b_sol())
    assert np.array_equal(np.array([float(v) for v in t["E"]]), tab.b_err())
    assert np.allclose(np.array([[float(v) for v in r] for r in t["AA"]]), A @ A, rtol=1e-13, atol=1e-14)
    assert np.allclose(np.array([float(v) for v in t["EA"]]), tab.b_err() @ A, rtol=1e-12, atol=1e-15)
    DB = np.array([[float(v) for v in r] for r in t["DB"]])
    assert np.array_equal(DB, tab.dense_b())
    DQ = np.array([[float(v) for v in r] for r in t["DQ"]])
    assert np.allclose(DQ, (DB.T @ A).T, rtol=1e-12, atol=1e-12)
    # row sums of A are the nodes: q_i = q0 + c_i h p0 + ... is exact
    assert np.abs(A.sum(1) - np.array([float(c) for c in t["C"]])).max() < 2e-15


def test_header_is_current():
    import gen_tables as g

    hdr = (ROOT / "galax_b200" / "csrc" / "gx_tables.h").read_text()
    t = g.tables()
    for v in (t["AA"][13][0], t["EA"][5], t["DQ"][0][2]):
        assert repr(float(v)) in hdr


def test_dopri5_tables_match_oracle_tableau():
    import gen_tables as g

    from oracle import dopri5_tableau as t5

    t = g.tables("dp5")
    A = np.array([[float(v) for v in row] for row in t["A"]])
    assert np.array_equal(A, t5.a_matrix())
    assert np.array_equal(np.array([float(v) for v in t["B"]]), t5.b_sol())
    assert np.array_equal(np.array([float(v) for v in t["E"]]), t5.b_err())
    assert np.array_equal(np.array([[float(v) for v in r] for r in t["DB"]]), t5.dense_b())
    assert np.allclose(np.array([[float(v) for v in r] for r in t["AA"]]), A @ A, rtol=1e-13, atol=1e-15)
