"""Dormand-Prince 5(4) tableau (7 stages, FSAL) as used by ``diffrax.Dopri5`` -- TEST INFRASTRUCTURE.

The reference's experimental stream simulator defaults to it
(``/root/reference/src/galax/dynamics/_src/experimental/stream.py:32-41``).  Coefficients: Dormand & Prince (1980),
J. Comp. Appl. Math. 6, 19; ``verify()`` checks the rooted-tree order conditions (order 5 for b, order 4 for the
embedded b_hat).  Dense output: diffrax's ``_Dopri5Interpolation`` is a ``FourthOrderPolynomialInterpolation``: the
quartic through y0, y1, f0 h, f1 h and y_mid = y0 + sum_i C_MID[i] k_i, with Shampine's mid-point weights (the
same numbers as torchdiffeq's ``DPS_C_MID``).  ``dense_b()`` rewrites that quartic as stage weights
b_i(theta) = sum_m DENSE_B[i][m-1] theta^m so that it shares the evaluation code of Dopri8.
"""
from __future__ import annotations

from fractions import Fraction as F

import numpy as np

from . import dopri8_tableau as t8

N_STAGES = 7
ORDER = 5
C = [F(0), F(1, 5), F(3, 10), F(4, 5), F(8, 9), F(1), F(1)]
A = [
    [],
    [F(1, 5)],
    [F(3, 40), F(9, 40)],
    [F(44, 45), F(-56, 15), F(32, 9)],
    [F(19372, 6561), F(-25360, 2187), F(64448, 6561), F(-212, 729)],
    [F(9017, 3168), F(-355, 33), F(46732, 5247), F(49, 176), F(-5103, 18656)],
    [F(35, 384), F(0), F(500, 1113), F(125, 192), F(-2187, 6784), F(11, 84)],
]
B_SOL = [F(35, 384), F(0), F(500, 1113), F(125, 192), F(-2187, 6784), F(11, 84), F(0)]
# the embedded weights diffrax (like torchdiffeq) uses: b_sol - b_hat is 2/3 of the textbook Dormand-Prince error
# coefficients (71/57600, ...).  Settled by the reference's StreamSimulator doctest, which this reproduces to 2e-9 (the
# textbook estimate takes 11 % shorter steps and lands 4e-3 kpc away after 8 Gyr).
B_HAT = [F(1951, 21600), F(0), F(22642, 50085), F(451, 720), F(-12231, 42400), F(649, 6300), F(1, 60)]
C_MID = [F(6025192743, 30085553152) / 2, F(0), F(51252292925, 65400821598) / 2, F(-2691868925, 45128329728) / 2,
         F(187940372067, 1594534317056) / 2, F(-1776094331, 19743644256) / 2, F(11237099, 235043384) / 2]


def a_matrix() -> np.ndarray:
    M = np.zeros((N_STAGES, N_STAGES))
    for i, row in enumerate(A):
        for j, v in enumerate(row):
            M[i, j] = float(v)
    return M


def b_sol() -> np.ndarray:
    return np.array([float(v) for v in B_SOL])


def b_err() -> np.ndarray:
    return np.array([float(s - h) for s, h in zip(B_SOL, B_HAT)])


def c_vec() -> np.ndarray:
    return np.array([float(v) for v in C])


def dense_b_exact() -> list[list[F]]:
    """Quartic through (y0, y1, f0 h, f1 h, y_mid) as stage weights: columns are theta^1 .. theta^6."""
    out = []
    for i in range(N_STAGES):
        d0, d6 = F(int(i == 0)), F(int(i == N_STAGES - 1))
        b, cm = B_SOL[i], C_MID[i]
        th1 = d0
        th2 = d6 - 4 * d0 - 5 * b + 16 * cm
        th3 = 5 * d0 - 3 * d6 + 14 * b - 32 * cm
        th4 = 2 * d6 - 2 * d0 - 8 * b + 16 * cm
        out.append([th1, th2, th3, th4, F(0), F(0)])
    return out


def dense_b() -> np.ndarray:
    return np.array([[float(v) for v in row] for row in dense_b_exact()])


def verify(tol=5e-15):
    Am = a_matrix()
    rows = np.abs(Am.sum(axis=1) - c_vec()).max()
    r5 = max(abs(r) for _, r in t8.order_residuals(b_sol(), 5, Am=Am))
    bh = np.array([float(v) for v in B_HAT])
    r4 = max(abs(r) for _, r in t8.order_residuals(bh, 4, Am=Am))
    r5h = max(abs(r) for _, r in t8.order_residuals(bh, 5, Am=Am))
    r6 = max(abs(r) for _, r in t8.order_residuals(b_sol(), 6, Am=Am))
    assert rows < tol and r5 < tol and r4 < tol and r5h > 1e-6 and r6 > 1e-6
    # the mid-point rule is (at least) 4th order at theta = 1/2, and the quartic hits y1 and f1 at theta = 1
    cm = np.array([float(v) for v in C_MID])
    rmid = max(abs(r) for _, r in t8.order_residuals(cm, 4, theta=0.5, Am=Am))
    Bd = dense_b()
    assert rmid < 1e-9 and np.abs(Bd.sum(axis=1) - b_sol()).max() < 1e-15
    d1 = Bd @ np.arange(1, 7)
    e_last = np.zeros(N_STAGES)
    e_last[-1] = 1
    assert np.abs(d1 - e_last).max() < 1e-14
    return {"row_sum": rows, "order5_sol": r5, "order4_hat": r4, "mid_order4": rmid}


if __name__ == "__main__":
    print(verify())
