"""Prince-Dormand RK8(7)13M tableau + FSAL stage, as used by ``diffrax.Dopri8`` (0.7.0).

TEST INFRASTRUCTURE.  diffrax is a third-party dependency of the reference
(``/root/reference/uv.lock:644``) that is not vendored and not importable here; the
coefficients below are the published Prince & Dormand (1981) rational approximations
("RK8(7)13M", J. Comp. Appl. Math. 7, 67), which is what ``diffrax/_solver/dopri8.py``
tabulates.  ``verify()`` checks them against the Runge-Kutta order conditions (all 200 rooted
trees up to order 8 for ``B_SOL``; all 85 up to order 7 for the embedded ``B_HAT``), so a
mistyped digit cannot survive.

Dense output: see ``DENSE_B`` below and ``oracle/derive_dense_output.py``.
"""
from __future__ import annotations

from fractions import Fraction as F
from functools import lru_cache

import numpy as np

C = [F(0), F(1, 18), F(1, 12), F(1, 8), F(5, 16), F(3, 8), F(59, 400), F(93, 200),
     F(5490023248, 9719169821), F(13, 20), F(1201146811, 1299019798), F(1), F(1), F(1)]

A = [
    [],
    [F(1, 18)],
    [F(1, 48), F(1, 16)],
    [F(1, 32), 0, F(3, 32)],
    [F(5, 16), 0, F(-75, 64), F(75, 64)],
    [F(3, 80), 0, 0, F(3, 16), F(3, 20)],
    [F(29443841, 614563906), 0, 0, F(77736538, 692538347), F(-28693883, 1125000000), F(23124283, 1800000000)],
    [F(16016141, 946692911), 0, 0, F(61564180, 158732637), F(22789713, 633445777), F(545815736, 2771057229),
     F(-180193667, 1043307555)],
    [F(39632708, 573591083), 0, 0, F(-433636366, 683701615), F(-421739975, 2616292301), F(100302831, 723423059),
     F(790204164, 839813087), F(800635310, 3783071287)],
    [F(246121993, 1340847787), 0, 0, F(-37695042795, 15268766246), F(-309121744, 1061227803),
     F(-12992083, 490766935), F(6005943493, 2108947869), F(393006217, 1396673457), F(123872331, 1001029789)],
    [F(-1028468189, 846180014), 0, 0, F(8478235783, 508512852), F(1311729495, 1432422823),
     F(-10304129995, 1701304382), F(-48777925059, 3047939560), F(15336726248, 1032824649),
     F(-45442868181, 3398467696), F(3065993473, 597172653)],
    [F(185892177, 718116043), 0, 0, F(-3185094517, 667107341), F(-477755414, 1098053517),
     F(-703635378, 230739211), F(5731566787, 1027545527), F(5232866602, 850066563), F(-4093664535, 808688257),
     F(3962137247, 1805957418), F(65686358, 487910083)],
    [F(403863854, 491063109), 0, 0, F(-5068492393, 434740067), F(-411421997, 543043805), F(652783627, 914296604),
     F(11173962825, 925320556), F(-13158990841, 6184727034), F(3936647629, 1978049680), F(-160528059, 685178525),
     F(248638103, 1413531060), 0],
]
B_SOL = [F(14005451, 335480064), 0, 0, 0, 0, F(-59238493, 1068277825), F(181606767, 758867731),
         F(561292985, 797845732), F(-1041891430, 1371343529), F(760417239, 1151165299), F(118820643, 751138087),
         F(-528747749, 2220607170), F(1, 4), 0]
B_HAT = [F(13451932, 455176623), 0, 0, 0, 0, F(-808719846, 976000145), F(1757004468, 5645159321),
         F(656045339, 265891186), F(-3867574721, 1518517206), F(465885868, 322736535), F(53011238, 667516719),
         F(2, 45), 0, 0]
# FSAL: stage 14 is evaluated at y1, i.e. its row of A is B_SOL.
A.append([B_SOL[j] for j in range(13)])

N_STAGES = 14

# Dense output: y(theta) = y0 + sum_i b_i(theta) k_i,  b_i(theta) = sum_{m=1..6} DENSE_B[i][m-1] theta^m
# (diffrax ``_Dopri8Interpolation``: ``theta * polyval(eval_coeffs[i], theta)``).  Frozen output of
# ``python -m oracle.derive_dense_output`` (see that module for the derivation and its caveats), with
# the theta^1 column snapped to its exact value [i == 0].
DENSE_B = [
    [1.0, -6.691018173783315, 19.999006933368626, -30.061056828966635, 22.139650499809203, -6.344834939286462],
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    [0.0, -7.614265804585636, 52.22735327929503, -121.49996277313414, 116.44221495503238, -39.61079198522323],
    [0.0, 10.729828206995803, -46.89529333174393, 83.17378648512963, -67.14512895979595, 20.376120406616955],
    [0.0, 0.5010105848592362, -0.08581287575176573, 9.520887022608708, -16.567313740971997, 7.334739678660368],
    [0.0, 4.352779189231888, -35.97576607870187, 87.9744965874054, -89.99136937851385, 32.880100066764946],
    [0.0, -0.7606570723149959, 6.469460633091836, -17.512977453034445, 22.62357948159191, -10.15884255841119],
    [0.0, -1.2356641964061477, 10.26007488020532, -28.56236272873234, 32.23628249753497, -12.54014297009153],
    [0.0, 4.68718979313984, -33.31834766465918, 80.9399770753902, -82.10232756187993, 29.555398819256105],
    [0.0, -6.319348569127368, 46.32029783816455, -114.22497801652129, 116.2664567950579, -41.792428047573885],
    [0.0, 2.3501460419902416, -19.000973613268876, 50.25219062985484, -53.90204458786452, 20.30068152928823],
]


def dense_b() -> np.ndarray:
    return np.array(DENSE_B)


def dense_weights(theta: float) -> np.ndarray:
    """b_i(theta), i = 1..14."""
    th = np.array([theta**m for m in range(1, 7)])
    return dense_b() @ th


def a_matrix() -> np.ndarray:
    M = np.zeros((N_STAGES, N_STAGES))
    for i, row in enumerate(A):
        for j, v in enumerate(row):
            M[i, j] = float(v)
    return M


def b_sol() -> np.ndarray:
    return np.array([float(v) for v in B_SOL])


def b_err() -> np.ndarray:
    """y_error = h * sum (b_sol - b_hat)_i f_i  (diffrax ``b_error``)."""
    return np.array([float(F(s) - F(h)) for s, h in zip(B_SOL, B_HAT)])


def c_vec() -> np.ndarray:
    return np.array([float(v) for v in C])


# ----------------------------------------------------------------------------------------
# rooted trees and order conditions


@lru_cache(maxsize=None)
def trees(order: int):
    """All rooted trees of the given order as canonical nested tuples."""
    if order == 1:
        return ((),)
    out = set()

    def parts(n, maxpart):
        if n == 0:
            yield ()
            return
        for p in range(min(n, maxpart), 0, -1):
            for rest in parts(n - p, p):
                yield (p, *rest)

    def combos(sizes):
        if not sizes:
            yield ()
            return
        for t in trees(sizes[0]):
            for rest in combos(sizes[1:]):
                yield (t, *rest)

    for sizes in parts(order - 1, order - 1):
        for kids in combos(sizes):
            out.add(tuple(sorted(kids)))
    return tuple(sorted(out))


def _order(t):
    return 1 + sum(_order(k) for k in t)


def _gamma(t):
    g = _order(t)
    for k in t:
        g *= _gamma(k)
    return g


def _phi(t, Am):
    """Vector of elementary weights Phi_i(t)."""
    out = np.ones(Am.shape[0])
    for k in t:
        out = out * (Am @ _phi(k, Am))
    return out


def order_residuals(b, max_order, theta=1.0, Am=None):
    Am = a_matrix() if Am is None else Am
    res = []
    for p in range(1, max_order + 1):
        for t in trees(p):
            res.append((p, float(np.dot(b, _phi(t, Am)) - theta**p / _gamma(t))))
    return res


def verify(tol=5e-15):
    Am = a_matrix()
    rows = np.abs(Am.sum(axis=1) - c_vec()).max()
    r8 = max(abs(r) for _, r in order_residuals(b_sol(), 8))
    r7 = max(abs(r) for _, r in order_residuals(np.array([float(v) for v in B_HAT]), 7))
    r8hat = max(abs(r) for _, r in order_residuals(np.array([float(v) for v in B_HAT]), 8))
    assert rows < tol, rows
    assert r8 < tol, r8
    assert r7 < tol, r7
    assert r8hat > 1e-6  # the embedded solution really is only 7th order
    return {"row_sum": rows, "order8_sol": r8, "order7_hat": r7, "order8_hat": r8hat}


if __name__ == "__main__":
    print(verify())
