"""CPU: include/gx_portable_math.h (fdlibm log / exp / log1p, the bit-reproducible functions shared by the
reference-order kernels and the C oracle) against 50-digit mpmath: each < 1 ulp."""
import ctypes as C

import mpmath as mp
import numpy as np
import pytest

from oracle import cref

mp.mp.dps = 50


def _eval(op, x):
    L = cref.lib()
    L.oc_pm_eval.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    L.oc_pm_eval(op, len(x), x.ctypes.data, out.ctypes.data)
    return out


def _worst_ulp(op, x, f):
    out = _eval(op, x)
    worst = 0.0
    for xi, oi in zip(x, out):
        t = f(mp.mpf(float(xi)))
        worst = max(worst, abs(float((mp.mpf(float(oi)) - t) / mp.mpf(float(np.spacing(abs(float(t))))))))
    return worst


@pytest.mark.parametrize(
    "op,f,xs",
    [
        (0, mp.log, lambda r: np.concatenate([np.exp(r.uniform(-700, 700, 3000)), 1 + r.uniform(-0.3, 0.5, 3000),
                                              1 + r.uniform(-1, 1, 500) * 2.0**-21])),
        (1, mp.exp, lambda r: np.concatenate([r.uniform(-700, 700, 3000), r.uniform(-2, 2, 3000)])),
        (2, mp.log1p, lambda r: np.concatenate([np.exp(r.uniform(-40, 12, 3000)), r.uniform(0.0625, 8, 3000),
                                                -r.uniform(0, 0.99, 1000)])),
    ],
)  # fmt: skip
def test_within_one_ulp(op, f, xs):
    assert _worst_ulp(op, xs(np.random.default_rng(op)), f) < 1.0


def test_special_values():
    assert _eval(0, [1.0])[0] == 0.0 and _eval(1, [0.0])[0] == 1.0 and _eval(2, [0.0])[0] == 0.0
    assert np.isneginf(_eval(0, [0.0])[0]) and np.isnan(_eval(0, [-1.0])[0]) and np.isposinf(_eval(0, [np.inf])[0])
    assert _eval(1, [-800.0])[0] == 0.0 and np.isposinf(_eval(1, [800.0])[0]) and np.isnan(_eval(1, [np.nan])[0])
    assert np.isneginf(_eval(2, [-1.0])[0]) and np.isnan(_eval(2, [-2.0])[0])
    # subnormal argument of log, subnormal result of exp
    assert abs(_eval(0, [5e-324])[0] / float(mp.log(mp.mpf(5e-324))) - 1) < 1e-15
    assert abs(_eval(1, [-740.0])[0] / float(mp.exp(-740)) - 1) < 1e-9
