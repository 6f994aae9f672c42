"""Generate ``gx_tables.h``: Dopri8 (Prince-Dormand RK8(7)13M + FSAL) tables in Nystrom form.

The Hamiltonian field is (dq/dt, dp/dt) = (p, a(q)), so a stage's position only needs the stage
accelerations a_l:   q_i = q0 + c_i h p0 + h^2 sum_l AA[i][l] a_l,   AA = A @ A.
Likewise q1 / the q error / the q dense output use  b^T A,  e^T A  and  B(theta)^T A.
All products are formed in exact rational arithmetic and rounded once.

The tableau itself is typed in here independently of ``oracle/dopri8_tableau.py`` (the oracle is test
infrastructure and is not imported by product code); ``tests/test_tables.py`` checks that the two agree
and that this table satisfies the order conditions.

Run:  python galax_b200/csrc/gen_tables.py
"""
from fractions import Fraction as F
from pathlib import Path

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
A.append([B_SOL[j] for j in range(13)])  # FSAL stage
C = [F(0), F(1, 18), F(1, 12), F(1, 8), F(5, 16), F(3, 8), F(59, 400), F(93, 200), F(5490023248, 9719169821),
     F(13, 20), F(1201146811, 1299019798), F(1), F(1), F(1)]

# degree-6 continuous extension, b_i(theta) = sum_m DENSE_B[i][m-1] theta^m (see DESIGN.md "Dense output")
DENSE_B = [
    [1.0, -6.691018173783315, 19.999006933368626, -30.061056828966635, 22.139650499809203, -6.344834939286462],
    [0.0] * 6, [0.0] * 6, [0.0] * 6, [0.0] * 6,
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

N = 14

# ---- Dormand-Prince 5(4), 7 stages FSAL (diffrax.Dopri5; the reference's experimental StreamSimulator default,
# dynamics/_src/experimental/stream.py:32-41).  Dense output = diffrax's FourthOrderPolynomialInterpolation with
# Shampine's mid-point weights, rewritten as stage weights b_i(theta) (see oracle/dopri5_tableau.py).
A5 = [
    [],
    [F(1, 5)],
    [F(3, 40), F(9, 40)],
    [F(44, 45), F(-56, 15), F(32, 9)],
    [F(19372, 6561), F(-25360, 2187), F(64448, 6561), F(-212, 729)],
    [F(9017, 3168), F(-355, 33), F(46732, 5247), F(49, 176), F(-5103, 18656)],
    [F(35, 384), F(0), F(500, 1113), F(125, 192), F(-2187, 6784), F(11, 84)],
]
B5_SOL = [F(35, 384), F(0), F(500, 1113), F(125, 192), F(-2187, 6784), F(11, 84), F(0)]
# the embedded weights diffrax (like torchdiffeq) uses: b_sol - b_hat is 2/3 of the textbook Dormand-Prince error
# coefficients (71/57600, ...).  Settled by the reference's StreamSimulator doctest, which this reproduces to 2e-9 (the
# textbook estimate takes 11 % shorter steps and lands 4e-3 kpc away after 8 Gyr).
B5_HAT = [F(1951, 21600), F(0), F(22642, 50085), F(451, 720), F(-12231, 42400), F(649, 6300), F(1, 60)]
C5 = [F(0), F(1, 5), F(3, 10), F(4, 5), F(8, 9), F(1), F(1)]
C5_MID = [F(6025192743, 30085553152) / 2, F(0), F(51252292925, 65400821598) / 2, F(-2691868925, 45128329728) / 2,
          F(187940372067, 1594534317056) / 2, F(-1776094331, 19743644256) / 2, F(11237099, 235043384) / 2]


def dense5():
    out = []
    for i in range(7):
        d0, d6 = F(int(i == 0)), F(int(i == 6))
        b, cm = B5_SOL[i], C5_MID[i]
        out.append([d0, d6 - 4 * d0 - 5 * b + 16 * cm, 5 * d0 - 3 * d6 + 14 * b - 32 * cm,
                    2 * d6 - 2 * d0 - 8 * b + 16 * cm, F(0), F(0)])
    return out


def amat(Arows=None, n=N):
    Arows = A if Arows is None else Arows
    return [[F(Arows[i][j]) if j < len(Arows[i]) else F(0) for j in range(n)] for i in range(n)]


def tables(which="dp8"):
    if which == "dp8":
        n, Am, bs, bh, cs = N, amat(), B_SOL, B_HAT, C
        DB = [[F(v) for v in row] for row in DENSE_B]
    else:
        n, Am, bs, bh, cs = 7, amat(A5, 7), B5_SOL, B5_HAT, C5
        DB = dense5()
    AA = [[sum(Am[i][j] * Am[j][l] for j in range(n)) for l in range(n)] for i in range(n)]
    e = [F(s) - F(h) for s, h in zip(bs, bh)]
    eA = [sum(e[j] * Am[j][l] for j in range(n)) for l in range(n)]
    DQ = [[sum(DB[j][m] * Am[j][l] for j in range(n)) for m in range(6)] for l in range(n)]
    return dict(A=Am, AA=AA, C=cs, B=[F(b) for b in bs], E=e, EA=eA, DB=DB, DQ=DQ)


def fmt(v):
    return repr(float(v))


def emit_namespace(out, ns, order, t):
    n = len(t["B"])
    out.append(f"namespace {ns} {{")
    out.append(f"constexpr int NS = {n};")

    # Values live in __constant__ memory (an unrolled access becomes a c[bank][imm] operand of the DFMA, no
    # extra instruction); the sparsity pattern is constexpr so zero coefficients vanish at compile time.
    def mat(name, M, cols):
        out.append(f"__constant__ double {name}[{len(M)}][{cols}] = {{")
        for row in M:
            out.append("    {" + ", ".join(fmt(v) for v in row) + "},")
        out.append("};")
        out.append(f"__device__ constexpr bool {name}_NZ[{len(M)}][{cols}] = {{")
        for row in M:
            out.append("    {" + ", ".join("true" if float(v) != 0.0 else "false" for v in row) + "},")
        out.append("};")

    def vec(name, v):
        out.append(f"__constant__ double {name}[{len(v)}] = {{" + ", ".join(fmt(x) for x in v) + "};")
        out.append(f"__device__ constexpr bool {name}_NZ[{len(v)}] = {{" + ", ".join("true" if float(x) != 0.0 else "false" for x in v) + "};")

    mat("A", t["A"], n)      # p-stage weights (only the last row, b_sol, is used by the kernels)
    mat("AA", t["AA"], n)    # q_i = q0 + CN[i] h p0 + h^2 sum_l AA[i][l] a_l
    vec("CN", t["C"])
    vec("B", t["B"])         # p1 = p0 + h sum_l B[l] a_l
    vec("E", t["E"])         # err_p = h sum_l E[l] a_l
    vec("EA", t["EA"])       # err_q = h^2 sum_l EA[l] a_l
    mat("DB", t["DB"], 6)    # p(theta) = p0 + h sum_l (sum_m DB[l][m] theta^(m+1)) a_l
    mat("DQ", t["DQ"], 6)    # q(theta) = q0 + theta h p0 + h^2 sum_l (sum_m DQ[l][m] theta^(m+1)) a_l
    out.append(f"}}  // namespace {ns}")
    # compile-time view used by the solver-templated kernels
    out.append(f"struct Tab{ns.capitalize()} {{")
    out.append(f"    static constexpr int NS = {n};")
    out.append(f"    static constexpr int ORDER = {order};  // diffrax error_order")
    for name in ("AA", "DB", "DQ"):
        out.append(f"    __device__ __forceinline__ static double {name}(int i, int j) {{ return {ns}::{name}[i][j]; }}")
        out.append(f"    __host__ __device__ static constexpr bool {name}_NZ(int i, int j) {{ return {ns}::{name}_NZ[i][j]; }}")
    for name in ("CN", "B", "E", "EA"):
        out.append(f"    __device__ __forceinline__ static double {name}(int i) {{ return {ns}::{name}[i]; }}")
        out.append(f"    __host__ __device__ static constexpr bool {name}_NZ(int i) {{ return {ns}::{name}_NZ[i]; }}")
    out.append("};")


LOG_TAB_BITS = 8


def log_table(n_candidates=3000):
    """(inv_c_j, -ln(inv_c_j)) for the 2^LOG_TAB_BITS mantissa intervals [1 + j/256, 1 + (j+1)/256).

    Gal's accurate-table method: among the doubles next to 1/(interval midpoint), inv_c_j is the one whose
    -ln(inv_c_j) lies closest to a double (typically < 2^-11 ulp away), so the tabulated logarithm carries no
    rounding error worth mentioning and ln m = logc_j + log1p(m inv_c_j - 1) holds to ~2^-64 for the table.
    Deterministic (fixed candidate window); takes about a minute of mpmath."""
    import mpmath as mp
    import numpy as np

    mp.mp.dps = 40
    n = 1 << LOG_TAB_BITS
    rows = []
    for j in range(n):
        x0 = float(1 / (mp.mpf(1) + (mp.mpf(j) + mp.mpf("0.5")) / n))
        best = None
        x = x0
        for _ in range(n_candidates // 2):
            x = np.nextafter(x, 0.0)
        for _ in range(n_candidates):
            lv = -mp.log(mp.mpf(float(x)))
            lf = float(lv)
            err = abs(lv - mp.mpf(lf)) / mp.mpf(np.spacing(lf) if lf != 0 else 1.0)
            if best is None or err < best[0]:
                best = (err, float(x), lf)
            x = np.nextafter(x, 2.0)
        rows.append((best[1], best[2], float(best[0])))
    return rows


def emit_log_table(out):
    out.append("// log table of gx_math.cuh::log_pos: {inv_c_j, -ln(inv_c_j)}, j = top %d mantissa bits" % LOG_TAB_BITS)
    out.append("constexpr int LOG_TAB_BITS = %d;" % LOG_TAB_BITS)
    out.append("__device__ const double2 LOG_TAB[%d] = {" % (1 << LOG_TAB_BITS))
    rows = log_table()
    out.append("// (largest distance of a tabulated logarithm from its exact value, rows j >= 4 -- the only ones reachable with")
    out.append("//  k = 0, where the entry's own ulp matters: %.2e ulp)" % max(r[2] for r in rows[4:]))
    for a, b, _ in rows:
        out.append("    {%s, %s}," % (fmt(a), fmt(b)))
    out.append("};")


def emit():
    out = ["// GENERATED by galax_b200/csrc/gen_tables.py -- do not edit.",
           "// Dopri8 = Prince-Dormand RK8(7)13M + FSAL stage; Dopri5 = Dormand-Prince 5(4) + FSAL stage;",
           "// both in Nystrom form for (dq, dp) = (p, a(q)).",
           "#pragma once", "namespace gx {"]
    emit_namespace(out, "dp8", 8, tables("dp8"))
    emit_namespace(out, "dp5", 5, tables("dp5"))
    emit_log_table(out)
    out.append("}  // namespace gx")
    Path(__file__).with_name("gx_tables.h").write_text("\n".join(out) + "\n")


if __name__ == "__main__":
    emit()
