/* galax_oracle.c -- plain-C CPU restatement of the galax hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * What it restates (reference file:line under /root/reference/src/galax unless noted):
 *   potentials    potential/_src/builtin/{miyamotonagai.py:73-78, hernquist.py:79-82,
 *                 nfw/base.py:326-338, powerlawcutoff.py:88-117}, r = safe_sqrt(x^2+y^2+z^2+tiny)
 *                 (potential/_src/utils.py:44-125); composite = sum in component order
 *                 (potential/_src/base_multi.py:39-82).  Gradients/Hessians are the closed-form
 *                 derivatives of those potentials (the reference uses jax.grad / jax.hessian).
 *   field         (dq/dt, dp/dt) = (p, -grad Phi(q))   dynamics/_src/orbit/field_hamiltonian.py:240-249
 *   fixed step    diffrax 0.7.0 SemiImplicitEuler + ConstantStepSize, reached through
 *                 field_hamiltonian.py:256-301 and orbit/solver.py:774-803
 *   adaptive      diffrax 0.7.0 Dopri8 + PIDController (legacy/integrator.py:37-39,161-168;
 *                 orbit/solver.py:121-141), per particle
 *   mock stream   legacy/mockstream/df/fardal15.py:49-94, df/chen24.py:61-137,
 *                 cluster/radius.py:198-215, register_api.py:77-88,
 *                 legacy/mockstream/mockstream_generator.py:128-158
 *
 * diffrax / jax are third-party and not vendored in /root/reference: the solver loop, PID rule,
 * initial-step heuristic, _clip_to_end and SaveAt interpolation are restated from the published
 * algorithm -- "parity unpinned" at the 1e-12 level, see oracle/__init__.py.
 *
 * Deliberately unoptimised: libm sqrt/log1p, true divisions, no FMA contraction
 * (compiled with -ffp-contract=off).  OpenMP over particles only, for the CPU-baseline timing.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OC_KIND_MN 0
#define OC_KIND_HERNQUIST 1
#define OC_KIND_NFW 2
#define OC_KIND_PLC 3
#define OC_KIND_LOG 4
#define OC_KIND_ISOCHRONE 5
#define OC_KIND_SATOH 6
#define OC_KIND_TRIAXIAL_HERNQUIST 7 /* p = (m_tot, r_s, q1, q2)      builtin/hernquist.py:160-176 */
#define OC_KIND_JAFFE 8              /* p = (m_tot, r_s)              builtin/jaffe.py:50-60 */
#define OC_KIND_BURKERT 9            /* p = (m, r_s)                  builtin/burkert.py:197-227 */
#define OC_KIND_STONE 10             /* p = (m_tot, r_c, r_h)         builtin/stoneostriker15.py:150-160 */
#define OC_KIND_HARMONIC 11          /* p = (omega_x, omega_y, omega_z) builtin/example.py:75-85 */
#define OC_KIND_HENON_HEILES 12      /* p = (coeff, timescale)        builtin/example.py:159-176 */
#define OC_PI 3.14159265358979323846
#define OC_BURKERT_C (3.0 * 0.69314718055994530942 - OC_PI / 2.0)
#define OC_MAX_COMP 16

#define OC_TINY 2.2250738585072014e-308

/* log, exp and log1p on the gradient path are the portable fdlibm forms of include/gx_portable_math.h (a public
 * header of the library, not product kernels): glibc's and CUDA's own functions are each < 1 ulp but round
 * differently, and the strict-parity test (tests/test_gpu_strict.py) compares the reference-order GPU kernel with this
 * file BIT FOR BIT.  Accuracy of the portable forms against 50-digit mpmath: tests/test_portable_math.py. */
#include "../include/gx_portable_math.h"
void oc_pm_eval(int op, int64_t n, const double *x, double *out) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = (op == 0) ? gx_pm_log(x[i]) : (op == 1) ? gx_pm_exp(x[i]) : (op == 2) ? gx_pm_log1p(x[i]) : gx_pm_pow(x[i], 0.125);
}

typedef struct {
    int kind;
    int group; /* components with equal group id are summed first (MN3 disk) */
    double p[8];
    double dp[8]; /* LinearParameter (potential/_src/params/core.py:25-110): p_k(t) = p[k] + dp[k] t */
} oc_component;

typedef struct {
    int n;
    double G;
    oc_component c[OC_MAX_COMP];
} oc_potential;

/* ---------------------------------------------------------------- incomplete gamma */

/* regularised lower incomplete gamma P(a,x): series for x < a+1, Lentz continued fraction else */
static double oc_gammainc_P(double a, double x) {
    if (x <= 0.0) return 0.0;
    double lg = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 1000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return sum * gx_pm_exp(-x + a * gx_pm_log(x) - lg);
    } else {
        const double FPMIN = 1e-300;
        double b = x + 1.0 - a, c = 1.0 / FPMIN, d = 1.0 / b, h = d;
        for (int i = 1; i < 1000; ++i) {
            double an = -i * (i - a);
            b += 2.0;
            d = an * d + b;
            if (fabs(d) < FPMIN) d = FPMIN;
            c = b + an / c;
            if (fabs(c) < FPMIN) c = FPMIN;
            d = 1.0 / d;
            double del = d * c;
            h *= del;
            if (fabs(del - 1.0) < 1e-17) break;
        }
        return 1.0 - gx_pm_exp(-x + a * gx_pm_log(x) - lg) * h;
    }
}

double oc_gammainc(double a, double x) { return oc_gammainc_P(a, x); }

/* ---------------------------------------------------------------- single components */

/* b^2 under the root of zeta = sqrt(z^2 + b^2) in the DERIVATIVES of Miyamoto-Nagai / Satoh terms: for b = 0
 * (KuzminPotential, builtin/kuzmin.py:82-84, Phi = -GM / sqrt(R^2 + (|z| + a)^2)) the closed forms divide by zeta, which
 * vanishes in the disk plane; the reference's autodiff gives d|z|/dz = 0 there, i.e. a zero z-force and no delta
 * function in the Hessian.  The smallest normal number under the root reproduces exactly that; b > 0 is unchanged. */
static inline double oc_b2(double b) {
    double b2 = b * b;
    return b2 == 0.0 ? OC_TINY : b2;
}

static double nfw_menc_shape(double s) {
    /* ln(1+s) - s/(1+s); alternating series below s = 2^-4 where the difference cancels */
    if (s < 0.0625) {
        double ser = 0.0;
        for (int k = 18; k >= 2; --k) {
            double c = ((k & 1) ? -1.0 : 1.0) * (k - 1.0) / k;
            ser = ser * s + c;
        }
        return ser * s * s;
    }
    return gx_pm_log1p(s) - s / (1.0 + s);
}

static double comp_potential(double G, const oc_component *c, const double q[3]) {
    const double x = q[0], y = q[1], z = q[2];
    const double *p = c->p;
    switch (c->kind) {
    case OC_KIND_MN: {
        double R2 = x * x + y * y;
        double zp = sqrt(z * z + p[2] * p[2]) + p[1];
        return -G * p[0] / sqrt(R2 + zp * zp);
    }
    case OC_KIND_HERNQUIST: {
        double r = sqrt(x * x + y * y + z * z + OC_TINY);
        return -G * p[0] / (r + p[1]);
    }
    case OC_KIND_NFW: {
        double r = sqrt(x * x + y * y + z * z + OC_TINY);
        double s = r / p[1];
        double phi0 = -G * p[0] / p[1];
        return phi0 * log1p(s) / s;
    }
    case OC_KIND_PLC: {
        double r = sqrt(x * x + y * y + z * z + OC_TINY);
        double ah = p[1] / 2, rc = p[2];
        double s2 = (r / rc) * (r / rc);
        double GM = G * p[0];
        double ga = 1.5 - ah;
        double t1 = GM * (oc_gammainc_P(ga, s2) * tgamma(ga)) * (ah - 1.5) / (r * tgamma(2.5 - ah));
        double t2 = GM * (oc_gammainc_P(1 - ah, s2) * tgamma(1 - ah)) / (rc * tgamma(ga));
        double pinf = ga > 0 ? GM * tgamma(1 - ah) / (rc * tgamma(ga)) : 0.0;
        return t1 + t2 - pinf;
    }
    case OC_KIND_LOG: {
        double sp = sin(p[5]), cp = cos(p[5]);
        double xr = x * cp + y * sp, yr = -x * sp + y * cp;
        double r2 = (xr / p[2]) * (xr / p[2]) + (yr / p[3]) * (yr / p[3]) + (z / p[4]) * (z / p[4]);
        return 0.5 * p[0] * p[0] * log(p[1] * p[1] + r2);
    }
    case OC_KIND_ISOCHRONE: {
        double r = sqrt(x * x + y * y + z * z + OC_TINY);
        return -G * p[0] / (p[1] + sqrt(r * r + p[1] * p[1]));
    }
    case OC_KIND_SATOH:
        return -G * p[0] / sqrt(x * x + y * y + z * z + p[1] * (p[1] + 2.0 * sqrt(z * z + p[2] * p[2])));
    case OC_KIND_TRIAXIAL_HERNQUIST: {
        double m = sqrt(x * x + (y / p[2]) * (y / p[2]) + (z / p[3]) * (z / p[3]) + OC_TINY);
        return -G * p[0] / (m + p[1]);
    }
    case OC_KIND_JAFFE: {
        double r = sqrt(x * x + y * y + z * z + OC_TINY);
        return -G * p[0] / p[1] * log(1.0 + p[1] / r);
    }
    case OC_KIND_BURKERT: {
        double r = sqrt(x * x + y * y + z * z + OC_TINY);
        double s = r / p[1], si = 1.0 / s;
        return -G * p[0] / (p[1] * OC_BURKERT_C) *
               (OC_PI - 2.0 * (1.0 + si) * atan(s) + 2.0 * (1.0 + si) * log1p(s) - (1.0 - si) * log1p(s * s));
    }
    case OC_KIND_STONE: {
        double r = sqrt(x * x + y * y + z * z + OC_TINY), rc = p[1], rh = p[2];
        double A = -2.0 * G * p[0] / (OC_PI * (rh - rc));
        return A * ((rh * atan2(r, rh) - rc * atan2(r, rc)) / r + 0.5 * log((r * r + rh * rh) / (r * r + rc * rc)));
    }
    case OC_KIND_HARMONIC:
        return 0.5 * ((p[0] * x) * (p[0] * x) + (p[1] * y) * (p[1] * y) + (p[2] * z) * (p[2] * z));
    case OC_KIND_HENON_HEILES:
        return ((x * x + y * y) / 2.0 + p[0] * (x * x * y - y * y * y / 3.0)) / (p[1] * p[1]);
    }
    return NAN;
}

static void log_matrix(const double *p, double M[3][3]) {
    double sp = sin(p[5]), cp = cos(p[5]);
    double i1 = 1.0 / (p[2] * p[2]), i2 = 1.0 / (p[3] * p[3]), i3 = 1.0 / (p[4] * p[4]);
    M[0][0] = cp * cp * i1 + sp * sp * i2;
    M[1][1] = sp * sp * i1 + cp * cp * i2;
    M[0][1] = M[1][0] = cp * sp * (i1 - i2);
    M[2][2] = i3;
    M[0][2] = M[2][0] = M[1][2] = M[2][1] = 0.0;
}

/* Burkert: 2 ln(1+s) + ln(1+s^2) - 2 atan(s), series below s = 0.3 (see oracle/potentials.py:_burkert_B) */
static double burkert_B(double s) {
    if (s < 0.3) {
        double s4 = (s * s) * (s * s), e = 0.0;
        for (int n = 8; n >= 0; --n) e = e * s4 + (1.0 / (double)(4 * n + 3) - s / (double)(4 * n + 4));
        return 4.0 * (s * s * s) * e;
    }
    return 2.0 * log1p(s) + log1p(s * s) - 2.0 * atan(s);
}

/* Stone-Ostriker: r_h atan(r/r_h) - r_c atan(r/r_c), Taylor series (16 terms) below r = 0.3 r_c */
static double stone_T(double r, double rc, double rh) {
    if (r < 0.3 * rc) {
        double uc = (r / rc) * (r / rc), uh = (r / rh) * (r / rh), acc = 0.0, pc = 1.0, ph = 1.0;
        for (int k = 1; k <= 16; ++k) {
            pc *= uc;
            ph *= uh;
            double term = (ph - pc) / (double)(2 * k + 1);
            acc += (k & 1) ? -term : term;
        }
        return r * acc;
    }
    return rh * atan2(r, rh) - rc * atan2(r, rc);
}

/* radial derivatives of a spherical component: d1 = dPhi/dr, d2 = d2Phi/dr2 */
static void comp_radial(double G, const oc_component *c, double r, double *d1, double *d2) {
    const double *p = c->p;
    double GM = G * p[0];
    switch (c->kind) {
    case OC_KIND_HERNQUIST: {
        double u = r + p[1];
        *d1 = GM / (u * u);
        *d2 = -2.0 * GM / (u * u * u);
        return;
    }
    case OC_KIND_NFW: {
        double rs = p[1], s = r / rs, m = nfw_menc_shape(s);
        *d1 = GM * m / (r * r);
        *d2 = GM * (s / (rs * (1.0 + s) * (1.0 + s) * r * r) - 2.0 * m / (r * r * r));
        return;
    }
    case OC_KIND_PLC: {
        double a = 1.5 - p[1] / 2, rc = p[2];
        double s2 = (r / rc) * (r / rc);
        double P = oc_gammainc_P(a, s2);
        double dP = pow(s2, a - 1.0) * exp(-s2) / tgamma(a);
        *d1 = GM * P / (r * r);
        *d2 = GM * (dP * 2.0 * r / (rc * rc) / (r * r) - 2.0 * P / (r * r * r));
        return;
    }
    case OC_KIND_TRIAXIAL_HERNQUIST: { /* r is the ellipsoidal radius here */
        double u = r + p[1];
        *d1 = GM / (u * u);
        *d2 = -2.0 * GM / (u * u * u);
        return;
    }
    case OC_KIND_JAFFE: {
        double a = p[1], ra = r * (r + a);
        *d1 = GM / ra;
        *d2 = -GM * (2.0 * r + a) / (ra * ra);
        return;
    }
    case OC_KIND_BURKERT: {
        double rs = p[1], s = r / rs, k = GM / OC_BURKERT_C;
        *d1 = k * burkert_B(s) / (r * r);
        *d2 = k * 4.0 * s * s / ((1.0 + s) * (1.0 + s * s) * rs * r * r) - 2.0 * (*d1) / r;
        return;
    }
    case OC_KIND_STONE: {
        double rc = p[1], rh = p[2], A = 2.0 * GM / (OC_PI * (rh - rc));
        *d1 = A * stone_T(r, rc, rh) / (r * r);
        *d2 = A * (rh * rh - rc * rc) / ((r * r + rh * rh) * (r * r + rc * rc)) - 2.0 * (*d1) / r;
        return;
    }
    case OC_KIND_ISOCHRONE: {
        double b = p[1], a = sqrt(r * r + b * b);
        *d1 = GM * r / (a * (b + a) * (b + a));
        *d2 = GM * (1.0 / (a * (b + a) * (b + a)) - r * r / (a * a * a * (b + a) * (b + a)) -
                    2.0 * r * r / (a * a * (b + a) * (b + a) * (b + a)));
        return;
    }
    }
    *d1 = *d2 = NAN;
}

static void comp_gradient(double G, const oc_component *c, const double q[3], double g[3]) {
    const double x = q[0], y = q[1], z = q[2];
    if (c->kind == OC_KIND_LOG) {
        double M[3][3], Mx[3];
        log_matrix(c->p, M);
        for (int i = 0; i < 3; ++i) Mx[i] = M[i][0] * x + M[i][1] * y + M[i][2] * z;
        double D = c->p[1] * c->p[1] + x * Mx[0] + y * Mx[1] + z * Mx[2];
        for (int i = 0; i < 3; ++i) g[i] = c->p[0] * c->p[0] * Mx[i] / D;
        return;
    }
    if (c->kind == OC_KIND_SATOH) {
        const double *p = c->p;
        double zeta = sqrt(z * z + oc_b2(p[2]));
        double S = x * x + y * y + z * z + p[1] * (p[1] + 2.0 * zeta);
        double f = G * p[0] / (S * sqrt(S));
        g[0] = f * x;
        g[1] = f * y;
        g[2] = f * z * (1.0 + p[1] / zeta);
        return;
    }
    if (c->kind == OC_KIND_MN) {
        const double *p = c->p;
        double zeta = sqrt(z * z + oc_b2(p[2]));
        double D2 = x * x + y * y + (p[1] + zeta) * (p[1] + zeta);
        double f = G * p[0] / (D2 * sqrt(D2));
        g[0] = f * x;
        g[1] = f * y;
        g[2] = f * z * (p[1] + zeta) / zeta;
        return;
    }
    if (c->kind == OC_KIND_HARMONIC) {
        g[0] = c->p[0] * c->p[0] * x;
        g[1] = c->p[1] * c->p[1] * y;
        g[2] = c->p[2] * c->p[2] * z;
        return;
    }
    if (c->kind == OC_KIND_HENON_HEILES) {
        double k = c->p[0], it2 = 1.0 / (c->p[1] * c->p[1]);
        g[0] = (x + 2.0 * k * x * y) * it2;
        g[1] = (y + k * (x * x - y * y)) * it2;
        g[2] = 0.0;
        return;
    }
    /* profiles in the (ellipsoidal) radius: grad = (F'/m) (x, y/q1^2, z/q2^2) */
    double i1 = 1.0, i2 = 1.0;
    if (c->kind == OC_KIND_TRIAXIAL_HERNQUIST) {
        i1 = 1.0 / (c->p[2] * c->p[2]);
        i2 = 1.0 / (c->p[3] * c->p[3]);
    }
    double r = sqrt(x * x + y * y * i1 + z * z * i2 + OC_TINY), d1, d2;
    comp_radial(G, c, r, &d1, &d2);
    double f = d1 / r;
    g[0] = f * x;
    g[1] = f * (y * i1);
    g[2] = f * (z * i2);
}

static void comp_hessian(double G, const oc_component *c, const double q[3], double H[9]) {
    const double x = q[0], y = q[1], z = q[2];
    if (c->kind == OC_KIND_LOG) {
        double M[3][3], Mx[3];
        log_matrix(c->p, M);
        for (int i = 0; i < 3; ++i) Mx[i] = M[i][0] * x + M[i][1] * y + M[i][2] * z;
        double D = c->p[1] * c->p[1] + x * Mx[0] + y * Mx[1] + z * Mx[2], vc2 = c->p[0] * c->p[0];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) H[3 * i + j] = vc2 * (M[i][j] / D - 2.0 * Mx[i] * Mx[j] / (D * D));
        return;
    }
    if (c->kind == OC_KIND_SATOH) {
        const double *p = c->p;
        double a = p[1], b = p[2];
        double zeta = sqrt(z * z + oc_b2(b));
        double S = x * x + y * y + z * z + a * (a + 2.0 * zeta);
        double f3 = G * p[0] / (S * sqrt(S)), f5 = 3.0 * f3 / S;
        double u[3] = {x, y, z * (1.0 + a / zeta)};
        double duz = 1.0 + (b != 0.0 ? a * b * b / (zeta * zeta * zeta) : 0.0);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) H[3 * i + j] = -f5 * (u[i] * u[j]);
        H[0] += f3;
        H[4] += f3;
        H[8] += f3 * duz;
        return;
    }
    if (c->kind == OC_KIND_MN) {
        const double *p = c->p;
        double a = p[1], b = p[2];
        double zeta = sqrt(z * z + oc_b2(b));
        double D2 = x * x + y * y + (a + zeta) * (a + zeta);
        double D = sqrt(D2);
        double f3 = G * p[0] / (D2 * D);
        double f5 = 3.0 * G * p[0] / (D2 * D2 * D);
        double u[3] = {x, y, z * (a + zeta) / zeta};
        double duz = 1.0 + (b != 0.0 ? a * b * b / (zeta * zeta * zeta) : 0.0);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) H[3 * i + j] = -f5 * (u[i] * u[j]);
        H[0] += f3;
        H[4] += f3;
        H[8] += f3 * duz;
        return;
    }
    if (c->kind == OC_KIND_HARMONIC) {
        for (int i = 0; i < 9; ++i) H[i] = 0.0;
        H[0] = c->p[0] * c->p[0];
        H[4] = c->p[1] * c->p[1];
        H[8] = c->p[2] * c->p[2];
        return;
    }
    if (c->kind == OC_KIND_HENON_HEILES) {
        double k = c->p[0], it2 = 1.0 / (c->p[1] * c->p[1]);
        for (int i = 0; i < 9; ++i) H[i] = 0.0;
        H[0] = (1.0 + 2.0 * k * y) * it2;
        H[1] = H[3] = 2.0 * k * x * it2;
        H[4] = (1.0 - 2.0 * k * y) * it2;
        return;
    }
    double dg[3] = {1.0, 1.0, 1.0};
    if (c->kind == OC_KIND_TRIAXIAL_HERNQUIST) {
        dg[1] = 1.0 / (c->p[2] * c->p[2]);
        dg[2] = 1.0 / (c->p[3] * c->p[3]);
    }
    double r = sqrt(x * x + y * y * dg[1] + z * z * dg[2] + OC_TINY), d1, d2;
    comp_radial(G, c, r, &d1, &d2);
    double n[3] = {x / r, y * dg[1] / r, z * dg[2] / r}; /* w / m */
    double f = d1 / r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double nn = n[i] * n[j];
            H[3 * i + j] = (d2 - f) * nn + (i == j ? f * dg[i] : 0.0);
        }
}

/* ---------------------------------------------------------------- time dependence */

static int is_time_dependent(const oc_potential *P) {
    for (int i = 0; i < P->n; ++i)
        for (int k = 0; k < 8; ++k)
            if (P->c[i].dp[k] != 0.0) return 1;
    return 0;
}

/* the potential frozen at time t (the reference evaluates every parameter as param(t) inside _potential) */
static void potential_at(const oc_potential *P, double t, oc_potential *out) {
    *out = *P;
    for (int i = 0; i < P->n; ++i)
        for (int k = 0; k < 8; ++k) {
            out->c[i].p[k] = P->c[i].p[k] + P->c[i].dp[k] * t;
            out->c[i].dp[k] = 0.0;
        }
}

/* ---------------------------------------------------------------- composite (sum in order) */

double oc_potential_value(const oc_potential *P, const double q[3]) {
    double total = 0.0, sub = 0.0;
    int have_total = 0;
    for (int i = 0; i < P->n; ++i) {
        double v = comp_potential(P->G, &P->c[i], q);
        int first = (i == 0) || (P->c[i].group != P->c[i - 1].group);
        sub = first ? v : sub + v;
        int last = (i == P->n - 1) || (P->c[i].group != P->c[i + 1].group);
        if (last) {
            total = have_total ? total + sub : sub;
            have_total = 1;
        }
    }
    return total;
}

static void sum_vec(const oc_potential *P, const double q[3], int len, double *out,
                    void (*fn)(double, const oc_component *, const double *, double *)) {
    double total[9], sub[9], v[9];
    int have_total = 0;
    for (int i = 0; i < P->n; ++i) {
        fn(P->G, &P->c[i], q, v);
        int first = (i == 0) || (P->c[i].group != P->c[i - 1].group);
        for (int k = 0; k < len; ++k) sub[k] = first ? v[k] : sub[k] + v[k];
        int last = (i == P->n - 1) || (P->c[i].group != P->c[i + 1].group);
        if (last) {
            for (int k = 0; k < len; ++k) total[k] = have_total ? total[k] + sub[k] : sub[k];
            have_total = 1;
        }
    }
    for (int k = 0; k < len; ++k) out[k] = total[k];
}

void oc_gradient(const oc_potential *P, const double q[3], double g[3]) { sum_vec(P, q, 3, g, comp_gradient); }
void oc_hessian(const oc_potential *P, const double q[3], double H[9]) { sum_vec(P, q, 9, H, comp_hessian); }

/* bulk evaluation: what bit 0 = Phi, 1 = grad, 2 = acceleration(-grad), 3 = Hessian */
void oc_potential_eval_t(const oc_potential *P0, double t, int64_t N, const double *xyz, unsigned what, double *phi,
                         double *grad, double *acc, double *hess) {
    oc_potential Pt;
    potential_at(P0, t, &Pt);
    const oc_potential *P = &Pt;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const double *q = xyz + 3 * i;
        if (what & 1u) phi[i] = oc_potential_value(P, q);
        if (what & 6u) {
            double g[3];
            oc_gradient(P, q, g);
            if (what & 2u) { grad[3 * i] = g[0]; grad[3 * i + 1] = g[1]; grad[3 * i + 2] = g[2]; }
            if (what & 4u) { acc[3 * i] = -g[0]; acc[3 * i + 1] = -g[1]; acc[3 * i + 2] = -g[2]; }
        }
        if (what & 8u) oc_hessian(P, q, hess + 9 * i);
    }
}

void oc_potential_eval(const oc_potential *P, int64_t N, const double *xyz, unsigned what, double *phi,
                       double *grad, double *acc, double *hess) {
    oc_potential_eval_t(P, 0.0, N, xyz, what, phi, grad, acc, hess);
}

/* ---------------------------------------------------------------- field */

/* acceleration at physical time t */
static inline void accel(const oc_potential *P, int td, double t, const double q[3], double a[3]) {
    double g[3];
    if (td) {
        oc_potential Pt;
        potential_at(P, t, &Pt);
        oc_gradient(&Pt, q, g);
    } else {
        oc_gradient(P, q, g);
    }
    a[0] = -g[0];
    a[1] = -g[1];
    a[2] = -g[2];
}

/* status codes shared with the CUDA library (include/galax_b200.h) */
#define OC_OK 0
#define OC_MAX_STEPS 1
#define OC_NONFINITE 2

static inline double clip_to_end(double tprev, double tnext, double t1, int keep) {
    /* diffrax _clip_to_end, fp64 tolerance 1e-10 */
    if (tnext > t1 - 1e-10) return keep ? t1 : tprev + 0.5 * (t1 - tprev);
    return tnext;
}

/* ---------------------------------------------------------------- fixed step
 * scheme 0: SemiImplicitEuler   q1 = q0 + p0*h ; p1 = p0 + (-grad Phi(q1))*h
 * scheme 1: LeapfrogMidpoint    y_{n+1} = y_{n-1} + f(y_n) * (t_{n+1} - t_{n-1}), first step Euler
 * Time is accumulated (tnext = tprev + dt0), the last step is clipped to t1, save times that are
 * not step boundaries are linearly interpolated (LocalLinearInterpolation).
 * Integration runs in tau = dir*t so t1 < t0 works (diffrax flips the sign the same way).
 * Outputs q_out,p_out are [N,T,3]. */
int oc_integrate_fixed(const oc_potential *P, int64_t N, const double *q0, const double *p0, double t0,
                       double t1, double dt0, int T, const double *ts, int scheme, int64_t max_steps,
                       double *q_out, double *p_out, int32_t *status, int64_t *nsteps) {
    double dir = (t1 >= t0) ? 1.0 : -1.0;
    double T0 = t0 * dir, T1 = t1 * dir;
    double h0 = dt0 * dir;
    if (!(h0 > 0.0) && T1 > T0) return -1;
    const int td = is_time_dependent(P);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < N; ++i) {
        double q[3] = {q0[3 * i], q0[3 * i + 1], q0[3 * i + 2]};
        double p[3] = {p0[3 * i], p0[3 * i + 1], p0[3 * i + 2]};
        double qm[3], pm[3], tm = T0; /* leapfrog-midpoint memory (t_{n-1}, y_{n-1}) */
        memcpy(qm, q, sizeof q);
        memcpy(pm, p, sizeof p);
        double tprev = T0, tnext = clip_to_end(T0, T0 + h0, T1, 1);
        int k = 0, st = OC_OK;
        int64_t n = 0;
        double *qo = q_out + (int64_t)3 * T * i, *po = p_out + (int64_t)3 * T * i;
        /* save times equal to t0 are y0 */
        while (k < T && ts[k] * dir <= T0) {
            for (int c = 0; c < 3; ++c) { qo[3 * k + c] = q[c]; po[3 * k + c] = p[c]; }
            ++k;
        }
        while (tprev < T1) {
            if (max_steps >= 0 && n >= max_steps) { st = OC_MAX_STEPS; break; }
            double h = tnext - tprev;
            double qn[3], pn[3], a[3];
            if (scheme == 0) {
                for (int c = 0; c < 3; ++c) qn[c] = q[c] + (p[c] * dir) * h;
                accel(P, td, tprev * dir, qn, a); /* diffrax SemiImplicitEuler: both terms are evaluated at t0 of the step */
                for (int c = 0; c < 3; ++c) pn[c] = p[c] + (a[c] * dir) * h;
            } else {
                double hh = tnext - tm;
                accel(P, td, tprev * dir, q, a);
                for (int c = 0; c < 3; ++c) {
                    qn[c] = qm[c] + (p[c] * dir) * hh;
                    pn[c] = pm[c] + (a[c] * dir) * hh;
                }
                memcpy(qm, q, sizeof q);
                memcpy(pm, p, sizeof p);
                tm = tprev;
            }
            ++n;
            while (k < T && ts[k] * dir <= tnext) {
                double th = (ts[k] * dir - tprev) / (tnext - tprev);
                for (int c = 0; c < 3; ++c) {
                    qo[3 * k + c] = q[c] + th * (qn[c] - q[c]);
                    po[3 * k + c] = p[c] + th * (pn[c] - p[c]);
                }
                ++k;
            }
            memcpy(q, qn, sizeof q);
            memcpy(p, pn, sizeof p);
            tprev = tnext;
            tnext = clip_to_end(tprev, tprev + h0, T1, 1);
            if (!(isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2]) && isfinite(p[0]) && isfinite(p[1]) &&
                  isfinite(p[2]))) { st = OC_NONFINITE; break; }
        }
        for (; k < T; ++k) /* unreachable save times (error paths): NaN like an unfilled diffrax buffer (inf) */
            for (int c = 0; c < 3; ++c) { qo[3 * k + c] = NAN; po[3 * k + c] = NAN; }
        if (status) status[i] = st;
        if (nsteps) nsteps[i] = n;
    }
    return 0;
}

/* ---------------------------------------------------------------- Dopri8 + PID, per particle */

/* explicit FSAL Runge-Kutta pair: Dopri8 (ns = 14, order 8) or Dopri5 (ns = 7, order 5); the last stage sits at y1 */
typedef struct {
    int ns;
    double order; /* diffrax error_order: exponent 1/order in the step-size rule and the initial-step heuristic */
    double a[14][14];
    double b_sol[14];
    double b_err[14];
    double c[14];
    double dense[14][6]; /* b_i(theta) = sum_m dense[i][m-1] theta^m */
} oc_tableau;

typedef struct {
    double rtol, atol;
    double pcoeff, icoeff, dcoeff;
    double safety, factormin, factormax;
    double dtmin, dtmax; /* <= 0 / inf: unset */
    int force_dtmin;
    double dt0; /* <= 0 or NaN: Hairer initial-step heuristic */
} oc_pid;

static inline double rms6(const double v[6]) {
    double s = 0;
    for (int i = 0; i < 6; ++i) s += v[i] * v[i];
    return sqrt(s / 6.0);
}

static void field_dir(const oc_potential *P, int td, double dir, double tau, const double y[6], double f[6]) {
    double a[3];
    accel(P, td, tau * dir, y, a);
    for (int c = 0; c < 3; ++c) { f[c] = y[3 + c] * dir; f[3 + c] = a[c] * dir; }
}

static double select_initial_step(const oc_potential *P, int td, double dir, double tau0, const double y0[6],
                                  const double f0[6], double rtol, double atol, double order) {
    double sc[6], v[6];
    for (int i = 0; i < 6; ++i) sc[i] = atol + fabs(y0[i]) * rtol;
    for (int i = 0; i < 6; ++i) v[i] = y0[i] / sc[i];
    double d0 = rms6(v);
    for (int i = 0; i < 6; ++i) v[i] = f0[i] / sc[i];
    double d1 = rms6(v);
    int cond = (d0 < 1e-5) || (d1 < 1e-5);
    double h0 = cond ? 1e-6 : 0.01 * (d0 / d1);
    double y1[6], f1[6];
    for (int i = 0; i < 6; ++i) y1[i] = y0[i] + h0 * f0[i];
    field_dir(P, td, dir, tau0 + h0, y1, f1);
    for (int i = 0; i < 6; ++i) v[i] = (f1[i] - f0[i]) / sc[i];
    double d2 = rms6(v) / h0;
    double maxd = fmax(d1, d2);
    double h1 = (maxd <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : gx_pm_pow(0.01 / maxd, 1.0 / order);
    return fmin(100.0 * h0, h1);
}

/* t0v: per-particle start times (stride t0_stride = 0 broadcasts a scalar).  Saves at ts[0..T) must lie
 * in [t0_i, t1]; q_out,p_out are [N,T,3].  n_acc / n_tot: accepted / attempted steps per particle. */
int oc_integrate_dopri8(const oc_potential *P, const oc_tableau *tab, const oc_pid *pid, int64_t N,
                        const double *q0, const double *p0, const double *t0v, int t0_stride, double t1,
                        int T, const double *ts, int64_t max_steps, double *q_out, double *p_out,
                        int32_t *status, int32_t *n_acc, int32_t *n_tot) {
    const double order = tab->order;
    const int ns = tab->ns;
    const int td = is_time_dependent(P);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < N; ++i) {
        double t0 = t0v[(int64_t)t0_stride * i];
        double dir = (t1 >= t0) ? 1.0 : -1.0;
        double T0 = t0 * dir, T1 = t1 * dir;
        double y[6] = {q0[3 * i], q0[3 * i + 1], q0[3 * i + 2], p0[3 * i], p0[3 * i + 1], p0[3 * i + 2]};
        double *qo = q_out + (int64_t)3 * T * i, *po = p_out + (int64_t)3 * T * i;
        int k = 0, st = OC_OK;
        int32_t nacc = 0, ntot = 0;
        while (k < T && ts[k] * dir <= T0) {
            for (int c = 0; c < 3; ++c) { qo[3 * k + c] = y[c]; po[3 * k + c] = y[3 + c]; }
            ++k;
        }
        double f0[6];
        field_dir(P, td, dir, T0, y, f0);
        double tprev = T0, tnext;
        double prev_inv = 1.0, prev_prev_inv = 1.0;
        int at_dtmin = 0;
        {
            /* PIDController.init: heuristic when dt0 is None, then clamp to [dtmin, dtmax].  The heuristic's exponent
             * is 1/(error_order + 1) (Hairer II.4 with p = order): inferred from the reference's 8-digit OrbitSolver
             * doctests, which this reproduces to 4e-9 (1/error_order: 6e-8, i.e. a different first step). */
            double h = (pid->dt0 > 0.0) ? pid->dt0
                                        : select_initial_step(P, td, dir, T0, y, f0, pid->rtol, pid->atol, order + 1.0);
            if (pid->dtmax > 0.0 && isfinite(pid->dtmax)) h = fmin(h, pid->dtmax);
            if (pid->dtmin > 0.0) { at_dtmin = h <= pid->dtmin; h = fmax(h, pid->dtmin); }
            tnext = clip_to_end(T0, T0 + h, T1, 1);
        }
        double K[14][6], flast[6];
        while (tprev < T1) {
            if (max_steps >= 0 && ntot >= max_steps) { st = OC_MAX_STEPS; break; }
            double h = tnext - tprev;
            for (int c = 0; c < 6; ++c) K[0][c] = f0[c] * h;
            double ys[6];
            for (int s = 1; s < ns; ++s) {
                for (int c = 0; c < 6; ++c) {
                    double inc = 0.0;
                    for (int j = 0; j < s; ++j) inc += tab->a[s][j] * K[j][c];
                    ys[c] = y[c] + inc;
                }
                field_dir(P, td, dir, tprev + tab->c[s] * h, ys, flast);
                for (int c = 0; c < 6; ++c) K[s][c] = flast[c] * h;
            }
            /* a[ns-1][:] == b_sol, so ys is y1 and K[ns-1] = f(y1) h (FSAL) */
            double y1[6], err[6], sc[6];
            for (int c = 0; c < 6; ++c) {
                y1[c] = ys[c];
                double e = 0.0;
                for (int j = 0; j < ns; ++j) e += tab->b_err[j] * K[j][c];
                err[c] = e;
            }
            ++ntot;
            int nan1 = 0;
            for (int c = 0; c < 6; ++c) nan1 |= isnan(y1[c]);
            for (int c = 0; c < 6; ++c) {
                double yc = nan1 ? y[c] : y1[c];
                sc[c] = err[c] / (pid->atol + fmax(fabs(y[c]), fabs(yc)) * pid->rtol);
                if (isnan(sc[c])) sc[c] = INFINITY;
            }
            double serr = rms6(sc);
            int keep = serr < 1.0;
            if (pid->dtmin > 0.0) keep = keep || at_dtmin;
            double inv = 1.0 / serr;
            double c1 = (pid->icoeff + pid->pcoeff + pid->dcoeff) / order;
            double c2 = -(pid->pcoeff + 2.0 * pid->dcoeff) / order;
            double c3 = pid->dcoeff / order;
            /* x^c through the portable exp / log (include/gx_portable_math.h): libm's pow differs between platforms in
             * the last bit, and the last bit of the factor decides accept / reject sequences (tests/test_gpu_strict.py) */
            double fac1 = (c1 == 0.0) ? 1.0 : gx_pm_pow(inv, c1);
            double fac2 = (c2 == 0.0) ? 1.0 : gx_pm_pow(prev_inv, c2);
            double fac3 = (c3 == 0.0) ? 1.0 : gx_pm_pow(prev_prev_inv, c3);
            double fmin_ = keep ? 1.0 : pid->factormin;
            double factor = pid->safety * fac1 * fac2 * fac3;
            if (factor < fmin_) factor = fmin_;
            if (factor > pid->factormax) factor = pid->factormax;
            double dt = h * factor;
            if (inv == 0.0 || isinf(inv)) { inv = 1.0; prev_inv = 1.0; }
            if (pid->dtmax > 0.0 && isfinite(pid->dtmax)) dt = fmin(dt, pid->dtmax);
            if (pid->dtmin > 0.0) {
                at_dtmin = dt <= pid->dtmin;
                dt = fmax(dt, pid->dtmin);
            }
            if (keep) {
                /* SaveAt(ts): y(theta) = y0 + sum_i b_i(theta) k_i  (diffrax _Dopri8Interpolation) */
                while (k < T && ts[k] * dir <= tnext) {
                    double th = (ts[k] * dir - tprev) / (tnext - tprev);
                    double bw[14];
                    for (int j = 0; j < ns; ++j) {
                        double pv = tab->dense[j][5];
                        for (int m = 4; m >= 0; --m) pv = pv * th + tab->dense[j][m];
                        bw[j] = pv * th;
                    }
                    for (int c = 0; c < 6; ++c) {
                        double inc = 0.0;
                        for (int j = 0; j < ns; ++j) inc += bw[j] * K[j][c];
                        double v = y[c] + inc;
                        if (c < 3) qo[3 * k + c] = v; else po[3 * k + c - 3] = v;
                    }
                    ++k;
                }
                /* FSAL: diffrax carries f(t1, y1) itself (not k = f*h) into the next step */
                for (int c = 0; c < 6; ++c) { y[c] = y1[c]; f0[c] = flast[c]; }
                prev_prev_inv = prev_inv;
                prev_inv = inv;
                tprev = tnext;
                ++nacc;
                int fin = 1;
                for (int c = 0; c < 6; ++c) fin &= isfinite(y[c]);
                if (!fin) { st = OC_NONFINITE; break; }
            }
            if (tprev > T1) tprev = T1;
            tnext = clip_to_end(tprev, tprev + dt, T1, keep);
        }
        for (; k < T; ++k)
            for (int c = 0; c < 3; ++c) { qo[3 * k + c] = NAN; po[3 * k + c] = NAN; }
        if (status) status[i] = st;
        if (n_acc) n_acc[i] = nacc;
        if (n_tot) n_tot[i] = ntot;
    }
    return 0;
}

/* ---------------------------------------------------------------- stream release (DF) */

static inline void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* King (1962) tidal radius cbrt(G m / (Omega^2 - d2Phi/dr2))   cluster/radius.py:198-215 */
static double tidal_radius(const oc_potential *P, const double x[3], const double v[3], double mass,
                           double *omega_out) {
    double r = norm3(x), L[3], H[9];
    cross3(x, v, L);
    double om[3] = {L[0] / (r * r), L[1] / (r * r), L[2] / (r * r)};
    double omega = norm3(om);
    oc_hessian(P, x, H);
    double rh[3] = {x[0] / r, x[1] / r, x[2] / r};
    double d2 = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) d2 += rh[i] * H[3 * i + j] * rh[j];
    if (omega_out) *omega_out = omega;
    return cbrt(P->G * mass / (omega * omega - d2));
}

/* Fardal+15 release (df/fardal15.py:49-94).  normals is [4,M] = (n_kr, n_kvphi, n_kz, n_kvz). */
void oc_release_fardal(const oc_potential *P, int64_t M, const double *xq, const double *xp, const double *mass,
                       const double *normals, double *q_lead, double *p_lead, double *q_trail, double *p_trail) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; ++i) {
        const double *x = xq + 3 * i, *v = xp + 3 * i;
        double omega, rt = tidal_radius(P, x, v, mass[i], &omega);
        double r = norm3(x), rhat[3], L[3], zhat[3], phiv[3], phihat[3];
        for (int c = 0; c < 3; ++c) rhat[c] = x[c] / r;
        double vc = omega * rt;
        cross3(x, v, L);
        double Ln = norm3(L);
        for (int c = 0; c < 3; ++c) zhat[c] = L[c] / Ln;
        double vr = v[0] * rhat[0] + v[1] * rhat[1] + v[2] * rhat[2];
        for (int c = 0; c < 3; ++c) phiv[c] = v[c] - vr * rhat[c];
        double pn = norm3(phiv);
        for (int c = 0; c < 3; ++c) phihat[c] = phiv[c] / pn;
        double kr = 2.0 + normals[0 * M + i] * 0.5;
        double kvphi = kr * (0.3 + normals[1 * M + i] * 0.5);
        double kz = 0.0 + normals[2 * M + i] * 0.5;
        double kvz = 0.0 + normals[3 * M + i] * 0.5;
        for (int c = 0; c < 3; ++c) {
            q_trail[3 * i + c] = x[c] + rt * (kr * rhat[c] + kz * zhat[c]);
            p_trail[3 * i + c] = v[c] + vc * (kvphi * phihat[c] + kvz * zhat[c]);
            q_lead[3 * i + c] = x[c] - rt * (kr * rhat[c] - kz * zhat[c]);
            p_lead[3 * i + c] = v[c] - vc * (kvphi * phihat[c] - kvz * zhat[c]);
        }
    }
}

/* Chen+24 release (df/chen24.py:61-137).  posvel is [M,6], the multivariate-normal draws. */
void oc_release_chen(const oc_potential *P, int64_t M, const double *xq, const double *xp, const double *mass,
                     const double *posvel, double *q_lead, double *p_lead, double *q_trail, double *p_trail) {
    const double D2R = 0.017453292519943295;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; ++i) {
        const double *x = xq + 3 * i, *v = xp + 3 * i, *pv = posvel + 6 * i;
        double rt = tidal_radius(P, x, v, mass[i], 0);
        double r = norm3(x), xh[3], L[3], zh[3], phiv[3], yh[3];
        for (int c = 0; c < 3; ++c) xh[c] = x[c] / r;
        cross3(x, v, L);
        double Ln = norm3(L);
        for (int c = 0; c < 3; ++c) zh[c] = L[c] / Ln;
        double vr = v[0] * xh[0] + v[1] * xh[1] + v[2] * xh[2];
        for (int c = 0; c < 3; ++c) phiv[c] = v[c] - vr * xh[c];
        double pn = norm3(phiv);
        for (int c = 0; c < 3; ++c) yh[c] = phiv[c] / pn;
        double Dr = pv[0] * rt;
        double vesc = sqrt(2.0 * P->G * mass[i] / Dr);
        double Dv = pv[3] * vesc;
        double phi = pv[1] * D2R, theta = pv[2] * D2R, alpha = pv[4] * D2R, beta = pv[5] * D2R;
        double ct = cos(theta), st = sin(theta), cp = cos(phi), sp = sin(phi);
        double ca = cos(alpha), sa = sin(alpha), cb = cos(beta), sb = sin(beta);
        for (int c = 0; c < 3; ++c) {
            q_trail[3 * i + c] = x[c] + (Dr * ct * cp) * xh[c] + (Dr * ct * sp) * yh[c] + (Dr * st) * zh[c];
            p_trail[3 * i + c] = v[c] + (Dv * cb * ca) * xh[c] + (Dv * cb * sa) * yh[c] + (Dv * sb) * zh[c];
            q_lead[3 * i + c] = x[c] - (Dr * ct * cp) * xh[c] - (Dr * ct * sp) * yh[c] + (Dr * st) * zh[c];
            p_lead[3 * i + c] = v[c] - (Dv * cb * ca) * xh[c] - (Dv * cb * sa) * yh[c] + (Dv * sb) * zh[c];
        }
    }
}

void oc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oc_num_threads(void) {
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
#endif
    return n;
}
