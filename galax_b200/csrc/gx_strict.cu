// gx_strict.cu -- reference-order ("strict") fixed-step integrator: GX_SCHEME_STRICT of gx_integrate_fixed.
//
// The fast kernels of gx_kernels.cu regroup the composite by kind, seed 1/x and 1/sqrt(x) from MUFU, read force tables
// and fuse the state update into FMAs: every one of those changes the last bit of a step, and after 10^4 steps of a
// chaotic-enough orbit the last bits are what decides the digits.  This translation unit is the other end of the
// trade: the arithmetic of the reference, operation by operation --
//   * per component in composite order, summed as AbstractCompositePotential._gradient does
//     (potential/_src/base_multi.py:48-55), an MN3 disk as the sum of its three Miyamoto-Nagai terms first
//     (builtin/mn3.py:121-130);
//   * IEEE division and square root, no floating-point contraction (this file is compiled with -fmad=false);
//   * log1p / exp / log from include/gx_portable_math.h (fdlibm, bit-reproducible on host and device);
//   * diffrax's SemiImplicitEuler update q1 = q0 + p0 h ; p1 = p0 + (-grad Phi(q1)) h as a multiply and an add, and
//     its time grid t_{n+1} = fl(t_n + dt0) walked step by step (dynamics/_src/orbit/field_hamiltonian.py:256-301).
// Its results are reproducible bit for bit by a plain C program on any IEEE-754 CPU; tests/test_gpu_strict.py holds it
// to exactly that for all 10^4 particles of config C1 in the three named Milky-Way models, and then measures the fast
// kernels against it.  A few times slower than the fast kernels (true divisions, 64-bit libm-style logarithm).
//
// Supported: static composites of Miyamoto-Nagai, Hernquist, NFW and PowerLawCutoff terms (the three named models
// and anything built from their parts).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/galax_b200.h"
#include "../../include/gx_portable_math.h"
// the Runge-Kutta tables (generic form: A, b_err, c, dense-output weights); a second copy of the constants, under
// another namespace so that the two translation units of the library do not define the same symbols
#define gx gxs_tab
#include "gx_tables.h"
#undef gx

namespace gxs {

constexpr double TINY = 2.2250738585072014e-308;

struct StrictComp {
    int kind, group;
    double p[3];
    double lg;  // PowerLawCutoff: lgamma(3/2 - alpha/2), from the host's libm (a constant of the potential)
};
struct StrictPot {
    int n;
    double G;
    StrictComp c[GX_MAX_COMPONENTS];
};

// regularised lower incomplete gamma P(a, x): series below x = a + 1, Lentz continued fraction above
__device__ double gammainc_P(double a, double lg, double x) {
    if (x <= 0.0) return 0.0;
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 1000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return sum * gx_pm_exp(-x + a * gx_pm_log(x) - lg);
    }
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

// ln(1+s) - s/(1+s); alternating series below s = 2^-4 where the difference cancels
__device__ double nfw_menc_shape(double s) {
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

__device__ void comp_gradient(double G, const StrictComp &c, double x, double y, double z, double g[3]) {
    const double *p = c.p;
    if (c.kind == GX_KIND_MIYAMOTO_NAGAI) {
        double b2 = p[2] * p[2];
        if (b2 == 0.0) b2 = TINY;  // Kuzmin (b = 0): zero z-force in the disk plane, as in oracle/galax_oracle.c
        double zeta = sqrt(z * z + b2);
        double D2 = x * x + y * y + (p[1] + zeta) * (p[1] + zeta);
        double f = G * p[0] / (D2 * sqrt(D2));
        g[0] = f * x;
        g[1] = f * y;
        g[2] = f * z * (p[1] + zeta) / zeta;
        return;
    }
    double r = sqrt(x * x + y * y + z * z + TINY), d1;
    double GM = G * p[0];
    if (c.kind == GX_KIND_HERNQUIST) {
        double u = r + p[1];
        d1 = GM / (u * u);
    } else if (c.kind == GX_KIND_NFW) {
        double rs = p[1], s = r / rs, m = nfw_menc_shape(s);
        d1 = GM * m / (r * r);
    } else {  // PowerLawCutoff: dPhi/dr = G M P(a, (r/r_c)^2) / r^2
        double a = 1.5 - p[1] / 2, rc = p[2];
        double s2 = (r / rc) * (r / rc);
        double P = gammainc_P(a, c.lg, s2);
        d1 = GM * P / (r * r);
    }
    double f = d1 / r;
    g[0] = f * x;
    g[1] = f * y;
    g[2] = f * z;
}

// composite: components of one group (an MN3 disk) are summed first, then the groups in order
__device__ void gradient(const StrictPot &P, double x, double y, double z, double out[3]) {
    double total[3] = {0, 0, 0}, sub[3] = {0, 0, 0}, v[3];
    bool have_total = false;
    for (int i = 0; i < P.n; ++i) {
        comp_gradient(P.G, P.c[i], x, y, z, v);
        const bool first = (i == 0) || (P.c[i].group != P.c[i - 1].group);
        for (int k = 0; k < 3; ++k) sub[k] = first ? v[k] : sub[k] + v[k];
        const bool last = (i == P.n - 1) || (P.c[i].group != P.c[i + 1].group);
        if (last) {
            for (int k = 0; k < 3; ++k) total[k] = have_total ? total[k] + sub[k] : sub[k];
            have_total = true;
        }
    }
    out[0] = total[0]; out[1] = total[1]; out[2] = total[2];
}

struct StrictArgs {
    const double *q0, *p0, *ts;
    double *q, *p;
    int *status;
    long long N, max_steps;
    long long sn, sk, sc;
    double t0, t1, dt0;
    int T, scheme;
};

__device__ __forceinline__ double clip_to_end(double tnext, double t1) { return (tnext > t1 - 1e-10) ? t1 : tnext; }

__global__ void __launch_bounds__(64) k_integrate_fixed_strict(const __grid_constant__ StrictPot P, const StrictArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const double dir = (a.t1 >= a.t0) ? 1.0 : -1.0;
    const double T0 = a.t0 * dir, T1 = a.t1 * dir, h0 = a.dt0 * dir;
    double q[3] = {a.q0[3 * i], a.q0[3 * i + 1], a.q0[3 * i + 2]};
    double p[3] = {a.p0[3 * i], a.p0[3 * i + 1], a.p0[3 * i + 2]};
    double qm[3] = {q[0], q[1], q[2]}, pm[3] = {p[0], p[1], p[2]}, tm = T0;  // LeapfrogMidpoint memory
    double tprev = T0, tnext = clip_to_end(T0 + h0, T1);
    double *qo = a.q + i * a.sn, *po = a.p + i * a.sn;
    int k = 0, st = GX_OK;
    long long n = 0;
    while (k < a.T && a.ts[k] * dir <= T0) {  // save times equal to t0 return y0
        for (int c = 0; c < 3; ++c) { qo[k * a.sk + c * a.sc] = q[c]; po[k * a.sk + c * a.sc] = p[c]; }
        ++k;
    }
    while (tprev < T1) {
        if (a.max_steps >= 0 && n >= a.max_steps) { st = GX_MAX_STEPS_REACHED; break; }
        const double h = tnext - tprev;
        double qn[3], pn[3], g[3];
        if (a.scheme == GX_SCHEME_SEMI_IMPLICIT_EULER) {
            for (int c = 0; c < 3; ++c) qn[c] = q[c] + (p[c] * dir) * h;
            gradient(P, qn[0], qn[1], qn[2], g);
            for (int c = 0; c < 3; ++c) pn[c] = p[c] + (-g[c] * dir) * h;
        } else {
            const double hh = tnext - tm;
            gradient(P, q[0], q[1], q[2], g);
            for (int c = 0; c < 3; ++c) {
                qn[c] = qm[c] + (p[c] * dir) * hh;
                pn[c] = pm[c] + (-g[c] * dir) * hh;
                qm[c] = q[c];
                pm[c] = p[c];
            }
            tm = tprev;
        }
        ++n;
        while (k < a.T && a.ts[k] * dir <= tnext) {  // LocalLinearInterpolation
            const double th = (a.ts[k] * dir - tprev) / (tnext - tprev);
            for (int c = 0; c < 3; ++c) {
                qo[k * a.sk + c * a.sc] = q[c] + th * (qn[c] - q[c]);
                po[k * a.sk + c * a.sc] = p[c] + th * (pn[c] - p[c]);
            }
            ++k;
        }
        for (int c = 0; c < 3; ++c) { q[c] = qn[c]; p[c] = pn[c]; }
        tprev = tnext;
        tnext = clip_to_end(tprev + h0, T1);
        if (!(isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2]) && isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) {
            st = GX_NONFINITE;
            break;
        }
    }
    const double NANV = __longlong_as_double(0x7ff8000000000000LL);
    for (; k < a.T; ++k)
        for (int c = 0; c < 3; ++c) { qo[k * a.sk + c * a.sc] = NANV; po[k * a.sk + c * a.sc] = NANV; }
    if (a.status) a.status[i] = st;
}


// ------------------------------------------------------------------------------------------------
// Dopri8 / Dopri5 + PID controller in the reference's own (generic, 6-component) form: k_i = f(y0 + sum_j a_ij k_j) h,
// error = sum_j e_j k_j, SaveAt from y0 + sum_j b_j(theta) k_j, factor = safety * (1/err)^(1/order) through the
// portable pow.  Operation for operation the loop of oracle/galax_oracle.c::oc_integrate_dopri8 (a restatement of
// diffrax's diffeqsolve + PIDController.adapt_step_size as galax drives them: dynamics/_src/legacy/integrator.py:161-245,
// dynamics/_src/orbit/solver.py:121-141): same results, same accept / reject sequence, bit for bit.
struct StrictDpArgs {
    const double *q0, *p0, *t0v, *ts;
    double *q, *p;
    int *status, *n_acc, *n_tot;
    long long N, max_steps;
    long long sn, sk, sc;
    double t0s, t1;
    double rtol, atol, pcoeff, icoeff, dcoeff, safety, factormin, factormax, dtmin, dtmax, dt0;
    int T;
};

__device__ __forceinline__ double rms6(const double v[6]) {
    double s = 0;
    for (int i = 0; i < 6; ++i) s += v[i] * v[i];
    return sqrt(s / 6.0);
}

__device__ void field_dir(const StrictPot &P, double dir, const double y[6], double f[6]) {
    double g[3];
    gradient(P, y[0], y[1], y[2], g);
    for (int c = 0; c < 3; ++c) { f[c] = y[3 + c] * dir; f[3 + c] = -g[c] * dir; }
}

__device__ double select_initial_step(const StrictPot &P, double dir, const double y0[6], const double f0[6], double rtol,
                                      double atol, double order) {
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
    field_dir(P, dir, y1, f1);
    for (int i = 0; i < 6; ++i) v[i] = (f1[i] - f0[i]) / sc[i];
    double d2 = rms6(v) / h0;
    double maxd = fmax(d1, d2);
    double h1 = (maxd <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : gx_pm_pow(0.01 / maxd, 1.0 / order);
    return fmin(100.0 * h0, h1);
}

__device__ __forceinline__ double clip_to_end_keep(double tprev, double tnext, double t1, int keep) {
    if (tnext > t1 - 1e-10) return keep ? t1 : tprev + 0.5 * (t1 - tprev);
    return tnext;
}

template <class TB>
__global__ void __launch_bounds__(64) k_integrate_adaptive_strict(const __grid_constant__ StrictPot P, const StrictDpArgs a) {
    constexpr int ns = TB::NS;
    const double order = TB::ORDER;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const double t0 = a.t0v ? a.t0v[i] : a.t0s;
    const double dir = (a.t1 >= t0) ? 1.0 : -1.0;
    const double T0 = t0 * dir, T1 = a.t1 * dir;
    double y[6] = {a.q0[3 * i], a.q0[3 * i + 1], a.q0[3 * i + 2], a.p0[3 * i], a.p0[3 * i + 1], a.p0[3 * i + 2]};
    double *qo = a.q + i * a.sn, *po = a.p + i * a.sn;
    int k = 0, st = GX_OK;
    int nacc = 0, ntot = 0;
    while (k < a.T && a.ts[k] * dir <= T0) {
        for (int c = 0; c < 3; ++c) { qo[k * a.sk + c * a.sc] = y[c]; po[k * a.sk + c * a.sc] = y[3 + c]; }
        ++k;
    }
    double f0[6];
    field_dir(P, dir, y, f0);
    double tprev = T0, tnext;
    double prev_inv = 1.0, prev_prev_inv = 1.0;
    int at_dtmin = 0;
    {
        double h = (a.dt0 > 0.0) ? a.dt0 : select_initial_step(P, dir, y, f0, a.rtol, a.atol, order + 1.0);
        if (a.dtmax > 0.0 && isfinite(a.dtmax)) h = fmin(h, a.dtmax);
        if (a.dtmin > 0.0) { at_dtmin = h <= a.dtmin; h = fmax(h, a.dtmin); }
        tnext = clip_to_end_keep(T0, T0 + h, T1, 1);
    }
    double K[ns][6], flast[6];
    while (tprev < T1) {
        if (a.max_steps >= 0 && ntot >= a.max_steps) { st = GX_MAX_STEPS_REACHED; break; }
        double h = tnext - tprev;
        for (int c = 0; c < 6; ++c) K[0][c] = f0[c] * h;
        double ys[6];
        for (int s = 1; s < ns; ++s) {
            for (int c = 0; c < 6; ++c) {
                double inc = 0.0;
                for (int j = 0; j < s; ++j) inc += TB::A(s, j) * K[j][c];
                ys[c] = y[c] + inc;
            }
            field_dir(P, dir, ys, flast);
            for (int c = 0; c < 6; ++c) K[s][c] = flast[c] * h;
        }
        double y1[6], err[6], sc[6];
        for (int c = 0; c < 6; ++c) {
            y1[c] = ys[c];
            double e = 0.0;
            for (int j = 0; j < ns; ++j) e += TB::E(j) * K[j][c];
            err[c] = e;
        }
        ++ntot;
        int nan1 = 0;
        for (int c = 0; c < 6; ++c) nan1 |= isnan(y1[c]);
        for (int c = 0; c < 6; ++c) {
            double yc = nan1 ? y[c] : y1[c];
            sc[c] = err[c] / (a.atol + fmax(fabs(y[c]), fabs(yc)) * a.rtol);
            if (isnan(sc[c])) sc[c] = __longlong_as_double(0x7ff0000000000000LL);
        }
        double serr = rms6(sc);
        int keep = serr < 1.0;
        if (a.dtmin > 0.0) keep = keep || at_dtmin;
        double inv = 1.0 / serr;
        double c1 = (a.icoeff + a.pcoeff + a.dcoeff) / order;
        double c2 = -(a.pcoeff + 2.0 * a.dcoeff) / order;
        double c3 = a.dcoeff / order;
        double fac1 = (c1 == 0.0) ? 1.0 : gx_pm_pow(inv, c1);
        double fac2 = (c2 == 0.0) ? 1.0 : gx_pm_pow(prev_inv, c2);
        double fac3 = (c3 == 0.0) ? 1.0 : gx_pm_pow(prev_prev_inv, c3);
        double fmin_ = keep ? 1.0 : a.factormin;
        double factor = a.safety * fac1 * fac2 * fac3;
        if (factor < fmin_) factor = fmin_;
        if (factor > a.factormax) factor = a.factormax;
        double dt = h * factor;
        if (inv == 0.0 || isinf(inv)) { inv = 1.0; prev_inv = 1.0; }
        if (a.dtmax > 0.0 && isfinite(a.dtmax)) dt = fmin(dt, a.dtmax);
        if (a.dtmin > 0.0) {
            at_dtmin = dt <= a.dtmin;
            dt = fmax(dt, a.dtmin);
        }
        if (keep) {
            while (k < a.T && a.ts[k] * dir <= tnext) {
                double th = (a.ts[k] * dir - tprev) / (tnext - tprev);
                double bw[ns];
                for (int j = 0; j < ns; ++j) {
                    double pv = TB::DB(j, 5);
                    for (int m = 4; m >= 0; --m) pv = pv * th + TB::DB(j, m);
                    bw[j] = pv * th;
                }
                for (int c = 0; c < 6; ++c) {
                    double inc = 0.0;
                    for (int j = 0; j < ns; ++j) inc += bw[j] * K[j][c];
                    double v = y[c] + inc;
                    if (c < 3) qo[k * a.sk + c * a.sc] = v; else po[k * a.sk + (c - 3) * a.sc] = v;
                }
                ++k;
            }
            for (int c = 0; c < 6; ++c) { y[c] = y1[c]; f0[c] = flast[c]; }
            prev_prev_inv = prev_inv;
            prev_inv = inv;
            tprev = tnext;
            ++nacc;
            int fin = 1;
            for (int c = 0; c < 6; ++c) fin &= isfinite(y[c]) ? 1 : 0;
            if (!fin) { st = GX_NONFINITE; break; }
        }
        if (tprev > T1) tprev = T1;
        tnext = clip_to_end_keep(tprev, tprev + dt, T1, keep);
    }
    const double NANV = __longlong_as_double(0x7ff8000000000000LL);
    for (; k < a.T; ++k)
        for (int c = 0; c < 3; ++c) { qo[k * a.sk + c * a.sc] = NANV; po[k * a.sk + c * a.sc] = NANV; }
    if (a.status) a.status[i] = st;
    if (a.n_acc) a.n_acc[i] = nacc;
    if (a.n_tot) a.n_tot[i] = ntot;
}

// tableau views for the strict kernels (generic A, not the Nystrom products the fast kernels use)
struct STabDp8 {
    static constexpr int NS = 14;
    static constexpr int ORDER = 8;
    __device__ __forceinline__ static double A(int i, int j) { return gxs_tab::dp8::A[i][j]; }
    __device__ __forceinline__ static double E(int j) { return gxs_tab::dp8::E[j]; }
    __device__ __forceinline__ static double DB(int j, int m) { return gxs_tab::dp8::DB[j][m]; }
};
struct STabDp5 {
    static constexpr int NS = 7;
    static constexpr int ORDER = 5;
    __device__ __forceinline__ static double A(int i, int j) { return gxs_tab::dp5::A[i][j]; }
    __device__ __forceinline__ static double E(int j) { return gxs_tab::dp5::E[j]; }
    __device__ __forceinline__ static double DB(int j, int m) { return gxs_tab::dp5::DB[j][m]; }
};

}  // namespace gxs

// gx_component.reserved carries the summation group: consecutive components with the same non-zero value are one
// reference component (an MN3 disk) and are summed first; 0 = a component of its own.
static int strict_pot(const gx_potential *pot, gxs::StrictPot &P) {
    P.n = pot->n;
    P.G = pot->G;
    int next_group = 1 << 20;
    for (int i = 0; i < pot->n; ++i) {
        const gx_component &c = pot->c[i];
        if (c.kind != GX_KIND_MIYAMOTO_NAGAI && c.kind != GX_KIND_HERNQUIST && c.kind != GX_KIND_NFW &&
            c.kind != GX_KIND_POWERLAWCUTOFF)
            return GX_ERR_UNSUPPORTED;
        if (c.reserved < 0) return GX_ERR_UNSUPPORTED;  // nested summation groups
        for (int k = 0; k < 8; ++k)
            if (c.dp[k] != 0.0) return GX_ERR_UNSUPPORTED;
        P.c[i].kind = c.kind;
        P.c[i].group = c.reserved != 0 ? c.reserved : next_group++;
        P.c[i].p[0] = c.p[0]; P.c[i].p[1] = c.p[1]; P.c[i].p[2] = c.p[2];
        P.c[i].lg = (c.kind == GX_KIND_POWERLAWCUTOFF) ? lgamma(1.5 - c.p[1] / 2) : 0.0;
    }
    return 0;
}

// Called by gx_integrate_fixed (gx_kernels.cu) when GX_SCHEME_STRICT is set; arguments already validated there.
int gx_strict_integrate_fixed(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0,
                              double t1, double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps,
                              int32_t layout, double *q, double *p, int32_t *status, void *stream) {
    using namespace gxs;
    StrictPot P;
    if (int rc = strict_pot(pot, P)) return rc;
    StrictArgs a;
    a.q0 = q0; a.p0 = p0; a.ts = ts; a.q = q; a.p = p; a.status = status;
    a.N = N; a.max_steps = max_steps; a.t0 = t0; a.t1 = t1; a.dt0 = dt0; a.T = T; a.scheme = scheme;
    if (layout == GX_LAYOUT_T3N) { a.sn = 1; a.sk = 3 * N; a.sc = N; }
    else { a.sn = 3LL * T; a.sk = 3; a.sc = 1; }
    const int block = (N >= 148LL * 64 * 4) ? 64 : 32;
    gxs::k_integrate_fixed_strict<<<(int)((N + block - 1) / block), block, 0, (cudaStream_t)stream>>>(P, a);
    return cudaGetLastError() == cudaSuccess ? 0 : GX_ERR_CUDA;
}

// Called by the adaptive entries (gx_kernels.cu) when GX_SOLVER_STRICT is or-ed into `solver`; arguments validated there.
int gx_strict_integrate_adaptive(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                                 const double *p0, int64_t N, const double *t0, double t0_scalar, double t1,
                                 const double *ts, int32_t T, int64_t max_steps, int32_t layout, double *q, double *p,
                                 int32_t *status, int32_t *n_accepted, int32_t *n_attempted, void *stream) {
    using namespace gxs;
    StrictPot P;
    if (int rc = strict_pot(pot, P)) return rc;
    StrictDpArgs a;
    a.q0 = q0; a.p0 = p0; a.t0v = t0; a.ts = ts; a.q = q; a.p = p;
    a.status = status; a.n_acc = n_accepted; a.n_tot = n_attempted;
    a.N = N; a.max_steps = max_steps; a.t0s = t0_scalar; a.t1 = t1;
    a.rtol = pid->rtol; a.atol = pid->atol;
    a.pcoeff = pid->pcoeff; a.icoeff = pid->icoeff; a.dcoeff = pid->dcoeff;
    a.safety = pid->safety; a.factormin = pid->factormin; a.factormax = pid->factormax;
    a.dtmin = pid->dtmin; a.dtmax = pid->dtmax;
    a.dt0 = (pid->dt0 > 0.0) ? pid->dt0 : -1.0;
    a.T = T;
    if (layout == GX_LAYOUT_T3N) { a.sn = 1; a.sk = 3 * N; a.sc = N; }
    else { a.sn = 3LL * T; a.sk = 3; a.sc = 1; }
    const int block = (N >= 148LL * 64 * 4) ? 64 : 32;
    const int grid = (int)((N + block - 1) / block);
    if (solver == GX_SOLVER_DOPRI5) k_integrate_adaptive_strict<STabDp5><<<grid, block, 0, (cudaStream_t)stream>>>(P, a);
    else k_integrate_adaptive_strict<STabDp8><<<grid, block, 0, (cudaStream_t)stream>>>(P, a);
    return cudaGetLastError() == cudaSuccess ? 0 : GX_ERR_CUDA;
}
