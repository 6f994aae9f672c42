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

}  // namespace gxs

// Called by gx_integrate_fixed (gx_kernels.cu) when GX_SCHEME_STRICT is set; arguments already validated there.
// gx_component.reserved carries the summation group: consecutive components with the same non-zero value are one
// reference component (an MN3 disk) and are summed first; 0 = a component of its own.
int gx_strict_integrate_fixed(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0,
                              double t1, double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps,
                              int32_t layout, double *q, double *p, int32_t *status, void *stream) {
    using namespace gxs;
    StrictPot P;
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
    StrictArgs a;
    a.q0 = q0; a.p0 = p0; a.ts = ts; a.q = q; a.p = p; a.status = status;
    a.N = N; a.max_steps = max_steps; a.t0 = t0; a.t1 = t1; a.dt0 = dt0; a.T = T; a.scheme = scheme;
    if (layout == GX_LAYOUT_T3N) { a.sn = 1; a.sk = 3 * N; a.sc = N; }
    else { a.sn = 3LL * T; a.sk = 3; a.sc = 1; }
    const int block = (N >= 148LL * 64 * 4) ? 64 : 32;
    gxs::k_integrate_fixed_strict<<<(int)((N + block - 1) / block), block, 0, (cudaStream_t)stream>>>(P, a);
    return cudaGetLastError() == cudaSuccess ? 0 : GX_ERR_CUDA;
}
