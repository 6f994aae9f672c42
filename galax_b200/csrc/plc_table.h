// plc_table.h -- host-side construction and per-device cache of the PowerLawCutoff force table.
//
// G(s) = P(a, s^2) / s^3 (s = r / r_c, P the regularised lower incomplete gamma function) is what the
// PowerLawCutoff force needs: Phi'(r)/r = (G M / r_c^3) G(s)  (reference: builtin/powerlawcutoff.py:88-117, whose
// gradient collapses to G M P(a, s^2) / r^2).  The direct evaluation (log, exp, a ~60-term series) made
// BovyMWPotential2014 ten times slower than MilkyWayPotential, so G is tabulated once per exponent a: 32 intervals
// per octave of s over [2^-11, 2^3), a degree-9 polynomial each (relative error < 3e-16, checked at build time
// against the long-double series on a finer grid; degree 13 on 8 intervals per octave, the first layout, is no more
// accurate and costs 4 more FMAs and 9 more loads per evaluation).  Outside the range the kernels fall back to the
// series.
//
// The table depends only on a; it is built in long double on the host, uploaded once per (device, a) and kept for
// the life of the process (immutable after creation, so sharing it between streams and threads is safe).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <cmath>

#include <mutex>
#include <vector>

#include <string.h>

#include "../../include/galax_b200.h"
#include "gx_potential.cuh"

namespace gx {

// P(a, x) in long double: positive-term series (x < a + 40) -- enough for x = s^2 < 64.
static long double plc_P_ld(long double a, long double x) {
    if (x <= 0.0L) return 0.0L;
    long double ap = a, del = 1.0L / a, sum = del;
    for (int n = 0; n < 4000; ++n) {
        ap += 1.0L;
        del *= x / ap;
        sum += del;
        if (del < sum * 1e-22L) break;
    }
    long double v = sum * expl(-x + a * logl(x) - lgammal(a));
    return v > 1.0L ? 1.0L : v;
}

static long double plc_G_ld(long double a, long double s) { return plc_P_ld(a, s * s) / (s * s * s); }

// Chebyshev interpolation on [-1, 1] at PLC_DEG+1 nodes, converted to monomial coefficients in long double.
template <int n = PLC_DEG + 1, int NCHECK = 40, class F>
static void fit_interval(const F &fun, long double s0, long double s1, double *coef, double *max_rel_err) {
    const long double PI = 3.141592653589793238462643383279502884L;
    long double f[n], c[n];
    for (int k = 0; k < n; ++k) {
        long double tk = cosl(PI * (k + 0.5L) / n);
        f[k] = fun(0.5L * (s0 + s1) + 0.5L * (s1 - s0) * tk);
    }
    for (int j = 0; j < n; ++j) {  // Chebyshev coefficients
        long double sum = 0.0L;
        for (int k = 0; k < n; ++k) sum += f[k] * cosl(PI * j * (k + 0.5L) / n);
        c[j] = 2.0L * sum / n;
    }
    c[0] *= 0.5L;
    // Chebyshev -> monomial: accumulate T_j(t) by the three-term recurrence on coefficient vectors
    long double mono[n] = {0}, Tm1[n] = {0}, Tm0[n] = {0}, Tn[n];
    Tm1[0] = 1.0L;  // T_0
    Tm0[1] = 1.0L;  // T_1
    for (int i = 0; i < n; ++i) mono[i] += c[0] * Tm1[i];
    if (n > 1)
        for (int i = 0; i < n; ++i) mono[i] += c[1] * Tm0[i];
    for (int j = 2; j < n; ++j) {
        for (int i = 0; i < n; ++i) Tn[i] = -Tm1[i];
        for (int i = 0; i + 1 < n; ++i) Tn[i + 1] += 2.0L * Tm0[i];
        for (int i = 0; i < n; ++i) { mono[i] += c[j] * Tn[i]; Tm1[i] = Tm0[i]; Tm0[i] = Tn[i]; }
    }
    for (int i = 0; i < n; ++i) coef[i] = (double)mono[i];
    // verify in double Horner (what the device does) on a fine grid
    double worst = 0.0;
    for (int k = 0; k <= NCHECK; ++k) {
        double t = -1.0 + 2.0 * k / (double)NCHECK;
        double v = coef[n - 1];
        for (int i = n - 2; i >= 0; --i) v = fma(v, t, coef[i]);
        long double s = 0.5L * (s0 + s1) + 0.5L * (s1 - s0) * (long double)t;
        long double ref = fun(s);
        double rel = (double)fabsl(((long double)v - ref) / ref);
        if (rel > worst) worst = rel;
    }
    if (max_rel_err && worst > *max_rel_err) *max_rel_err = worst;
}

static void plc_fit_interval(long double a, long double s0, long double s1, double *coef, double *max_rel_err) {
    fit_interval([a](long double s) { return plc_G_ld(a, s); }, s0, s1, coef, max_rel_err);
}

// NFW force shape F(s) = (ln(1+s) - s/(1+s)) / s^3 in long double; below s = 0.1 the alternating series
// sum_{k>=2} (-1)^k (k-1)/k s^(k-3) (the two closed-form terms cancel to O(s^2)).
static long double nfw_F_ld(long double s) {
    if (s < 0.1L) {
        long double sum = 0.0L, pw = 1.0L / s;  // s^(k-3) for k = 2
        for (int k = 2; k < 80; ++k) {
            sum += ((k & 1) ? -1.0L : 1.0L) * (long double)(k - 1) / (long double)k * pw;
            pw *= s;
        }
        return sum;
    }
    return (log1pl(s) - s / (1.0L + s)) / (s * s * s);
}

// The universal NFW force table on the current device (built and uploaded on first use, immutable afterwards), or
// nullptr if it cannot be built to 1e-14 (then the kernels keep the closed form).
static const double *nfw_table(double *max_rel_err_out = nullptr) {
    struct Entry { int device; double *dev_ptr; double max_rel_err; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    for (const auto &e : cache)
        if (e.device == dev) {
            if (max_rel_err_out) *max_rel_err_out = e.max_rel_err;
            return e.dev_ptr;
        }
    static std::vector<double> host;  // the fit does not depend on the device: done once per process
    static double worst = 0.0;
    if (host.empty()) {
        host.resize((size_t)NFW_NINT * (PLC_DEG + 1));
        for (int j = 0; j < NFW_NINT; ++j) {
            const int e = NFW_E_LO + j / PLC_SUB, sub = j % PLC_SUB;
            const long double base = ldexpl(1.0L, e);
            fit_interval(nfw_F_ld, base * (1.0L + sub / (long double)PLC_SUB),
                         base * (1.0L + (sub + 1) / (long double)PLC_SUB), host.data() + (size_t)j * (PLC_DEG + 1), &worst);
        }
    }
    double *d = nullptr;
    if (worst < 1e-14 && cudaMalloc(&d, host.size() * sizeof(double)) == cudaSuccess) {
        if (cudaMemcpy(d, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(d);
            d = nullptr;
        }
    }
    cache.push_back({dev, d, worst});
    if (max_rel_err_out) *max_rel_err_out = worst;
    return d;
}

struct PlcTableEntry { int device; double a; double *dev_ptr; double max_rel_err; };

// Returns the device table for exponent a on the current device (building and uploading it on first use), or
// nullptr if it cannot be built to 1e-14 (then the kernels use the series).
static const double *plc_table_for(double a, double *max_rel_err_out = nullptr) {
    static std::mutex mu;
    static std::vector<PlcTableEntry> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    for (const auto &e : cache)
        if (e.device == dev && e.a == a) {
            if (max_rel_err_out) *max_rel_err_out = e.max_rel_err;
            return e.dev_ptr;
        }
    std::vector<double> host((size_t)PLC_NINT * (PLC_DEG + 1));
    double worst = 0.0;
    for (int j = 0; j < PLC_NINT; ++j) {
        const int e = PLC_E_LO + j / PLC_SUB, sub = j % PLC_SUB;
        const long double base = ldexpl(1.0L, e);
        plc_fit_interval(a, base * (1.0L + sub / (long double)PLC_SUB), base * (1.0L + (sub + 1) / (long double)PLC_SUB),
                         host.data() + (size_t)j * (PLC_DEG + 1), &worst);
    }
    double *d = nullptr;
    if (worst < 1e-14 && cudaMalloc(&d, host.size() * sizeof(double)) == cudaSuccess) {
        if (cudaMemcpy(d, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(d);
            d = nullptr;
        }
    }
    cache.push_back({dev, a, d, worst});
    if (max_rel_err_out) *max_rel_err_out = worst;
    return d;
}

// ---- combined spherical force table of one composite (SPH_* in gx_potential.cuh) --------------------------------
// S(u) = sum_i Phi_i'(r)/r at r = sqrt(u) over the composite's spherical components, in long double from the same fp64
// parameters (G m as the kernels form it) the closed forms use.
struct SphComp { int kind; double GM, p1, p2; };  // Hernquist (GM, c) / NFW (GM, r_s) / PowerLawCutoff (GM, r_c, a)
static long double sph_S_ld(const std::vector<SphComp> &cs, long double u) {
    const long double r = sqrtl(u);
    long double sum = 0.0L;
    for (const SphComp &c : cs) {
        if (c.kind == GX_KIND_HERNQUIST) {
            const long double w = r + (long double)c.p1;
            sum += (long double)c.GM / (r * w * w);
        } else if (c.kind == GX_KIND_NFW) {
            const long double rs = c.p1;
            sum += (long double)c.GM / (rs * rs * rs) * nfw_F_ld(r / rs);
        } else {  // PowerLawCutoff: P(a, s^2) == 1 to long-double precision beyond s^2 = 64 (a <= 3/2)
            const long double rc = c.p1, s = r / rc;
            const long double g = (s * s < 64.0L) ? plc_G_ld(c.p2, s) : 1.0L / (s * s * s);
            sum += (long double)c.GM / (rc * rc * rc) * g;
        }
    }
    return sum;
}

// First octave (of u = r^2) of the table of these components: the 22 octaves end at (8 x the largest scale radius)^2
// rounded up to a power of two; components without a scale (Kepler = Hernquist with c = 0) leave the default.
static int sph_e_lo(const std::vector<SphComp> &cs) {
    double smax = 0.0;
    for (const SphComp &c : cs)
        if (c.p1 > smax && std::isfinite(c.p1)) smax = c.p1;
    if (!(smax > 0.0)) return SPH_E_LO;
    const int e_hi_r = (int)ceil(log2(8.0 * smax));  // r up to 2^e_hi_r
    int e_lo = 2 * e_hi_r - SPH_OCTAVES;
    if (e_lo < -900) e_lo = -900;
    if (e_lo > 900) e_lo = 900;
    return e_lo;
}

// Fit of the table (host only; also used by gx_spherical_force_table): coef[SPHW_NINT][SPHW_ROW],
// returns the worst relative error of the fp64 Horner evaluation against the long-double function on a 41-point grid
// per interval.
template <int SUB_BITS = SPHW_SUB_BITS, int ROW = SPHW_ROW>
static double sph_table_fit(const std::vector<SphComp> &cs, double *coef) {
    double worst = 0.0;
    const int e_lo = sph_e_lo(cs);
    const int nint = SPH_OCTAVES << SUB_BITS;
    for (int j = 0; j < nint; ++j) {
        const int e = e_lo + (j >> SUB_BITS), sub = j & ((1 << SUB_BITS) - 1);
        const long double base = ldexpl(1.0L, e), nsub = (long double)(1 << SUB_BITS);
        // (the narrow intervals of the wide format are checked on a coarser grid: 4x as many of them)
        fit_interval<ROW, (SUB_BITS >= 7 ? 10 : 40)>([&cs](long double u) { return sph_S_ld(cs, u); }, base * (1.0L + sub / nsub),
                          base * (1.0L + (sub + 1) / nsub), coef + (size_t)j * ROW, &worst);
    }
    return worst;
}

// The rows as the kernels hold them (GX_SPH_ARG_DIFF): coefficient k of the interval [2^e (1 + sub/2^B), ...) times
// 2^(-k (e - B - 1)), i.e. monomials in u - centre instead of in t = (u - centre) / half-width.  Exact scalings.
template <int SUB_BITS, int ROW>
static void sph_rows_for_device(const std::vector<SphComp> &cs, std::vector<double> &rows) {
#if GX_SPH_ARG_DIFF
    const int e_lo = sph_e_lo(cs);
    const int nint = SPH_OCTAVES << SUB_BITS;
    for (int j = 0; j < nint; ++j) {
        const int sh = e_lo + (j >> SUB_BITS) - SUB_BITS - 1;
        for (int k = 1; k < ROW; ++k) rows[(size_t)j * ROW + k] = ldexp(rows[(size_t)j * ROW + k], -k * sh);
    }
#else
    (void)cs; (void)rows;
#endif
}

static bool sph_same(const std::vector<SphComp> &a, const std::vector<SphComp> &b) {  // (field by field: the struct has padding)
    if (a.size() != b.size()) return false;
    for (size_t k = 0; k < a.size(); ++k)
        if (a[k].kind != b[k].kind || a[k].GM != b[k].GM || a[k].p1 != b[k].p1 || a[k].p2 != b[k].p2) return false;
    return true;
}
constexpr size_t SPH_CACHE_MAX = 4096;  // x 132 KB = 540 MB of device memory at most (a parameter scan: 15-25 ms of host time per new set)
// The device table of this set of spherical components on the current device: fitted (12-25 ms of host time) and
// uploaded on FIRST use of a potential by an integrator, immutable afterwards and kept for the life of the process -- at
// most SPH_CACHE_MAX distinct (device, parameter set) entries; beyond that, if the fit misses 1e-14, or when the
// caller's stream is being captured and the table is not there yet: nullptr (the caller then runs the composite
// through the runtime-count kernels).
static const double *sph_table_for(const std::vector<SphComp> &cs, double *max_rel_err_out = nullptr, bool may_upload = true) {
    struct Entry { int device; std::vector<SphComp> cs; double *dev_ptr; double max_rel_err; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    static std::vector<std::vector<double>> fits;
    int dev = 0;
    if (cs.empty() || cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    const std::vector<double> *fitted = nullptr;
    double fitted_err = 0.0;
    for (size_t k = 0; k < cache.size(); ++k) {
        const Entry &e = cache[k];
        if (sph_same(e.cs, cs)) {
            if (e.device == dev) {
                if (max_rel_err_out) *max_rel_err_out = e.max_rel_err;
                return e.dev_ptr;
            }
            fitted = &fits[k];
            fitted_err = e.max_rel_err;
        }
    }
    if (cache.size() >= SPH_CACHE_MAX || !may_upload) return nullptr;
    std::vector<double> host;
    double worst = 0.0;
    if (fitted) {
        host = *fitted;  // (fitted for another device of this process)
        worst = fitted_err;
    } else {
        host.resize((size_t)SPHW_NINT * SPHW_ROW);
        worst = sph_table_fit<SPHW_SUB_BITS, SPHW_ROW>(cs, host.data());
    }
    double *d = nullptr;
    if (worst < 1e-14) {
        std::vector<double> img = host;
        sph_rows_for_device<SPHW_SUB_BITS, SPHW_ROW>(cs, img);
        if (cudaMalloc(&d, img.size() * sizeof(double)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(d);
            return nullptr;
        }
    }
    cache.push_back({dev, cs, d, worst});
    fits.push_back(std::move(host));
    if (max_rel_err_out) *max_rel_err_out = worst;
    return d;
}

}  // namespace gx
