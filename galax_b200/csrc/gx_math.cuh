// gx_math.cuh -- fp64 building blocks for the sm_100a kernels.
//
// The B200 FP64 pipe issues one DFMA-class warp instruction every 2 cycles per SM sub-partition;
// IEEE division / sqrt / log1p from libdevice cost 20-60 such instructions plus slow-path branches.
// The integrators only need the acceleration to ~1e-15 relative (its rounding error enters the state
// scaled by h*|a|/|p| ~ 1e-3), so the hot path uses MUFU-seeded reciprocal / rsqrt with one cubically
// convergent correction (max error ~1 ulp, no branches) and a lean log1p.  Everything here is plain
// IEEE arithmetic on normal inputs: no --use_fast_math, no flush-to-zero beyond the MUFU seed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gx_tables.h"

namespace gx {

// ---- MUFU seeds (relative error ~2^-20; only the upper 32 bits of the operand are examined)
__device__ __forceinline__ double rcp_seed(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// 1/x for normal x: seed e0 ~ 2^-20, one cubic step -> e0^3 ~ 2^-60, final rounding ~0.5-1 ulp.
__device__ __forceinline__ double rcp_fast(double x) {
    double y = rcp_seed(x);
    double e = fma(-x, y, 1.0);
    double t = fma(e, e, e);
    return fma(y, t, y);
}

// x^(-1/2) for normal x > 0: y1 = y0 (1 + h/2 + 3h^2/8), h = 1 - x y0^2.
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y = rsqrt_seed(x);
    double t = x * y;
    double h = fma(-t, y, 1.0);
    double u = y * h;
    return fma(u, fma(0.375, h, 0.5), y);
}

// sqrt(x) with one Newton correction on top of x*rsqrt_fast(x) (~0.5 ulp); rs receives rsqrt_fast(x).
__device__ __forceinline__ double sqrt_rs(double x, double &rs) {
    rs = rsqrt_fast(x);
    double g = x * rs;
    double r = fma(-g, g, x);
    return fma(r, 0.5 * rs, g);
}

// ---- log1p(s), s >= 0, given u = fl(1+s) and inv_u ~ 1/u (any accuracy >= the MUFU seed's).
// The rounding error of the sum, c = s - (u - 1), is exact for 0 <= s < 2^52 (u - 1 is exact because 1 is a
// multiple of ulp(u); the second difference is exact by Sterbenz) and is folded back in: log1p = log u + c/u.
// log u by the classic argument reduction u = 2^k m, m in [sqrt(1/2), sqrt(2)), f = m-1, w = f/(2+f),
// log m = 2 atanh(w) with the degree-7 minimax polynomial in w^2 of fdlibm's e_log.c (public domain,
// |error| < 2^-58.45).
__device__ __forceinline__ double log1p_pos(double s, double u, double inv_u) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    const double c = s - (u - 1.0);
    int hi = __double2hiint(u), lo = __double2loint(u);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    // m in [sqrt(1/2), sqrt(2)): if the mantissa is >= sqrt(2), halve m and bump k
    int i = (hi + 0x95f64) & 0x100000;
    k += (i >> 20);
    double m = __hiloint2double(hi | (i ^ 0x3ff00000), lo);
    double f = m - 1.0;
    double w = f * rcp_fast(2.0 + f);
    double z = w * w;
    double z2 = z * z;
    // even / odd split of the polynomial (two shorter dependency chains)
    double t1 = z2 * fma(z2, fma(z2, Lg6, Lg4), Lg2);
    double t2 = z * fma(z2, fma(z2, fma(z2, Lg7, Lg5), Lg3), Lg1);
    double hf = 0.5 * f;
    double hfsq = hf * f;
    double dk = (double)k;
    // log m = f - hfsq + w (hfsq + R);  log1p = k ln2 + log m + c/u
    double lo_part = fma(w, hfsq + (t1 + t2), fma(dk, ln2_lo, c * inv_u));
    return fma(dk, ln2_hi, f - (hfsq - lo_part));
}

// ---- log1p(s) for s >= 2^-6, table driven: u = fl(1+s) = 2^k m, j = top 8 mantissa bits of m,
//   ln u = k ln2 + logc_j + log1p(r),  r = m inv_c_j - 1 (one FMA, |r| <= 2^-9),  log1p(r) = r + r^2 p(r), deg p = 4
// (truncation r^7/7 < 2^-65).  {inv_c_j, logc_j = -ln(inv_c_j)} come from gx_tables.h (256 x 16 B, L1-resident).
// 12 FP64 instructions instead of the 33 of the atanh form above; error <= ~1.5 ulp of the result.  Not for s
// near 0: logc_j + log1p(r) cancels there (the NFW shape uses its series below s = 2^-4 anyway).
// The constants that are not 32-bit-immediate doubles sit in the constant bank (GX_LOG_CONST_BANK=1): as literals the
// compiler re-materialises each with two UMOV/IMAD.MOV per use, and on this pipe mix every instruction costs an
// issue slot (FP64 two): -7 instructions per NFW evaluation.
#ifndef GX_LOG_CONST_BANK
#define GX_LOG_CONST_BANK 1
#endif
#if GX_LOG_CONST_BANK
__constant__ double LOG_C[4] = {6.93147180559945286227e-01, -1.0 / 6.0, 0.2, 1.0 / 3.0};
#endif
__device__ __forceinline__ double log1p_tab(double s, double u, double inv_u) {
#if GX_LOG_CONST_BANK
    const double LN2 = LOG_C[0], C6 = LOG_C[1], C5 = LOG_C[2], C3 = LOG_C[3];
#else
    const double LN2 = 6.93147180559945286227e-01, C6 = -1.0 / 6.0, C5 = 0.2, C3 = 1.0 / 3.0;
#endif
    const double c = s - (u - 1.0);
    const int hi = __double2hiint(u);
    const int k = (hi >> 20) - 1023;
    const int j = (hi >> (20 - LOG_TAB_BITS)) & ((1 << LOG_TAB_BITS) - 1);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(u));
    const double2 t = __ldg(&LOG_TAB[j]);
    const double r = fma(m, t.x, -1.0);
    double p = fma(r, C6, C5);
    p = fma(p, r, -0.25);
    p = fma(p, r, C3);
    p = fma(p, r, -0.5);
    const double lo = fma(r * r, p, fma(c, inv_u, r));  // log1p(r) + c/u
    return fma((double)k, LN2, t.y) + lo;
}

__device__ __forceinline__ double log1p_pos(double s) {
    double u = 1.0 + s;
    return (s < 0.015625) ? log1p_pos(s, u, rcp_seed(u)) : log1p_tab(s, u, rcp_seed(u));
}

// ln(1+s) - s/(1+s), the NFW enclosed-mass shape.  Below s = 2^-4 the two terms cancel to O(s^2) and the
// alternating series sum_{k>=2} (-1)^k (k-1)/k s^k is used instead (17 terms: 0.0625^16 < 6e-20).
__device__ __forceinline__ double nfw_menc_shape(double s, double &inv_u) {
    if (__double2hiint(s) < 0x3fb00000) {  // s < 2^-4 for s >= 0, decided on the integer pipe (no FP64 DSETP)
        inv_u = rcp_fast(1.0 + s);
        double p = 17.0 / 18.0;  // k = 18
#pragma unroll
        for (int k = 17; k >= 2; --k) p = fma(-p, s, (double)(k - 1) / (double)k);
        return p * s * s;
    }
    const double u = 1.0 + s;
    inv_u = rcp_fast(u);
    return fma(-s, inv_u, log1p_tab(s, u, inv_u));
}
__device__ __forceinline__ double nfw_menc_shape(double s) {
    double inv_u;
    return nfw_menc_shape(s, inv_u);
}

// ---- regularised lower incomplete gamma P(a, x), a > 0, x >= 0 (Bovy bulge: a = 0.6).
//   P(a,x) = x^a e^-x / Gamma(a) * sum_{n>=0} x^n / (a (a+1) ... (a+n))
// The series has only positive terms (no cancellation) and is used for every x below the point where
// Q = 1 - P drops under 2^-54 (then P rounds to 1); the reciprocals 1/(a+n) are tabulated on the host, so a term
// costs one multiply and one add.  On exit *dP = x^(a-1) e^-x / Gamma(a) = dP/dx (needed by the Hessian).
constexpr int PLC_NT = 128;
struct GammaTab {
    double a, lgam, xcut;     // lgam = lgamma(a); P(a, x >= xcut) == 1 in fp64
    double inv[PLC_NT];       // inv[n] = 1 / (a + n)
};

__device__ __noinline__ double gammainc_P(const GammaTab &g, double x, double *dP) {
    if (!(x > 0.0)) {
        if (dP) *dP = (g.a == 1.0) ? 1.0 : ((g.a > 1.0) ? 0.0 : __longlong_as_double(0x7ff0000000000000LL));
        return 0.0;
    }
    const double pref = exp(fma(g.a, log(x), -x) - g.lgam);  // x^a e^-x / Gamma(a)
    if (dP) *dP = pref / x;
    if (x >= g.xcut) return 1.0;
    double del = g.inv[0], sum = del;
    for (int n = 1; n < PLC_NT; ++n) {
        del *= x * g.inv[n];
        sum += del;
        if (del < sum * 2e-17) break;
    }
    return fmin(sum * pref, 1.0);
}

}  // namespace gx
