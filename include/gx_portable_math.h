/* gx_portable_math.h -- bit-reproducible log, exp, log1p for fp64.
 *
 * The fdlibm algorithms (Sun Microsystems' freely distributable libm: e_log.c, e_exp.c, s_log1p.c), written with
 * nothing but IEEE +, -, *, / on doubles and integer operations on their bit patterns.  Compiled WITHOUT
 * floating-point contraction (gcc -ffp-contract=off, nvcc -fmad=false) the same source gives the same bits on an
 * x86-64 host and on an sm_100a device, which libm's and libdevice's own log/exp/log1p do not (each is accurate to
 * < 1 ulp, but they round differently).
 *
 * Users: the reference-order ("strict") integrator kernels of libgalax_b200.so (galax_b200/csrc/gx_strict.cu), whose
 * results must be reproducible bit for bit on a CPU, and the CPU restatement in oracle/galax_oracle.c that checks
 * them.  Accuracy (tests/test_portable_math.py, against 50-digit mpmath): < 1 ulp each.
 *
 * The NFW potential of the reference needs log1p (potential/_src/builtin/nfw/base.py:326-338), its PowerLawCutoff
 * potential the incomplete gamma function, i.e. exp and log (potential/_src/builtin/powerlawcutoff.py:88-117).
 */
#ifndef GX_PORTABLE_MATH_H
#define GX_PORTABLE_MATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GX_PM_FN __host__ __device__ static inline
#else
#define GX_PM_FN static inline
#endif

GX_PM_FN int64_t gx_pm_bits(double x) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    int64_t b;
    memcpy(&b, &x, sizeof b);
    return b;
#endif
}
GX_PM_FN double gx_pm_from_bits(int64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double x;
    memcpy(&x, &b, sizeof x);
    return x;
#endif
}
GX_PM_FN int32_t gx_pm_hi(double x) { return (int32_t)(gx_pm_bits(x) >> 32); }
GX_PM_FN uint32_t gx_pm_lo(double x) { return (uint32_t)(gx_pm_bits(x) & 0xffffffffLL); }
GX_PM_FN double gx_pm_with_hi(double x, int32_t hi) {
    return gx_pm_from_bits((int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint64_t)gx_pm_lo(x)));
}

#define GX_PM_LN2_HI 6.93147180369123816490e-01 /* 0x3fe62e42 0xfee00000 */
#define GX_PM_LN2_LO 1.90821492927058770002e-10 /* 0x3dea39ef 0x35793c76 */
#define GX_PM_LG1 6.666666666666735130e-01
#define GX_PM_LG2 3.999999999940941908e-01
#define GX_PM_LG3 2.857142874366239149e-01
#define GX_PM_LG4 2.222219843214978396e-01
#define GX_PM_LG5 1.818357216161805012e-01
#define GX_PM_LG6 1.531383769920937332e-01
#define GX_PM_LG7 1.479819860511658591e-01

/* natural logarithm, x > 0 finite (x <= 0, inf, nan: the IEEE special values) */
GX_PM_FN double gx_pm_log(double x) {
    int32_t hx = gx_pm_hi(x);
    const uint32_t lx = gx_pm_lo(x);
    int32_t k = 0;
    if (hx < 0x00100000) { /* x < 2^-1022 */
        if (((hx & 0x7fffffff) | (int32_t)lx) == 0) return -1.0 / (x * x) /* log(+-0) = -inf */;
        if (hx < 0) return (x - x) / (x - x); /* log(-#) = nan */
        k -= 54;
        x *= 1.80143985094819840000e+16; /* 2^54: scale up a subnormal */
        hx = gx_pm_hi(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int32_t i = (hx + 0x95f64) & 0x100000;
    x = gx_pm_with_hi(x, hx | (i ^ 0x3ff00000)); /* normalise x or x/2 */
    k += (i >> 20);
    const double f = x - 1.0;
    const double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) { /* |f| < 2^-20 */
        if (f == 0.0) return (k == 0) ? 0.0 : dk * GX_PM_LN2_HI + dk * GX_PM_LN2_LO;
        const double R = f * f * (0.5 - 0.33333333333333333 * f);
        return (k == 0) ? f - R : dk * GX_PM_LN2_HI - ((R - dk * GX_PM_LN2_LO) - f);
    }
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (GX_PM_LG2 + w * (GX_PM_LG4 + w * GX_PM_LG6));
    const double t2 = z * (GX_PM_LG1 + w * (GX_PM_LG3 + w * (GX_PM_LG5 + w * GX_PM_LG7)));
    const double R = t2 + t1;
    if (((hx - 0x6147a) | (0x6b851 - hx)) > 0) {
        const double hfsq = 0.5 * f * f;
        return (k == 0) ? f - (hfsq - s * (hfsq + R)) : dk * GX_PM_LN2_HI - ((hfsq - (s * (hfsq + R) + dk * GX_PM_LN2_LO)) - f);
    }
    return (k == 0) ? f - s * (f - R) : dk * GX_PM_LN2_HI - ((s * (f - R) - dk * GX_PM_LN2_LO) - f);
}

/* log(1 + x), x > -1 */
GX_PM_FN double gx_pm_log1p(double x) {
    const int32_t hx = gx_pm_hi(x);
    const int32_t ax = hx & 0x7fffffff;
    int32_t k = 1, hu = 0;
    double f = 0.0, c = 0.0;
    if (hx < 0x3FDA827A) { /* x < 0.41422 */
        if (ax >= 0x3ff00000) { /* x <= -1 */
            if (x == -1.0) return -1.0 / ((x + 1.0) * (x + 1.0)); /* -inf */
            return (x - x) / (x - x);
        }
        if (ax < 0x3e200000) { /* |x| < 2^-29 */
            if (ax < 0x3c900000) return x; /* |x| < 2^-54 */
            return x - x * x * 0.5;
        }
        if (hx > 0 || hx <= (int32_t)0xbfd2bec3) { /* -0.2929 < x < 0.41422 */
            k = 0;
            f = x;
            hu = 1;
        }
    }
    if (hx >= 0x7ff00000) return x + x;
    if (k != 0) {
        double u;
        if (hx < 0x43400000) {
            u = 1.0 + x;
            hu = gx_pm_hi(u);
            k = (hu >> 20) - 1023;
            c = (k > 0) ? 1.0 - (u - x) : x - (u - 1.0); /* the rounding error of 1 + x */
            c /= u;
        } else {
            u = x;
            hu = gx_pm_hi(u);
            k = (hu >> 20) - 1023;
            c = 0.0;
        }
        hu &= 0x000fffff;
        if (hu < 0x6a09e) {
            u = gx_pm_with_hi(u, hu | 0x3ff00000); /* normalise u */
        } else {
            k += 1;
            u = gx_pm_with_hi(u, hu | 0x3fe00000); /* normalise u/2 */
            hu = (0x00100000 - hu) >> 2;
        }
        f = u - 1.0;
    }
    const double dk = (double)k;
    const double hfsq = 0.5 * f * f;
    if (hu == 0) { /* |f| < 2^-20 */
        if (f == 0.0) {
            if (k == 0) return 0.0;
            c += dk * GX_PM_LN2_LO;
            return dk * GX_PM_LN2_HI + c;
        }
        const double R = hfsq * (1.0 - 0.66666666666666666 * f);
        return (k == 0) ? f - R : dk * GX_PM_LN2_HI - ((R - (dk * GX_PM_LN2_LO + c)) - f);
    }
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double R = z * (GX_PM_LG1 + z * (GX_PM_LG2 + z * (GX_PM_LG3 + z * (GX_PM_LG4 + z * (GX_PM_LG5 + z * (GX_PM_LG6 + z * GX_PM_LG7))))));
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * GX_PM_LN2_HI - ((hfsq - (s * (hfsq + R) + (dk * GX_PM_LN2_LO + c))) - f);
}

/* exponential */
GX_PM_FN double gx_pm_exp(double x) {
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    const double INVLN2 = 1.44269504088896338700e+00;
    int32_t hx = gx_pm_hi(x);
    const int32_t xsb = (int32_t)(((uint32_t)hx >> 31) & 1u);
    hx &= 0x7fffffff;
    double hi = 0.0, lo = 0.0;
    int32_t k = 0;
    if (hx >= 0x40862E42) { /* |x| >= 709.78... */
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | (int32_t)gx_pm_lo(x)) != 0) return x + x; /* nan */
            return (xsb == 0) ? x : 0.0;                                      /* exp(+-inf) */
        }
        if (x > 7.09782712893383973096e+02) return 1.0e300 * 1.0e300; /* overflow */
        if (x < -7.45133219101941108420e+02) return 0.0;               /* underflow */
    }
    if (hx > 0x3fd62e42) {     /* |x| > 0.5 ln2 */
        if (hx < 0x3FF0A2B2) { /* and |x| < 1.5 ln2 */
            hi = (xsb == 0) ? x - GX_PM_LN2_HI : x + GX_PM_LN2_HI;
            lo = (xsb == 0) ? GX_PM_LN2_LO : -GX_PM_LN2_LO;
            k = 1 - xsb - xsb;
        } else {
            k = (int32_t)(INVLN2 * x + ((xsb == 0) ? 0.5 : -0.5));
            const double t = (double)k;
            hi = x - t * GX_PM_LN2_HI;
            lo = t * GX_PM_LN2_LO;
        }
        x = hi - lo;
    } else if (hx < 0x3e300000) { /* |x| < 2^-28 */
        return 1.0 + x;
    }
    const double t = x * x;
    const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) return gx_pm_with_hi(y, gx_pm_hi(y) + (int32_t)((uint32_t)k << 20));
    y = gx_pm_with_hi(y, gx_pm_hi(y) + (int32_t)((uint32_t)(k + 1000) << 20));
    return y * 9.33263618503218878990e-302; /* 2^-1000 */
}

/* x^y for x >= 0 as exp(y log x): a few ulp for |y log x| < 10, which is all its users need (the step-size controller's
 * factor = safety * (1/err)^(1/order), the initial-step heuristic's (0.01/d)^(1/(order+1))).  pow(inf, y > 0) = inf,
 * pow(0, y > 0) = 0, nan propagates. */
GX_PM_FN double gx_pm_pow(double x, double y) { return gx_pm_exp(y * gx_pm_log(x)); }

#endif /* GX_PORTABLE_MATH_H */
