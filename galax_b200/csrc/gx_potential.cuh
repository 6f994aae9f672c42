// gx_potential.cuh -- analytic composite potentials on the device.
//
// Replaces, for the supported component classes, the reference's autodiff path
//   AbstractPotential._gradient/_hessian  = jax.grad / jax.hessian of _potential
//     (/root/reference/src/galax/potential/_src/base.py:170-179,230-239)
//   AbstractCompositePotential._potential/_gradient/_hessian = sum over components
//     (/root/reference/src/galax/potential/_src/base_multi.py:39-82)
// with hand-derived closed forms of
//   MiyamotoNagai   builtin/miyamotonagai.py:73-78
//   Hernquist       builtin/hernquist.py:79-82     (r = safe_sqrt(x^2+y^2+z^2 + tiny), utils.py:44-125)
//   NFW             builtin/nfw/base.py:326-338
//   PowerLawCutoff  builtin/powerlawcutoff.py:88-117
// MN3 disks arrive already expanded into three MiyamotoNagai components (builtin/mn3.py:90-119 is
// host-side parameter algebra).
//
// Components are regrouped by kind on the host so that all spherical components share one r, 1/r and
// the flattened ones share R^2, z^2; the parameters live in the kernel-parameter constant bank
// (__grid_constant__), i.e. every access is a uniform constant-cache read.
#pragma once
#include "gx_math.cuh"

namespace gx {

constexpr int MAX_MN = 6;
constexpr int MAX_HERN = 4;
constexpr int MAX_NFW = 2;
constexpr int MAX_PLC = 1;
constexpr int MAX_LOG = 2;
constexpr int MAX_ISO = 2;
constexpr int MAX_SATOH = 2;
constexpr int MAX_RAD = 4;
constexpr int MAX_HARM = 1;
constexpr int MAX_HENON = 1;
constexpr double TINY = 2.2250738585072014e-308;

struct DevMN { double GM, a, b2, ab2; };             // ab2 = a*b^2 (Hessian)
struct DevHern { double GM, c; };
struct DevNFW { double GM, rs, inv_rs, GM_rs3, GM_inv_rs, pad_; };  // GM_rs3 = GM / rs^3 (force table); {inv_rs, GM_rs3} one 16-byte load
// ga: a = 3/2 - alpha/2; ga2: a - 1/2; tail = Gamma(a2)/(rc Gamma(a)).
// tab: optional device table of G(s) = P(a, s^2) / s^3, s = r/r_c, as degree-(PLC_DEG) polynomials on 2^PLC_SUB_BITS
// intervals per octave of s in [2^PLC_E_LO, 2^PLC_E_HI) (built on the host in long double, see plc_table.h);
// GM_rc3 = GM/rc^3.  Degree 9 on 32 intervals per octave is at the rounding floor for the value (2.1e-16; 5e-15 for the
// derivative the Hessian needs) -- the first table (degree 13 on 8 intervals) spent 4 more FMAs and, with 8-byte loads,
// 9 more load instructions per evaluation.  A row is PLC_DEG + 1 = 10 doubles (80 B, 16-byte aligned): five 16-byte loads.
constexpr int PLC_DEG = 9, PLC_E_LO = -11, PLC_E_HI = 3, PLC_SUB_BITS = 5, PLC_SUB = 1 << PLC_SUB_BITS;
constexpr int PLC_NINT = (PLC_E_HI - PLC_E_LO) * PLC_SUB;
constexpr double PLC_S_ONE = 8.0;  // = 2^PLC_E_HI: s^2 = 64 > xcut (<= 41 for every 0 < a <= 3/2), so P == 1
struct DevPLC { double GM, inv_rc, tail, GM_rc3; const double *tab; GammaTab ga, ga2; };

#ifndef GX_HERN_PAIR
#define GX_HERN_PAIR 1
#endif
// SMEM: the table was staged into shared memory (plc_stage below), rows of PLC_STRIDE doubles.
constexpr int PLC_STRIDE = PLC_DEG + 1;
static_assert(PLC_DEG % 2 == 1 && PLC_STRIDE % 2 == 0, "rows are read as pairs of doubles (16-byte loads)");
__device__ __forceinline__ double2 lds_v2f64(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
// Piecewise-polynomial table lookup: value and (optionally) derivative with respect to s of a function tabulated on
// 2^PLC_SUB_BITS intervals per octave of s in [2^E_LO, 2^E_LO + NINT / 2^PLC_SUB_BITS octaves); returns false when s is
// outside the tabulated range.  Shared by the PowerLawCutoff and the NFW force tables.
template <bool SMEM, int E_LO, int NINT, int B = PLC_SUB_BITS, bool ESTRIN = false>
__device__ __forceinline__ bool poly_table_eval(const double *tab, double s, double &G, double *dG,
                                                unsigned smem_base = 0) {
    if (!SMEM && tab == nullptr) return false;
    const int hi = __double2hiint(s);
    // interval index from the bits of s > 0: (hi >> (20 - B)) = exponent field * 2^B + top B mantissa bits
    const unsigned j = (unsigned)(hi >> (20 - B)) - (unsigned)((1023 + E_LO) << B);
    if (j >= (unsigned)NINT) return false;  // s outside the table (also NaN / negative)
    // interval [2^e (1 + sub/2^B), 2^e (1 + (sub+1)/2^B)), t in [-1, 1):  t = 2^(B+1) m - (2^(B+1) + 2 sub + 1) with
    // m = s / 2^e in [1, 2).  Both operands come straight from the bits of s (exact).
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(s));
    constexpr int TOP = ((1 << B) - 1) << (20 - B), HALF = 1 << (19 - B), EXPC = (1023 + B + 1) << 20;
    const double cB = __hiloint2double((hi & TOP) | HALF | EXPC, 0);  // 2^(B+1) (1 + sub/2^B + 2^-(B+1))
    const double t = fma(m, (double)(2 << B), -cB);
    double v, d = 0.0;
    if (SMEM && ESTRIN) {
        // Estrin's scheme: 12 FP64 instructions, 4 deep (Horner: 9, 9 deep) -- for the latency-bound Dopri kernels,
        // whose right-hand side ends on this polynomial.  Value only.
        static_assert(PLC_DEG == 9, "Estrin form written for degree 9");
        const unsigned a0 = (smem_base ? smem_base : (unsigned)__cvta_generic_to_shared(tab)) + j * (unsigned)(PLC_STRIDE * 8);
        const double2 c01 = lds_v2f64(a0), c23 = lds_v2f64(a0 + 16), c45 = lds_v2f64(a0 + 32), c67 = lds_v2f64(a0 + 48),
                      c89 = lds_v2f64(a0 + 64);
        const double t2 = t * t, t4 = t2 * t2, t8 = t4 * t4;
        const double p01 = fma(c01.y, t, c01.x), p23 = fma(c23.y, t, c23.x), p45 = fma(c45.y, t, c45.x),
                     p67 = fma(c67.y, t, c67.x), p89 = fma(c89.y, t, c89.x);
        const double q0 = fma(p23, t2, p01), q1 = fma(p67, t2, p45);
        v = fma(p89, t8, fma(q1, t4, q0));
    } else if (SMEM) {
        // 32-bit shared-window address; the fixed-step kernels pass the table's base (plc_smem_base) so that the window
        // base (S2R SR_CgaCtaId + LEA) is not re-derived in every step
        const unsigned a0 = (smem_base ? smem_base : (unsigned)__cvta_generic_to_shared(tab)) + j * (unsigned)(PLC_STRIDE * 8);
        double2 c = lds_v2f64(a0 + (PLC_DEG - 1) * 8);
        v = c.y;
        if (dG) d = v;
        v = fma(v, t, c.x);
#pragma unroll
        for (int k = PLC_DEG - 3; k >= 0; k -= 2) {
            c = lds_v2f64(a0 + k * 8);
            if (dG) d = fma(d, t, v);
            v = fma(v, t, c.y);
            if (dG) d = fma(d, t, v);
            v = fma(v, t, c.x);
        }
    } else {
        const double2 *p = reinterpret_cast<const double2 *>(tab + (long long)j * (PLC_DEG + 1));
        double2 c = __ldg(p + (PLC_DEG - 1) / 2);
        v = c.y;
        if (dG) d = v;
        v = fma(v, t, c.x);
#pragma unroll
        for (int k = (PLC_DEG - 3) / 2; k >= 0; --k) {
            c = __ldg(p + k);
            if (dG) d = fma(d, t, v);
            v = fma(v, t, c.y);
            if (dG) d = fma(d, t, v);
            v = fma(v, t, c.x);
        }
    }
    G = v;
    if (dG) {
        const int e = ((hi >> 20) & 0x7ff) - 1023;
        *dG = d * __hiloint2double((1023 - e + B + 1) << 20, 0);  // dt/ds = 2^(B+1-e)
    }
    return true;
}
template <bool SMEM>
__device__ __forceinline__ bool plc_table_eval_at(const double *tab, double s, double &G, double *dG,
                                                  unsigned smem_base = 0) {
    if (tab == nullptr) return false;
    return poly_table_eval<SMEM, PLC_E_LO, PLC_NINT>(tab, s, G, dG, smem_base);
}
// NFW force table: F(s) = (ln(1+s) - s/(1+s)) / s^3, so that Phi'/r = G m M(s) / r^3 = (G m / r_s^3) F(s), s = r/r_s.
// One universal function (no parameter): 12 octaves [2^-7, 2^5) x 32 intervals x degree 9 = 30 KB, staged in shared
// memory by the fixed-step and Dopri kernels of the static models without a PowerLawCutoff component.  Replaces, per
// evaluation, the reciprocal of 1+s, the table logarithm and the small-s series switch (21 FP64 + 12 other
// instructions) by 12 FP64 + five 16-byte loads + 5 integer instructions, and is accurate to 2e-16 everywhere (the
// closed form loses digits to cancellation around s ~ 2^-4).  Outside the range the closed form is used.
constexpr int NFW_E_LO = -7, NFW_E_HI = 5;
constexpr int NFW_NINT = (NFW_E_HI - NFW_E_LO) * PLC_SUB;
#ifndef GX_NFW_TABLE
#define GX_NFW_TABLE 1
#endif
// Combined spherical force table of one composite: S(u) = sum over its spherical components of Phi_i'(r)/r as a
// function of u = r^2 (Hernquist GM/(r (r+c)^2), NFW (GM/r_s^3) F(r/r_s), PowerLawCutoff (GM/r_c^3) G(r/r_c)), fitted
// per potential on the host in long double (plc_table.h: sph_table_for).  Indexed by the bits of r^2, it removes from a
// right-hand side everything spherical: the rsqrt of r^2, the Hernquist reciprocals, the NFW / PowerLawCutoff lookups
// in s = r/r_s (MilkyWayPotential2022: 77 -> 60 FP64 instructions, 6 -> 4 MUFU chains), and the lookup no longer waits
// for sqrt(r^2).
// Layout: 2^7 = 128 intervals per octave of u, degree 5, rows of 6 doubles = 48 bytes = THREE 16-byte loads per lookup
// (an odd number of chunks: consecutive rows cover the eight bank groups without padding or swizzle), 22 octaves: 2816
// rows = 132 KB, held in dynamic shared memory by ONE CTA per SM.  The kernels that use it are bound by the
// shared-memory port (every lane reads another row), so a row is as short as the accuracy allows -- history: degree 9
// on 16 intervals, 80-byte rows, five loads (40 KB); degree 7 on 32 intervals, 64-byte swizzled rows, four loads
// (44 KB, +17..21 % on the port-bound fixed-step kernels); this one (+13..29 % again).
// The 22 octaves are placed per potential (sph_e_lo() in plc_table.h): up to 8 x the largest scale radius of the
// spherical components, rounded up to a power of two -- [2^-8, 2^14) kpc^2, i.e. 62 pc to 128 kpc, for the three
// Milky-Way models in galactic units; the library does not know the unit system, the scale radii do.
// Outside the range: the closed forms (spherical_fallback, out of line).
constexpr int SPH_E_LO = -8, SPH_OCTAVES = 22;
constexpr int SPHW_SUB_BITS = 7, SPHW_DEG = 5, SPHW_ROW = SPHW_DEG + 1;
#ifndef GX_SPH_ARG_DIFF
#define GX_SPH_ARG_DIFF 1
#endif
constexpr int SPHW_NINT = SPH_OCTAVES << SPHW_SUB_BITS;
constexpr int SPHW_BYTES = SPHW_NINT * SPHW_ROW * 8;
static_assert(SPHW_ROW == 6 && SPHW_BYTES % 16 == 0, "rows are three 16-byte chunks");
#ifndef GX_SPH_TABLE
#define GX_SPH_TABLE 1
#endif
__device__ __forceinline__ bool plc_table_eval(const DevPLC &c, double s, double &G, double *dG) {
    return plc_table_eval_at<false>(c.tab, s, G, dG);
}

// (LMJ09)LogarithmicPotential, builtin/logarithmic.py:45-108: Phi = vc^2/2 ln(rs^2 + x^T M x), M symmetric with
// M = R(phi)^T diag(1/q1^2, 1/q2^2, 1/q3^2) R(phi) (rotation about z): c11, c12, c22, c33.
struct DevLog { double vc2, rs2, c11, c12, c22, c33; };
// IsochronePotential, builtin/isochrone.py:80-90: Phi = -GM / (b + sqrt(r^2 + b^2)).
struct DevIso { double GM, b, b2; };
// SatohPotential, builtin/satoh.py:63-70: Phi = -GM / sqrt(R^2 + z^2 + a (a + 2 sqrt(z^2 + b^2))) -- a
// Miyamoto-Nagai form whose D^2 is smaller by b^2.
struct DevSatoh { double GM, a, b2, ab2; };

// Profiles in an ellipsoidal radius m^2 = x^2 + y^2/q1^2 + z^2/q2^2 (i1 = 1/q1^2, i2 = 1/q2^2; spherical: 1, 1):
//   RAD_HERNQUIST  TriaxialHernquistPotential, hernquist.py:160-176   K = GM,                    a = r_s
//   RAD_JAFFE      JaffePotential,             jaffe.py:50-60         K = GM,                    a = r_s
//   RAD_BURKERT    BurkertPotential,           burkert.py:197-227     K = GM/(3 ln2 - pi/2),     a = r_s, b = 1/r_s
//   RAD_STONE      StoneOstriker15Potential,   stoneostriker15.py     K = 2GM/(pi (r_h - r_c)),  a = r_c, b = r_h
enum { RAD_HERNQUIST = 0, RAD_JAFFE = 1, RAD_BURKERT = 2, RAD_STONE = 3 };
struct DevRad { int profile, pad_; double K, a, b, i1, i2; };
// HarmonicOscillatorPotential (example.py:75-85), HenonHeilesPotential (example.py:159-176): polynomials.
struct DevHarm { double w2x, w2y, w2z; };
struct DevHenon { double k, it2; };

// Time-dependent composites (LinearParameter, params/core.py:25-110): raw parameters and their rates; the integrators
// evaluate every right-hand side with the parameters of its own time (gradient_td below).  n == 0: static potential.
constexpr int TD_MAX = 14, TD_NP = 4;
struct DevTD {
    int n, pad_;
    double G;
    int kind[TD_MAX];
    double p[TD_MAX][TD_NP], dp[TD_MAX][TD_NP];
};

struct alignas(16) DevPot {
    int n_mn, n_hern, n_nfw, n_plc, n_log, n_iso, n_satoh, n_rad, n_harm, n_henon, pad0_, pad1_;  // 48 B: the arrays below
    // start 16-byte aligned, so {GM, a} of a disk or {GM, c} of a sphere are one 16-byte constant load
    DevMN mn[MAX_MN];
    DevHern hern[MAX_HERN];
    DevNFW nfw[MAX_NFW];
    DevPLC plc[MAX_PLC];
    DevLog lg[MAX_LOG];
    DevIso iso[MAX_ISO];
    DevSatoh satoh[MAX_SATOH];
    DevRad rad[MAX_RAD];
    DevHarm harm[MAX_HARM];
    DevHenon henon[MAX_HENON];
    const double *nfw_tab;  // universal NFW force table (nfw_table(), plc_table.h) or nullptr
    const double *sph_wide; // this composite's spherical force table S(r^2) (sph_table_for(), plc_table.h) or nullptr
    unsigned sph_j0w;       // ... and where it starts: (1023 + e_lo) << SPHW_SUB_BITS, the table covers u in [2^e_lo, 2^(e_lo + 22))
    unsigned pad_sph_;
    DevTD td;
};

// Burkert: B(s) = 2 ln(1+s) + ln(1+s^2) - 2 atan(s) = M(<r) C/m.  Below s = 0.3 the three terms cancel to
// O(s^3): B' = 4 s^2 / ((1+s)(1+s^2)) = 4 s^2 (1-s)/(1-s^4) integrates to 4 s^3 sum_n s^4n (1/(4n+3) - s/(4n+4)).
__device__ __forceinline__ double burkert_B(double s) {
    if (s < 0.3) {
        const double s4 = (s * s) * (s * s);
        double e = 0.0;
#pragma unroll
        for (int n = 8; n >= 0; --n) e = fma(e, s4, 1.0 / (double)(4 * n + 3) - s / (double)(4 * n + 4));
        return 4.0 * (s * s * s) * e;
    }
    return 2.0 * log1p(s) + log1p(s * s) - 2.0 * atan(s);
}
// Stone-Ostriker: T(m) = r_h atan(m/r_h) - r_c atan(m/r_c) = M(<r) pi (r_h - r_c)/(2 M); series below 0.3 r_c.
__device__ __forceinline__ double stone_T(double m, double rc, double rh) {
    if (m < 0.3 * rc) {
        const double uc = (m / rc) * (m / rc), uh = (m / rh) * (m / rh);
        double acc = 0.0, pc = 1.0, ph = 1.0;
#pragma unroll
        for (int k = 1; k <= 16; ++k) {
            pc *= uc; ph *= uh;
            const double term = (ph - pc) / (double)(2 * k + 1);
            acc += (k & 1) ? -term : term;
        }
        return m * acc;
    }
    return rh * atan2(m, rh) - rc * atan2(m, rc);
}

// f = F'(m)/m and (optionally) d2 = F''(m), phi = F(m) of a DevRad profile; m2 = m^2 (> 0).
__device__ __forceinline__ void rad_profile(const DevRad &c, double m2, double &f, double *d2, double *phi) {
    const double minv = rsqrt_fast(m2), m = m2 * minv;
    switch (c.profile) {
    case RAD_HERNQUIST: {
        const double iu = rcp_fast(m + c.a);
        const double d1 = c.K * iu * iu;
        f = d1 * minv;
        if (d2) *d2 = -2.0 * d1 * iu;
        if (phi) *phi = -c.K * iu;
        break;
    }
    case RAD_JAFFE: {
        const double iu = rcp_fast(m + c.a);
        const double d1 = c.K * minv * iu;  // GM / (m (m + a))
        f = d1 * minv;
        if (d2) *d2 = -d1 * (2.0 * m + c.a) * (minv * iu);
        if (phi) *phi = -c.K / c.a * log(1.0 + c.a * minv);
        break;
    }
    case RAD_BURKERT: {
        const double s = m * c.b;
        const double d1 = c.K * burkert_B(s) * (minv * minv);
        f = d1 * minv;
        if (d2) *d2 = fma(c.K * 4.0 * c.b * c.b * c.b, 1.0 / ((1.0 + s) * fma(s, s, 1.0)), -2.0 * f);
        if (phi) {
            const double si = 1.0 / s;
            *phi = -c.K * c.b * (3.14159265358979323846 - 2.0 * (1.0 + si) * atan(s) + 2.0 * (1.0 + si) * log1p(s) -
                                 (1.0 - si) * log1p(s * s));
        }
        break;
    }
    default: {  // RAD_STONE
        const double rc = c.a, rh = c.b;
        const double T = stone_T(m, rc, rh);
        f = c.K * T * (minv * minv) * minv;
        if (d2) *d2 = fma(c.K * (rh * rh - rc * rc), 1.0 / ((m2 + rh * rh) * (m2 + rc * rc)), -2.0 * f);
        if (phi) *phi = -c.K * (T * minv + 0.5 * log((m2 + rh * rh) / (m2 + rc * rc)));
        break;
    }
    }
}

// Static component counts let the compiler unroll and schedule the whole evaluation as one block of
// straight-line code; Runtime (-1) is the generic fallback for arbitrary composites of the four kinds.
template <int NMN, int NH, int NNFW, int NPLC, bool MN_SHARED_B = false, bool BASIC = false, bool BASIC_TAB = false>
struct Counts {
    static constexpr bool is_static = (NMN >= 0);
    // BASIC_TAB (with BASIC): the composite's spherical components -- however many Hernquist / NFW / PowerLawCutoff
    // terms -- are ONE lookup in its combined force table S(r^2), as for the three named models; only the
    // Miyamoto-Nagai terms are looped over at run time.  The host picks it when the table exists (P.sph_wide).
    static constexpr bool basic_tab = BASIC && BASIC_TAB;
    // BASIC (runtime counts only): a composite of the four basic kinds with constant parameters -- the loops over the
    // further kinds and the time-dependent branch are compiled out (the full runtime kernel is 85 KB of SASS against a
    // 32 KB instruction cache).
    static constexpr bool basic_only = BASIC;
    // all Miyamoto-Nagai terms have the same b (the three disks of an MN3 model, mn3.py:121-130): sqrt(z^2 + b^2)
    // is evaluated once.  Same bits as evaluating it per term; the host checks the equality before dispatching.
    static constexpr bool mn_shared_b = MN_SHARED_B;
    static constexpr int kMN = is_static ? NMN : MAX_MN, kH = is_static ? NH : MAX_HERN,
                         kNFW = is_static ? NNFW : MAX_NFW, kPLC = is_static ? NPLC : MAX_PLC;
    // the three specialised Milky-Way models contain none of the further kinds; the runtime path loops over them
    static constexpr int kLOG = (is_static || BASIC) ? 0 : MAX_LOG, kISO = (is_static || BASIC) ? 0 : MAX_ISO,
                         kSAT = (is_static || BASIC) ? 0 : MAX_SATOH;
    static constexpr int kRAD = (is_static || BASIC) ? 0 : MAX_RAD, kHARM = (is_static || BASIC) ? 0 : MAX_HARM,
                         kHENON = (is_static || BASIC) ? 0 : MAX_HENON;
    template <class PT> __device__ __forceinline__ static int mn(const PT &P) { return is_static ? NMN : P.n_mn; }
    template <class PT> __device__ __forceinline__ static int hern(const PT &P) { return is_static ? NH : P.n_hern; }
    template <class PT> __device__ __forceinline__ static int nfw(const PT &P) { return is_static ? NNFW : P.n_nfw; }
    template <class PT> __device__ __forceinline__ static int plc(const PT &P) { return is_static ? NPLC : P.n_plc; }
};
using CountsRuntime = Counts<-1, -1, -1, -1>;
using CountsBasic = Counts<-1, -1, -1, -1, false, true>;  // runtime counts of MN / Hernquist / NFW / PowerLawCutoff only
using CountsBasicTab = Counts<-1, -1, -1, -1, false, true, true>;  // ... with the spherical ones in the combined table
using CountsMW = Counts<1, 2, 1, 0>;      // MilkyWayPotential:      MN disk, NFW halo, 2 Hernquist
using CountsMW2022 = Counts<3, 2, 1, 0, true>;  // MilkyWayPotential2022:  MN3 disk (one b), NFW halo, 2 Hernquist
using CountsBovy = Counts<1, 0, 1, 1>;    // BovyMWPotential2014:    MN disk, PLC bulge, NFW halo

// Shared-memory copy of the (single) PowerLawCutoff table of a static model (35 KB).  Every lane reads a different
// interval, so from global memory each coefficient load of a step costs up to 32 L1 wavefronts per warp (the Bovy
// fixed-step kernel was load-bound, not FP64-bound); from shared memory it is a few.
template <class C>
__device__ __forceinline__ double *plc_smem() {
    __shared__ __align__(16) double t[PLC_NINT * PLC_STRIDE];
    return t;
}
// Shared-window address of the staged table, made opaque so that it stays in a register across the step loop.
template <class C>
__device__ __forceinline__ unsigned plc_smem_base() {
    unsigned b = 0;
    if constexpr (C::is_static && C::kPLC > 0) {
        b = (unsigned)__cvta_generic_to_shared(plc_smem<C>());
        asm volatile("" : "+r"(b));
    }
    return b;
}
// Call once per CTA, by all threads, before the first gradient<C, true>().
template <class C>
__device__ __forceinline__ void plc_stage(const DevPot &P) {
    if constexpr (C::is_static && C::kPLC > 0) {
        double *t = plc_smem<C>();
        const double *src = P.plc[0].tab;
        static_assert(PLC_STRIDE == PLC_DEG + 1, "rows are copied as they are");
        const double2 *src2 = reinterpret_cast<const double2 *>(src);
        double2 *t2 = reinterpret_cast<double2 *>(t);
        for (int idx = threadIdx.x; idx < PLC_NINT * PLC_STRIDE / 2; idx += blockDim.x) t2[idx] = __ldg(src2 + idx);
        __syncthreads();
    }
}

// The same for the NFW force table (static models with one NFW halo and no PowerLawCutoff table: both do not fit
// beside 5-6 resident CTAs).
template <class C>
__host__ __device__ constexpr bool nfw_tab_ok() { return GX_NFW_TABLE && C::is_static && C::kNFW == 1 && C::kPLC == 0; }
// The lookup moves 80 B per evaluation from randomly scattered rows through the SM's single shared-memory port (bank
// conflicts included: ~60 port cycles per warp evaluation).  The Dopri kernels (3 warps per scheduler) gain 2.5-4.5 %;
// of the fixed-step kernels MilkyWayPotential2022's heavier step (three disks) gains 6 %, but MilkyWayPotential's step
// (166 issue cycles per warp-step on each of four schedulers) saturates that port and runs 14 % SLOWER with the table.
// So the fixed-step kernels use it from GX_NFW_TABLE_FIXED_MIN_MN Miyamoto-Nagai terms up.
#ifndef GX_NFW_TABLE_FIXED_MIN_MN
#define GX_NFW_TABLE_FIXED_MIN_MN 3
#endif
template <class C>
__host__ __device__ constexpr bool nfw_tab_fixed_ok() { return nfw_tab_ok<C>() && C::kMN >= GX_NFW_TABLE_FIXED_MIN_MN; }
template <class C>
__device__ __forceinline__ double *nfw_smem() {
    __shared__ __align__(16) double t[NFW_NINT * PLC_STRIDE];
    return t;
}
// Call once per CTA, by all threads; returns the shared-window address of the table.
template <class C, bool ON>
__device__ __forceinline__ unsigned nfw_stage(const DevPot &P) {
    unsigned b = 0;
    if constexpr (ON) {  // (the host only picks a static model's integrator kernels when P.nfw_tab exists)
        double *t = nfw_smem<C>();
        const double2 *src2 = reinterpret_cast<const double2 *>(P.nfw_tab);
        double2 *t2 = reinterpret_cast<double2 *>(t);
        for (int idx = threadIdx.x; idx < NFW_NINT * PLC_STRIDE / 2; idx += blockDim.x) t2[idx] = __ldg(src2 + idx);
        __syncthreads();
        b = (unsigned)__cvta_generic_to_shared(t);
        asm volatile("" : "+r"(b));
    }
    return b;
}

// The combined spherical table S(r^2): which kernels use it.
//   mode 0: not used; 3: Horner (fixed-step kernels); 4: Estrin (latency-bound Dopri kernels).
template <class C>
__host__ __device__ constexpr bool sph_tab_ok() {
    return GX_SPH_TABLE && ((C::is_static && (C::kH + C::kNFW + C::kPLC > 0)) || C::basic_tab);
}
// Fixed step: the lookup moves 80 B per lane from scattered rows through the SM's one shared-memory port (~60 port
// cycles per warp with the bank conflicts); MilkyWayPotential's single-disk step is short enough to saturate it (the
// NFW table alone cost it 14 %), so the fixed-step kernels take the table from GX_SPH_TABLE_FIXED_MIN_MN disks up or
// when the model has a PowerLawCutoff bulge (whose own table went through the same port anyway).
#ifndef GX_SPH_TABLE_FIXED_MIN_MN
#define GX_SPH_TABLE_FIXED_MIN_MN 3
#endif
template <class C>
__host__ __device__ constexpr bool sph_tab_fixed_ok() {
    return sph_tab_ok<C>() && (C::basic_tab || C::kMN >= GX_SPH_TABLE_FIXED_MIN_MN || C::kPLC > 0);
}
// Fixed-step kernels of a static model that does NOT take the table on every step (MilkyWayPotential): out of every
// GX_SPH_MIX_PERIOD steps, GX_SPH_MIX_TABLE use the table and the rest the closed forms (0: never mix), evenly spread.
// Measured with the wide table: 5/8 2.71e11, 3/4 3.02e11, 7/8 3.03e11, 13/16 3.10e11, 15/16 3.04e11 particle-steps/s.
#ifndef GX_SPH_MIX_PERIOD
#define GX_SPH_MIX_PERIOD 16
#endif
#ifndef GX_SPH_MIX_TABLE
#define GX_SPH_MIX_TABLE 13
#endif
static_assert(GX_SPH_MIX_PERIOD == 0 || (GX_SPH_MIX_PERIOD & (GX_SPH_MIX_PERIOD - 1)) == 0, "a power of two (32-bit step counter)");
template <class C>
__host__ __device__ constexpr bool sph_mix_ok() {
    return GX_SPH_MIX_PERIOD > 0 && sph_tab_ok<C>() && !sph_tab_fixed_ok<C>();
}
__device__ __forceinline__ bool sph_mix_table_step(unsigned long long step) {
    // evenly spread (Bresenham): step * TABLE mod PERIOD < TABLE
    return (unsigned)((step * (unsigned)GX_SPH_MIX_TABLE) % (unsigned)(GX_SPH_MIX_PERIOD > 0 ? GX_SPH_MIX_PERIOD : 1)) <
           (unsigned)GX_SPH_MIX_TABLE;
}
// S(u) from the table at the start of the CTA's dynamic shared memory (base = its shared-window address); false outside
// the tabulated range.
template <bool ESTRIN = false>
__device__ __forceinline__ bool sph_wide_eval(double u, double &S, unsigned base, unsigned j0w) {
    const int hi = __double2hiint(u);
    constexpr int B = SPHW_SUB_BITS;
    const unsigned j = (unsigned)(hi >> (20 - B)) - j0w;
    if (j >= (unsigned)SPHW_NINT) return false;
#if GX_SPH_ARG_DIFF
    // polynomial argument = u - (centre of the interval), an exact difference; the device rows are the fitted monomial
    // coefficients in t in [-1, 1) rescaled by powers of two (plc_table.h::sph_rows_for_device), so every intermediate
    // of the evaluation is the power-of-two multiple of what the t form gives: the same bits, three integer
    // instructions fewer per lookup
    constexpr int BELOW = (1 << (20 - B)) - 1, HALF = 1 << (19 - B);
    const double t = u - __hiloint2double((hi & ~BELOW) | HALF, 0);
#else
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(u));
    constexpr int TOP = ((1 << B) - 1) << (20 - B), HALF = 1 << (19 - B), EXPC = (1023 + B + 1) << 20;
    const double cB = __hiloint2double((hi & TOP) | HALF | EXPC, 0);
    const double t = fma(m, (double)(2 << B), -cB);
#endif
    const unsigned a0 = base + j * (unsigned)(SPHW_ROW * 8);
    const double2 c01 = lds_v2f64(a0), c23 = lds_v2f64(a0 + 16u), c45 = lds_v2f64(a0 + 32u);
    if (ESTRIN) {  // 6 FP64, 3 deep (the latency-bound Dopri kernels, whose right-hand side ends on this polynomial)
        const double t2 = t * t;
        const double p01 = fma(c01.y, t, c01.x), p23 = fma(c23.y, t, c23.x), p45 = fma(c45.y, t, c45.x);
        S = fma(fma(p45, t2, p23), t2, p01);
    } else {       // Horner: 5 FP64 (the fixed-step kernels)
        double v = fma(c45.y, t, c45.x);
        v = fma(v, t, c23.y); v = fma(v, t, c23.x);
        v = fma(v, t, c01.y); v = fma(v, t, c01.x);
        S = v;
    }
    return true;
}
__device__ __forceinline__ double *dyn_smem() {
    extern __shared__ __align__(16) double gx_save_buf[];  // (the one dynamic shared array of the library: table, then save staging)
    return gx_save_buf;
}
// The same lookup in two halves (SPH = 5): the three loads, issued as soon as r^2 is known with the row index clamped
// into the table (always a valid address), and the Horner evaluation after the disk terms, whose arithmetic hides the
// loads' latency.  Same operations, same bits as sph_wide_eval<false>.  For launches that are bound by the dependent
// chain of one step (few warps per scheduler): C1's 10^4 particles 1.96 -> 1.68 ms (MilkyWayPotential), 2.01 -> 1.50
// (MW2022), 1.77 -> 1.31 (Bovy); full machines lose 2 % to the twelve registers the row occupies meanwhile, so the host
// picks this instantiation for CTAs of up to 512 threads only (the crossover, measured) -- the bits do not depend on it.
struct SphRow { double2 c01, c23, c45; double t; bool ok; };
__device__ __forceinline__ SphRow sph_wide_fetch(double u, unsigned base, unsigned j0w) {
    SphRow r;
    const int hi = __double2hiint(u);
    constexpr int B = SPHW_SUB_BITS;
    const unsigned j = (unsigned)(hi >> (20 - B)) - j0w;
    r.ok = j < (unsigned)SPHW_NINT;
    const unsigned jc = min(j, (unsigned)(SPHW_NINT - 1));
    constexpr int BELOW = (1 << (20 - B)) - 1, HALF = 1 << (19 - B);
    r.t = u - __hiloint2double((hi & ~BELOW) | HALF, 0);
    const unsigned a0 = base + jc * (unsigned)(SPHW_ROW * 8);
    r.c01 = lds_v2f64(a0); r.c23 = lds_v2f64(a0 + 16u); r.c45 = lds_v2f64(a0 + 32u);
    return r;
}
template <bool ESTRIN = false>
__device__ __forceinline__ double sph_wide_poly(const SphRow &r) {
    if (ESTRIN) {
        const double t2 = r.t * r.t;
        const double p01 = fma(r.c01.y, r.t, r.c01.x), p23 = fma(r.c23.y, r.t, r.c23.x), p45 = fma(r.c45.y, r.t, r.c45.x);
        return fma(fma(p45, t2, p23), t2, p01);
    }
    double v = fma(r.c45.y, r.t, r.c45.x);
    v = fma(v, r.t, r.c23.y); v = fma(v, r.t, r.c23.x);
    v = fma(v, r.t, r.c01.y); v = fma(v, r.t, r.c01.x);
    return v;
}
// Call once per CTA, by all threads: copies the wide table to the start of dynamic shared memory.
__device__ __forceinline__ unsigned sph_wide_stage(const DevPot &P) {
    double2 *t2 = reinterpret_cast<double2 *>(dyn_smem());
    const double2 *src2 = reinterpret_cast<const double2 *>(P.sph_wide);
    for (int idx = threadIdx.x; idx < SPHW_BYTES / 16; idx += blockDim.x) t2[idx] = __ldg(src2 + idx);
    __syncthreads();
    unsigned b = (unsigned)__cvta_generic_to_shared(dyn_smem());
    asm volatile("" : "+r"(b));
    return b;
}

// ---------------------------------------------------------------------------------------------
// gradient (hot path of the integrators): g = grad Phi(q).  ~1-2 ulp per term, branch-free except the
// small-s NFW series and the incomplete-gamma routine.
// The flattened + spherical part is returned as two scalars: grad = (fh x, fh y, fv z); the kinds that are neither
// (triaxial logarithmic / ellipsoidal profiles / polynomials) are added to (ex, ey, ez) by gradient<C>() below.
// Terms of the form GM_i w_i / r (Hernquist, NFW) are summed before the common factor 1/r is applied.
// Sum over the spherical components of Phi'(r)/r at r^2 = r2 (> 0), closed forms / per-kind tables.
template <class C, bool PLC_SMEM, bool NFW_TAB>
__device__ __forceinline__ double spherical_factor(const DevPot &P, double r2, unsigned plc_base, unsigned nfw_base) {
    double fs = 0.0, fr = 0.0;  // fs: terms Phi'/r as they are; fr: terms still to be divided by r
    const double rinv = rsqrt_fast(r2);
    const double r = r2 * rinv;
    const double rinv2 = rinv * rinv;  // (kept apart from the third factor: rinv^3 overflows at r -> 0)
#if GX_HERN_PAIR
    if constexpr (C::is_static && C::kH == 2) {
        // bulge + nucleus over one reciprocal: GM1/u1^2 + GM2/u2^2 = (GM1 u2^2 + GM2 u1^2) / (u1 u2)^2
        // (one MUFU seed + refinement instead of two: -1 FP64 and -3 other instructions per evaluation)
        const double u1 = r + P.hern[0].c, u2 = r + P.hern[1].c;
        const double w1 = u1 * u1, w2 = u2 * u2;
        const double num = fma(P.hern[1].GM, w1, P.hern[0].GM * w2);
        fr = num * rcp_fast(w1 * w2);
    } else
#endif
    {
#pragma unroll
        for (int i = 0; i < C::kH; ++i) {
            if (!C::is_static && i >= P.n_hern) break;
            const DevHern &c = P.hern[i];
            double u = r + c.c;
            double w = rcp_fast(u * u);  // Phi'/r = GM / ((r+c)^2 r)
            if (C::is_static && i == 0) fr = c.GM * w; else fr = fma(c.GM, w, fr);
        }
    }
#pragma unroll
    for (int i = 0; i < C::kISO; ++i) {
        if (i >= P.n_iso) break;
        const DevIso &c = P.iso[i];
        double ia = rsqrt_fast(r2 + c.b2);            // 1/a, a = sqrt(r^2 + b^2)
        double ibpa = rcp_fast(fma(r2 + c.b2, ia, c.b));  // 1/(b + a)
        fs = fma(c.GM * ia, ibpa * ibpa, fs);         // Phi'/r = GM / (a (b+a)^2)
    }
#pragma unroll
    for (int i = 0; i < C::kNFW; ++i) {
        if (!C::is_static && i >= P.n_nfw) break;
        const DevNFW &c = P.nfw[i];
        const double s = r * c.inv_rs;
        if constexpr (NFW_TAB) {
            double F;
            if (poly_table_eval<true, NFW_E_LO, NFW_NINT>(nullptr, s, F, nullptr, nfw_base)) {
                fs = fma(c.GM_rs3, F, fs);  // Phi'/r = (GM / rs^3) F(s)
                continue;
            }
        }
        double m = nfw_menc_shape(s);
        // Phi'/r = GM m(s) / r^3 = (GM m / r^2) / r
        if (C::is_static && C::kH == 0 && i == 0) fr = (c.GM * m) * rinv2; else fr = fma(c.GM * m, rinv2, fr);
    }
#pragma unroll
    for (int i = 0; i < C::kPLC; ++i) {
        if (!C::is_static && i >= P.n_plc) break;
        const DevPLC &c = P.plc[i];
        const double s = r * c.inv_rc;
        double Gs;
        if (s >= PLC_S_ONE) {
            fr = fma(c.GM, rinv2, fr);  // beyond the table P(a, s^2) == 1 in fp64: Kepler
        } else if (PLC_SMEM ? plc_table_eval_at<true>(plc_smem<C>(), s, Gs, nullptr, plc_base)
                            : plc_table_eval(c, s, Gs, nullptr)) {
            fs = fma(c.GM_rc3, Gs, fs);  // GM P(a, s^2) / r^3 = (GM / rc^3) G(s)
        } else {
            const double Pg = gammainc_P(c.ga, s * s, nullptr);
            fr = fma(c.GM * Pg, rinv2, fr);
        }
    }
    if (C::is_static && C::kPLC == 0 && !NFW_TAB) fs = fr * rinv; else fs = fma(fr, rinv, fs);
    return fs;
}
// the combined table's out-of-range path (r < 7.8 pc or r > 512 kpc): the closed forms, out of line
template <class C>
__device__ __noinline__ double spherical_fallback(const DevPot *P, double r2) {
    return spherical_factor<C, false, false>(*P, r2, 0u, 0u);
}

template <class C, bool PLC_SMEM = false, bool NFW_TAB = false, int SPH = 0>
__device__ __forceinline__ void gradient_factors(const DevPot &P, double x, double y, double z, double &fh, double &fv,
                                                 unsigned plc_base = 0, unsigned nfw_base = 0) {
    // (SPH != 0: nfw_base is the shared-window address of the combined spherical table)
    const double z2 = z * z;
    const double R2 = fma(y, y, fma(x, x, TINY));  // TINY (2^-1022) is below half an ulp of any normal x^2: same bits
    double fxy = 0.0, fz = 0.0, fs = 0.0;
    double zeta2 = 0.0, rz = 0.0;
    constexpr bool EARLY = GX_SPH_ARG_DIFF && (SPH == 5 || SPH == 6);  // (5: Horner like 3, 6: Estrin like 4, the row fetched before the disk terms)
    SphRow row;
    double r2e = 0.0;
    if constexpr (EARLY) {
        r2e = fma(z, z, R2);
        row = sph_wide_fetch(r2e, nfw_base, P.sph_j0w);
    }
#pragma unroll
    for (int i = 0; i < C::kMN; ++i) {
        if (!C::is_static && i >= P.n_mn) break;
        const DevMN &c = P.mn[i];
        if (!C::mn_shared_b || i == 0) {
            // (explicit FMAs / un-contractable adds below: a product feeding a sum is the compiler's to fuse or not, call
            //  site by call site, and the run-length kernel, its save step and the step-by-step kernel must agree bit for bit)
            zeta2 = fma(z, z, c.b2);
            rz = rsqrt_fast(zeta2);          // 1/zeta
        }
        double apz = fma(zeta2, rz, c.a);    // a + zeta
        double D2 = fma(apz, apz, R2);
        double rD = rsqrt_fast(D2);
        double f = (c.GM * rD) * (rD * rD);  // GM / D^3
        if constexpr (C::is_static && C::mn_shared_b) {
            // one zeta for all disks: sum_i f_i (a_i + zeta) first, the common 1/zeta once after the loop
            if (i == 0) { fxy = f; fz = f * apz; }
            else { fxy = __dadd_rn(fxy, f); fz = fma(f, apz, fz); }
        } else {
            const double w = apz * rz;           // (a+zeta)/zeta
            // (an explicit FMA: left to the compiler, the run-length and the step-by-step kernel contracted this sum
            //  differently for a two-disk runtime composite, and the two must give the same bits)
            if (C::is_static && i == 0) { fxy = f; fz = f * w; }
            else { fxy = __dadd_rn(fxy, f); fz = fma(f, w, fz); }
        }
    }
    if constexpr (C::is_static && C::mn_shared_b && C::kMN > 0) fz *= rz;
#pragma unroll
    for (int i = 0; i < C::kSAT; ++i) {
        if (i >= P.n_satoh) break;
        const DevSatoh &c = P.satoh[i];
        double zs2 = z2 + c.b2;
        double rzs = rsqrt_fast(zs2);
        double apz = fma(zs2, rzs, c.a);
        double rD = rsqrt_fast(fma(apz, apz, R2 - c.b2));
        double f = (c.GM * rD) * (rD * rD);
        fxy += f;
        fz = fma(f, apz * rzs, fz);
    }
    const bool any_sph = (C::is_static || SPH != 0) ? (SPH != 0 || C::kH + C::kNFW + C::kPLC > 0)
                                                    : (P.n_hern + P.n_nfw + P.n_plc + P.n_iso > 0);
    if (any_sph) {
        const double r2 = fma(z, z, R2);  // (R2 carries the TINY that keeps r > 0)
        if constexpr (EARLY) {
            fs = sph_wide_poly<SPH == 6>(row);
            if (!row.ok) fs = spherical_fallback<C>(&P, r2e);
        } else if constexpr (SPH != 0) {
            // SPH: 3 Horner (fixed-step kernels), 4 Estrin (Dopri kernels); 5 (Horner, early loads) is handled above
            if (!sph_wide_eval<SPH == 4>(r2, fs, nfw_base, P.sph_j0w)) fs = spherical_fallback<C>(&P, r2);
        } else {
            fs = spherical_factor<C, PLC_SMEM, NFW_TAB>(P, r2, plc_base, nfw_base);
        }
    }
    fh = __dadd_rn(fxy, fs);
    fv = __dadd_rn(fz, fs);
}

// the remaining kinds: e += grad of (triaxial logarithmic, ellipsoidal-radius profiles, polynomials)
template <class C>
__device__ __forceinline__ void gradient_extras(const DevPot &P, double x, double y, double z, double &gx_, double &gy_,
                                                double &gz_) {
#pragma unroll
    for (int i = 0; i < C::kLOG; ++i) {
        if (i >= P.n_log) break;
        const DevLog &c = P.lg[i];
        const double mx = fma(c.c11, x, c.c12 * y), my = fma(c.c12, x, c.c22 * y), mz = c.c33 * z;  // M x
        const double f = c.vc2 * rcp_fast(c.rs2 + fma(x, mx, fma(y, my, z * mz)));
        gx_ = fma(f, mx, gx_);
        gy_ = fma(f, my, gy_);
        gz_ = fma(f, mz, gz_);
    }
#pragma unroll 1
    for (int i = 0; i < C::kRAD; ++i) {
        if (i >= P.n_rad) break;
        const DevRad &c = P.rad[i];
        const double wy = y * c.i1, wz = z * c.i2;
        double f;
        rad_profile(c, fma(x, x, fma(y, wy, fma(z, wz, TINY))), f, nullptr, nullptr);
        gx_ = fma(f, x, gx_);
        gy_ = fma(f, wy, gy_);
        gz_ = fma(f, wz, gz_);
    }
    for (int i = 0; i < C::kHARM; ++i) {
        if (i >= P.n_harm) break;
        gx_ = fma(P.harm[i].w2x, x, gx_);
        gy_ = fma(P.harm[i].w2y, y, gy_);
        gz_ = fma(P.harm[i].w2z, z, gz_);
    }
    for (int i = 0; i < C::kHENON; ++i) {
        if (i >= P.n_henon) break;
        const double k = P.henon[i].k, it2 = P.henon[i].it2;
        gx_ = fma(fma(2.0 * k * x, y, x), it2, gx_);
        gy_ = fma(fma(k, x * x - y * y, y), it2, gy_);
    }
}

// Gradient of a time-dependent composite at time t: the closed forms of gradient_factors() evaluated from the raw
// parameters p_k(t) = p[k] + dp[k] t (kinds as in include/galax_b200.h; the host admits only these seven here).
__device__ __noinline__ void gradient_td(const DevTD &R, double t, double x, double y, double z, double &gx_, double &gy_,
                                         double &gz_) {
    const double z2 = z * z, R2 = fma(y, y, fma(x, x, TINY));
    const double r2 = R2 + z2, rinv = rsqrt_fast(r2), r = r2 * rinv, rinv2 = rinv * rinv;
    double fxy = 0.0, fz = 0.0, fs = 0.0, ex = 0.0, ey = 0.0, ez = 0.0;
    for (int i = 0; i < R.n; ++i) {
        const double p0 = fma(R.dp[i][0], t, R.p[i][0]), p1 = fma(R.dp[i][1], t, R.p[i][1]);
        const double p2 = fma(R.dp[i][2], t, R.p[i][2]), p3 = fma(R.dp[i][3], t, R.p[i][3]);
        const double GM = R.G * p0;
        switch (R.kind[i]) {
        case 0: {  // Miyamoto-Nagai (m, a, b)
            const double zeta2 = fma(p2, p2, z2 + TINY), rz = rsqrt_fast(zeta2), apz = fma(zeta2, rz, p1);  // (+tiny: b = 0, z = 0)
            const double rD = rsqrt_fast(fma(apz, apz, R2)), f = (GM * rD) * (rD * rD);
            fxy += f;
            fz = fma(f, apz * rz, fz);
            break;
        }
        case 1: {  // Hernquist (m, r_s)
            const double u = r + p1;
            fs = fma(GM * rinv, rcp_fast(u * u), fs);
            break;
        }
        case 2: {  // NFW (m, r_s)
            const double m = nfw_menc_shape(r * rcp_fast(p1));
            fs = fma((GM * m) * rinv, rinv2, fs);
            break;
        }
        case 5: {  // Isochrone (m, b)
            const double a2 = fma(p1, p1, r2), ia = rsqrt_fast(a2), ibpa = rcp_fast(fma(a2, ia, p1));
            fs = fma(GM * ia, ibpa * ibpa, fs);
            break;
        }
        case 6: {  // Satoh (m, a, b)
            const double b2 = p2 * p2, zs2 = (z2 + TINY) + b2, rzs = rsqrt_fast(zs2), apz = fma(zs2, rzs, p1);
            const double rD = rsqrt_fast(fma(apz, apz, R2 - b2)), f = (GM * rD) * (rD * rD);
            fxy += f;
            fz = fma(f, apz * rzs, fz);
            break;
        }
        case 7: {  // triaxial Hernquist (m, r_s, q1, q2)
            const double i1 = rcp_fast(p2 * p2), i2 = rcp_fast(p3 * p3), wy = y * i1, wz = z * i2;
            const double m2 = fma(x, x, fma(y, wy, fma(z, wz, TINY))), minv = rsqrt_fast(m2), u = fma(m2, minv, p1);
            const double f = GM * minv * rcp_fast(u * u);
            ex = fma(f, x, ex); ey = fma(f, wy, ey); ez = fma(f, wz, ez);
            break;
        }
        default: {  // 8: Jaffe (m, r_s): Phi'/r = GM / (r^2 (r + a))
            fs = fma(GM * rinv2, rcp_fast(r + p1), fs);
            break;
        }
        }
    }
    gx_ = fma(fxy + fs, x, ex);
    gy_ = fma(fxy + fs, y, ey);
    gz_ = fma(fz + fs, z, ez);
}

// gradient (hot path of the integrators): g = grad Phi(q).  ~1-2 ulp per term, branch-free except the
// small-s NFW series and the incomplete-gamma routine.
template <class C, bool PLC_SMEM = false, bool NFW_TAB = false, int SPH = 0>
__device__ __forceinline__ void gradient(const DevPot &P, double x, double y, double z, double &gx_, double &gy_,
                                         double &gz_, double t = 0.0, unsigned nfw_base = 0) {
    if (!C::is_static && !C::basic_only && P.td.n > 0) {  // time-dependent composite (runtime path only)
        gradient_td(P.td, t, x, y, z, gx_, gy_, gz_);
        return;
    }
    double fh, fv;
    gradient_factors<C, PLC_SMEM, NFW_TAB, SPH>(P, x, y, z, fh, fv, 0, nfw_base);
    gx_ = fh * x;
    gy_ = fh * y;
    gz_ = fv * z;
    gradient_extras<C>(P, x, y, z, gx_, gy_, gz_);
}

// ---------------------------------------------------------------------------------------------
// potential value and Hessian (bulk-evaluation kernel only; HBM-bound, so IEEE div/sqrt/log1p).
template <class C, class PT = DevPot>
__device__ __forceinline__ double potential_value(const PT &P, double x, double y, double z) {
    // the same MUFU-seeded primitives as the gradient (each within ~1 ulp); the PowerLawCutoff term keeps the series
    const double z2 = z * z, R2 = fma(y, y, x * x);
    double phi = 0.0;
    double zeta2 = 0.0, rz = 0.0;
    for (int i = 0; i < C::mn(P); ++i) {
        const DevMN &c = P.mn[i];
        if (!C::mn_shared_b || i == 0) {
            zeta2 = z2 + c.b2;
            rz = rsqrt_fast(zeta2);
        }
        const double apz = fma(zeta2, rz, c.a);
        phi = fma(-c.GM, rsqrt_fast(fma(apz, apz, R2)), phi);
    }
    const double r2 = (R2 + z2) + TINY;
    const double rinv = rsqrt_fast(r2);
    const double r = r2 * rinv;
    for (int i = 0; i < C::hern(P); ++i) phi = fma(-P.hern[i].GM, rcp_fast(r + P.hern[i].c), phi);
    for (int i = 0; i < C::nfw(P); ++i) {
        const double s = r * P.nfw[i].inv_rs;
        // ln(1+s)/s: s > 0 always (TINY); rinv * rs = 1/s without another reciprocal
        phi = fma(-P.nfw[i].GM_inv_rs * log1p_pos(s), rinv * P.nfw[i].rs, phi);
    }
    for (int i = 0; i < C::plc(P); ++i) {
        const DevPLC &c = P.plc[i];
        double s = r * c.inv_rc, s2 = s * s;
        double Pa = gammainc_P(c.ga, s2, nullptr);
        double Qa2 = 1.0 - gammainc_P(c.ga2, s2, nullptr);
        phi -= c.GM * (Pa / r + Qa2 * c.tail);
    }
    for (int i = 0; i < C::kSAT; ++i) {
        if (i >= P.n_satoh) break;
        const DevSatoh &c = P.satoh[i];
        phi -= c.GM / sqrt(R2 + z2 + c.a * (c.a + 2.0 * sqrt(z2 + c.b2)));
    }
    for (int i = 0; i < C::kISO; ++i) {
        if (i >= P.n_iso) break;
        phi -= P.iso[i].GM / (P.iso[i].b + sqrt(r * r + P.iso[i].b2));
    }
    for (int i = 0; i < C::kLOG; ++i) {
        if (i >= P.n_log) break;
        const DevLog &c = P.lg[i];
        const double q = fma(x, fma(c.c11, x, c.c12 * y), fma(y, fma(c.c12, x, c.c22 * y), c.c33 * z2));
        phi += 0.5 * c.vc2 * log(c.rs2 + q);
    }
    for (int i = 0; i < C::kRAD; ++i) {
        if (i >= P.n_rad) break;
        const DevRad &c = P.rad[i];
        double f, v;
        rad_profile(c, fma(x, x, fma(y, y * c.i1, fma(z, z * c.i2, TINY))), f, nullptr, &v);
        phi += v;
    }
    for (int i = 0; i < C::kHARM; ++i) {
        if (i >= P.n_harm) break;
        phi += 0.5 * fma(P.harm[i].w2x * x, x, fma(P.harm[i].w2y * y, y, P.harm[i].w2z * z2));
    }
    for (int i = 0; i < C::kHENON; ++i) {
        if (i >= P.n_henon) break;
        phi += (0.5 * R2 + P.henon[i].k * (x * x * y - y * y * y / 3.0)) * P.henon[i].it2;
    }
    return phi;
}

// ---------------------------------------------------------------------------------------------
// fused gradient + Hessian with the fast primitives (bulk kernel, C5 and the stream release).
// g = grad Phi; H[0..5] = (xx, xy, xz, yy, yz, zz).
template <class C, class PT = DevPot>
__device__ __forceinline__ void grad_hess(const PT &P, double x, double y, double z, double g[3], double H[6]) {
    const double z2 = z * z, R2 = fma(y, y, x * x);
    double gxy = 0.0, gz = 0.0;  // g = (gxy x, gxy y, gz) for the flattened part
    double zeta2 = 0.0, rz = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) H[k] = 0.0;
#pragma unroll
    for (int i = 0; i < C::kMN; ++i) {
        if (!C::is_static && i >= P.n_mn) break;
        const DevMN &c = P.mn[i];
        if (!C::mn_shared_b || i == 0) {
            zeta2 = z2 + c.b2;
            rz = rsqrt_fast(zeta2);
        }
        const double apz = fma(zeta2, rz, c.a);
        const double D2 = fma(apz, apz, R2);
        const double rD = rsqrt_fast(D2), rD2 = rD * rD;
        const double f3 = (c.GM * rD) * rD2;     // GM / D^3
        const double f5 = 3.0 * f3 * rD2;        // 3 GM / D^5
        const double w = apz * rz;               // (a + zeta) / zeta
        const double uz = z * w;
        const double duz = fma(c.ab2 * rz, rz * rz, 1.0);  // 1 + a b^2 / zeta^3
        gxy += f3;
        gz = fma(f3, w, gz);
        H[0] += fma(-f5 * x, x, f3);
        H[1] -= f5 * x * y;
        H[2] -= f5 * x * uz;
        H[3] += fma(-f5 * y, y, f3);
        H[4] -= f5 * y * uz;
        H[5] += fma(-f5 * uz, uz, f3 * duz);
    }
#pragma unroll
    for (int i = 0; i < C::kSAT; ++i) {
        if (i >= P.n_satoh) break;
        const DevSatoh &c = P.satoh[i];
        const double zeta2 = z2 + c.b2;
        const double rz = rsqrt_fast(zeta2);
        const double apz = fma(zeta2, rz, c.a);
        const double rD = rsqrt_fast(fma(apz, apz, R2 - c.b2)), rD2 = rD * rD;
        const double f3 = (c.GM * rD) * rD2;
        const double f5 = 3.0 * f3 * rD2;
        const double w = apz * rz;
        const double uz = z * w;
        const double duz = fma(c.ab2 * rz, rz * rz, 1.0);
        gxy += f3;
        gz = fma(f3, w, gz);
        H[0] += fma(-f5 * x, x, f3);
        H[1] -= f5 * x * y;
        H[2] -= f5 * x * uz;
        H[3] += fma(-f5 * y, y, f3);
        H[4] -= f5 * y * uz;
        H[5] += fma(-f5 * uz, uz, f3 * duz);
    }
    double d1r = 0.0, d2 = 0.0;  // sum over spherical components of Phi'/r and Phi''
    const bool any_sph = C::is_static ? (C::kH + C::kNFW + C::kPLC > 0)
                                      : (P.n_hern + P.n_nfw + P.n_plc + P.n_iso > 0);
    if (any_sph) {
        const double r2 = (R2 + z2) + TINY;
        const double rinv = rsqrt_fast(r2);
        const double r = r2 * rinv;
        const double rinv2 = rinv * rinv;
#pragma unroll
        for (int i = 0; i < C::kH; ++i) {
            if (!C::is_static && i >= P.n_hern) break;
            const double iu = rcp_fast(r + P.hern[i].c);
            const double d1 = P.hern[i].GM * iu * iu;  // GM / (r+c)^2
            d1r = fma(d1, rinv, d1r);
            d2 = fma(-2.0 * d1, iu, d2);
        }
#pragma unroll
        for (int i = 0; i < C::kNFW; ++i) {
            if (!C::is_static && i >= P.n_nfw) break;
            const DevNFW &c = P.nfw[i];
            double iu;
            const double m = nfw_menc_shape(r * c.inv_rs, iu);
            const double t = ((c.GM * m) * rinv) * rinv2;  // Phi'/r = GM m / r^3
            d1r += t;
            // Phi'' = GM s / (rs (1+s)^2 r^2) - 2 Phi'/r = (GM / rs^2) (1/r) / (1+s)^2 - 2 Phi'/r
            d2 += fma(c.GM_inv_rs * c.inv_rs * rinv, iu * iu, -2.0 * t);
        }
#pragma unroll
        for (int i = 0; i < C::kPLC; ++i) {
            if (!C::is_static && i >= P.n_plc) break;
            const DevPLC &c = P.plc[i];
            const double s = r * c.inv_rc;
            double Gs, dGs;
            if (s >= PLC_S_ONE) {
                const double t = (c.GM * rinv) * rinv2;  // P == 1, dP/dx == 0 in fp64
                d1r += t;
                d2 -= 2.0 * t;
            } else if (plc_table_eval(c, s, Gs, &dGs)) {
                // Phi'/r = K G(s), K = GM/rc^3;  Phi'' = d(r K G)/dr = K (G + s G')
                const double t = c.GM_rc3 * Gs;
                d1r += t;
                d2 += fma(c.GM_rc3 * s, dGs, t);
            } else {
                double dP;
                const double Pg = gammainc_P(c.ga, s * s, &dP);
                const double t = ((c.GM * Pg) * rinv) * rinv2;
                d1r += t;
                d2 += fma(c.GM * dP * 2.0 * c.inv_rc * c.inv_rc, rinv, -2.0 * t);
            }
        }
#pragma unroll
        for (int i = 0; i < C::kISO; ++i) {
            if (i >= P.n_iso) break;
            const DevIso &c = P.iso[i];
            const double ia = rsqrt_fast(r2 + c.b2);
            const double ibpa = rcp_fast(fma(r2 + c.b2, ia, c.b));
            const double t = (c.GM * ia) * (ibpa * ibpa);  // Phi'/r
            d1r += t;
            // Phi' = r t  =>  Phi'' = t - r^2 t (1/a^2 + 2/(a (b+a)))
            d2 += t - r2 * t * fma(2.0 * ia, ibpa, ia * ia);
        }
        // H += (Phi'/r) I + (Phi'' - Phi'/r) n n^T
        const double w = (d2 - d1r) * rinv2;
        H[0] += fma(w * x, x, d1r);
        H[1] = fma(w * x, y, H[1]);
        H[2] = fma(w * x, z, H[2]);
        H[3] += fma(w * y, y, d1r);
        H[4] = fma(w * y, z, H[4]);
        H[5] += fma(w * z, z, d1r);
    }
    g[0] = (gxy + d1r) * x;
    g[1] = (gxy + d1r) * y;
    g[2] = (gz + d1r) * z;
#pragma unroll
    for (int i = 0; i < C::kLOG; ++i) {
        if (i >= P.n_log) break;
        const DevLog &c = P.lg[i];
        const double mx = fma(c.c11, x, c.c12 * y), my = fma(c.c12, x, c.c22 * y), mz = c.c33 * z;
        const double iD = rcp_fast(c.rs2 + fma(x, mx, fma(y, my, z * mz)));
        const double f = c.vc2 * iD, f2 = 2.0 * f * iD;
        g[0] = fma(f, mx, g[0]); g[1] = fma(f, my, g[1]); g[2] = fma(f, mz, g[2]);
        // H = vc^2 (M / D - 2 (M x)(M x)^T / D^2)
        H[0] += fma(f, c.c11, -f2 * mx * mx);
        H[1] += fma(f, c.c12, -f2 * mx * my);
        H[2] -= f2 * mx * mz;
        H[3] += fma(f, c.c22, -f2 * my * my);
        H[4] -= f2 * my * mz;
        H[5] += fma(f, c.c33, -f2 * mz * mz);
    }
#pragma unroll 1
    for (int i = 0; i < C::kRAD; ++i) {
        if (i >= P.n_rad) break;
        const DevRad &c = P.rad[i];
        const double wy = y * c.i1, wz = z * c.i2;
        const double m2 = fma(x, x, fma(y, wy, fma(z, wz, TINY)));
        double f, d2;
        rad_profile(c, m2, f, &d2, nullptr);
        g[0] = fma(f, x, g[0]); g[1] = fma(f, wy, g[1]); g[2] = fma(f, wz, g[2]);
        // H += (F'/m) diag(1, i1, i2) + (F'' - F'/m)/m^2 w w^T
        const double w = (d2 - f) / m2;
        H[0] += fma(w * x, x, f);
        H[1] = fma(w * x, wy, H[1]);
        H[2] = fma(w * x, wz, H[2]);
        H[3] += fma(w * wy, wy, f * c.i1);
        H[4] = fma(w * wy, wz, H[4]);
        H[5] += fma(w * wz, wz, f * c.i2);
    }
    for (int i = 0; i < C::kHARM; ++i) {
        if (i >= P.n_harm) break;
        g[0] = fma(P.harm[i].w2x, x, g[0]); g[1] = fma(P.harm[i].w2y, y, g[1]); g[2] = fma(P.harm[i].w2z, z, g[2]);
        H[0] += P.harm[i].w2x; H[3] += P.harm[i].w2y; H[5] += P.harm[i].w2z;
    }
    for (int i = 0; i < C::kHENON; ++i) {
        if (i >= P.n_henon) break;
        const double k = P.henon[i].k, it2 = P.henon[i].it2;
        g[0] = fma(fma(2.0 * k * x, y, x), it2, g[0]);
        g[1] = fma(fma(k, x * x - y * y, y), it2, g[1]);
        H[0] += fma(2.0 * k, y, 1.0) * it2;
        H[1] += 2.0 * k * x * it2;
        H[3] += fma(-2.0 * k, y, 1.0) * it2;
    }
}

// ---------------------------------------------------------------------------------------------
// A time-dependent composite (LinearParameter, DevTD) frozen at one time, per thread: the derived constants of its
// components as build_devpot() forms them on the host for a static potential, so that potential_value / grad_hess
// evaluate it through the same code (stream release at each particle's own release time, energies along an orbit).
// Only the member names matter to those templates; the kinds a DevTD cannot hold have length-1 dummies (n = 0).
struct FrozenPot {
    int n_mn, n_hern, n_nfw, n_plc, n_log, n_iso, n_satoh, n_rad, n_harm, n_henon;
    DevMN mn[MAX_MN];
    DevHern hern[MAX_HERN];
    DevNFW nfw[MAX_NFW];
    DevIso iso[MAX_ISO];
    DevSatoh satoh[MAX_SATOH];
    DevRad rad[MAX_RAD];
    DevPLC plc[1];
    DevLog lg[1];
    DevHarm harm[1];
    DevHenon henon[1];
};
// false: more components of one kind than the static arrays hold (the host refuses such composites up front)
__device__ __forceinline__ bool freeze_td(const DevTD &R, double t, FrozenPot &F) {
    F.n_mn = F.n_hern = F.n_nfw = F.n_plc = F.n_log = F.n_iso = F.n_satoh = F.n_rad = F.n_harm = F.n_henon = 0;
    for (int i = 0; i < R.n; ++i) {
        const double p0 = fma(R.dp[i][0], t, R.p[i][0]), p1 = fma(R.dp[i][1], t, R.p[i][1]);
        const double p2 = fma(R.dp[i][2], t, R.p[i][2]), p3 = fma(R.dp[i][3], t, R.p[i][3]);
        const double GM = R.G * p0;
        switch (R.kind[i]) {
        case 0: {  // Miyamoto-Nagai (m, a, b)
            if (F.n_mn >= MAX_MN) return false;
            DevMN &m = F.mn[F.n_mn++];
            m.GM = GM; m.a = p1; m.b2 = p2 * p2; m.ab2 = p1 * m.b2;
            if (m.b2 == 0.0) m.b2 = TINY;
            break;
        }
        case 1: {  // Hernquist (m, r_s)
            if (F.n_hern >= MAX_HERN) return false;
            DevHern &h = F.hern[F.n_hern++];
            h.GM = GM; h.c = p1;
            break;
        }
        case 2: {  // NFW (m, r_s)
            if (F.n_nfw >= MAX_NFW) return false;
            DevNFW &n = F.nfw[F.n_nfw++];
            n.GM = GM; n.rs = p1; n.inv_rs = 1.0 / p1; n.GM_inv_rs = GM / p1; n.GM_rs3 = GM / (p1 * p1 * p1); n.pad_ = 0.0;
            break;
        }
        case 5: {  // Isochrone (m, b)
            if (F.n_iso >= MAX_ISO) return false;
            DevIso &c = F.iso[F.n_iso++];
            c.GM = GM; c.b = p1; c.b2 = p1 * p1;
            break;
        }
        case 6: {  // Satoh (m, a, b)
            if (F.n_satoh >= MAX_SATOH) return false;
            DevSatoh &m = F.satoh[F.n_satoh++];
            m.GM = GM; m.a = p1; m.b2 = p2 * p2; m.ab2 = p1 * m.b2;
            if (m.b2 == 0.0) m.b2 = TINY;
            break;
        }
        default: {  // 7: triaxial Hernquist (m, r_s, q1, q2); 8: Jaffe (m, r_s)
            if (F.n_rad >= MAX_RAD) return false;
            DevRad &r = F.rad[F.n_rad++];
            r.pad_ = 0; r.K = GM; r.a = p1; r.b = 0.0; r.i1 = r.i2 = 1.0;
            if (R.kind[i] == 7) { r.profile = RAD_HERNQUIST; r.i1 = 1.0 / (p2 * p2); r.i2 = 1.0 / (p3 * p3); }
            else r.profile = RAD_JAFFE;
            break;
        }
        }
    }
    return true;
}

}  // namespace gx
