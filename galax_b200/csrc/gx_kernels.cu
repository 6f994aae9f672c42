// gx_kernels.cu -- sm_100a kernels + C ABI of the galax hot path (see include/galax_b200.h).
//
//   K1  k_potential_eval     bulk Phi / grad / acc / Hessian                       (HBM-bound)
//   K2  k_integrate_fixed    one thread per particle, state in registers,
//                            SemiImplicitEuler / LeapfrogMidpoint + ConstantStepSize (FP64-pipe-bound)
//   K3  k_integrate_dopri8   persistent warps, per-lane work queue, per-particle PID step control,
//                            Dopri8 in Nystrom form with all 14 stage accelerations in registers
//   K4  k_stream_release     Fardal+15 / Chen+24 release conditions (then K3 with per-particle t0)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo (no fast-math).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <mutex>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/galax_b200.h"
#include "gx_potential.cuh"
#include "gx_tables.h"
#include "plc_table.h"

namespace gx {

// ================================================================================================
// host: gx_potential -> DevPot
// ================================================================================================

enum Model { MODEL_GENERIC = 0, MODEL_MW = 1, MODEL_MW2022 = 2, MODEL_BOVY = 3 };

// Q(a, x) ~ x^(a-1) e^-x / Gamma(a) (1 + (a-1)/x + ...): find where it drops below 2^-55 so that P == 1 in fp64.
static void fill_gamma_tab(GammaTab &g, double a) {
    g.a = a;
    g.lgam = lgamma(a);
    double x = 30.0;
    while (x < 80.0) {
        double q = exp((a - 1.0) * log(x) - x - g.lgam) * (1.0 + fabs(a - 1.0) / x);
        if (q < 2.7e-17) break;
        x += 0.25;
    }
    g.xcut = x;
    for (int n = 0; n < PLC_NT; ++n) g.inv[n] = 1.0 / (a + n);
}

// How an entry treats LinearParameter rates (gx_component.dp): freeze the parameters at one time (bulk evaluation),
// carry the raw parameters to the device for per-stage evaluation (integrators), or refuse.
enum TdMode { TD_REJECT = 0, TD_FREEZE = 1, TD_INTEGRATE = 2 };

// may_upload = false (the caller's stream is being captured into a graph): a combined spherical table that is not in the
// cache yet is not fitted and uploaded now -- the composite then runs through the runtime-count kernels.
// tabs: TAB_WIDE when the caller's kernels look the spherical components up in the combined table (the integrators); it is
// fitted and uploaded on first use only by the entries that need it.
enum { TAB_NONE = 0, TAB_WIDE = 2 };
static int build_devpot(const gx_potential *pot_in, DevPot &D, Model &model, bool use_device = true,
                        TdMode td_mode = TD_REJECT, double t_freeze = 0.0, bool may_upload = true, int tabs = TAB_NONE) {
    if (!pot_in || pot_in->n < 0 || pot_in->n > GX_MAX_COMPONENTS) return GX_ERR_BADARG;
    memset(&D, 0, sizeof D);
    gx_potential frozen = *pot_in;
    bool td = false;
    for (int i = 0; i < frozen.n; ++i)
        for (int k = 0; k < 8; ++k) {
            if (frozen.c[i].dp[k] != 0.0) td = true;
            frozen.c[i].p[k] += frozen.c[i].dp[k] * (td_mode == TD_FREEZE ? t_freeze : 0.0);
        }
    if (td && td_mode == TD_REJECT) return GX_ERR_UNSUPPORTED;
    if (td && td_mode == TD_INTEGRATE) {
        D.td.n = frozen.n;
        D.td.G = frozen.G;
        for (int i = 0; i < frozen.n; ++i) {
            const int kind = pot_in->c[i].kind;
            if (kind != GX_KIND_MIYAMOTO_NAGAI && kind != GX_KIND_HERNQUIST && kind != GX_KIND_NFW &&
                kind != GX_KIND_ISOCHRONE && kind != GX_KIND_SATOH && kind != GX_KIND_TRIAXIAL_HERNQUIST &&
                kind != GX_KIND_JAFFE)
                return GX_ERR_UNSUPPORTED;
            D.td.kind[i] = kind;
            for (int k = 0; k < TD_NP; ++k) { D.td.p[i][k] = pot_in->c[i].p[k]; D.td.dp[i][k] = pot_in->c[i].dp[k]; }
            for (int k = TD_NP; k < 8; ++k)
                if (pot_in->c[i].dp[k] != 0.0) return GX_ERR_UNSUPPORTED;
        }
    }
    const gx_potential *pot = &frozen;
    const double G = pot->G;
    for (int i = 0; i < pot->n; ++i) {
        const gx_component &c = pot->c[i];
        switch (c.kind) {
        case GX_KIND_MIYAMOTO_NAGAI: {
            if (D.n_mn >= MAX_MN) return GX_ERR_UNSUPPORTED;
            DevMN &m = D.mn[D.n_mn++];
            m.GM = G * c.p[0];
            m.a = c.p[1];
            m.b2 = c.p[2] * c.p[2];
            m.ab2 = c.p[1] * m.b2;
            // b = 0 (KuzminPotential): zeta = sqrt(z^2 + b^2) vanishes in the disk plane and 0 * (1/0) would be NaN.
            // With the smallest normal number under the root the in-plane z-force is exactly 0, as the reference's
            // |z| form gives (builtin/kuzmin.py:82-84); for b > 0 nothing changes (b^2 + tiny == b^2).
            if (m.b2 == 0.0) m.b2 = 2.2250738585072014e-308;
            break;
        }
        case GX_KIND_HERNQUIST: {
            if (D.n_hern >= MAX_HERN) return GX_ERR_UNSUPPORTED;
            DevHern &h = D.hern[D.n_hern++];
            h.GM = G * c.p[0];
            h.c = c.p[1];
            break;
        }
        case GX_KIND_NFW: {
            if (D.n_nfw >= MAX_NFW) return GX_ERR_UNSUPPORTED;
            DevNFW &n = D.nfw[D.n_nfw++];
            n.GM = G * c.p[0];
            n.rs = c.p[1];
            n.inv_rs = 1.0 / c.p[1];
            n.GM_inv_rs = n.GM / c.p[1];
            n.GM_rs3 = n.GM / (c.p[1] * c.p[1] * c.p[1]);
            n.pad_ = 0.0;
            if (use_device && D.nfw_tab == nullptr) D.nfw_tab = nfw_table();
            break;
        }
        case GX_KIND_POWERLAWCUTOFF: {
            if (D.n_plc >= MAX_PLC) return GX_ERR_UNSUPPORTED;
            double alpha = c.p[1], rc = c.p[2];
            if (!(alpha >= 0.0 && alpha < 2.0)) return GX_ERR_UNSUPPORTED;  // Phi needs Gamma(1 - alpha/2)
            DevPLC &p = D.plc[D.n_plc++];
            p.GM = G * c.p[0];
            p.inv_rc = 1.0 / rc;
            p.GM_rc3 = p.GM / (rc * rc * rc);
            p.tab = use_device ? plc_table_for(1.5 - alpha / 2) : nullptr;
            fill_gamma_tab(p.ga, 1.5 - alpha / 2);
            fill_gamma_tab(p.ga2, 1.0 - alpha / 2);
            p.tail = tgamma(p.ga2.a) / (rc * tgamma(p.ga.a));
            break;
        }
        case GX_KIND_LOGARITHMIC: {
            if (D.n_log >= MAX_LOG) return GX_ERR_UNSUPPORTED;
            const double vc = c.p[0], rs = c.p[1], q1 = c.p[2], q2 = c.p[3], q3 = c.p[4], phi = c.p[5];
            if (!(q1 > 0.0 && q2 > 0.0 && q3 > 0.0)) return GX_ERR_BADARG;
            const double sp = sin(phi), cp = cos(phi);
            DevLog &l = D.lg[D.n_log++];
            l.vc2 = vc * vc;
            l.rs2 = rs * rs;
            l.c11 = cp * cp / (q1 * q1) + sp * sp / (q2 * q2);
            l.c22 = sp * sp / (q1 * q1) + cp * cp / (q2 * q2);
            l.c12 = cp * sp * (1.0 / (q1 * q1) - 1.0 / (q2 * q2));
            l.c33 = 1.0 / (q3 * q3);
            break;
        }
        case GX_KIND_ISOCHRONE: {
            if (D.n_iso >= MAX_ISO) return GX_ERR_UNSUPPORTED;
            DevIso &s = D.iso[D.n_iso++];
            s.GM = G * c.p[0];
            s.b = c.p[1];
            s.b2 = c.p[1] * c.p[1];
            break;
        }
        case GX_KIND_SATOH: {
            if (D.n_satoh >= MAX_SATOH) return GX_ERR_UNSUPPORTED;
            DevSatoh &m = D.satoh[D.n_satoh++];
            m.GM = G * c.p[0];
            m.a = c.p[1];
            m.b2 = c.p[2] * c.p[2];
            m.ab2 = c.p[1] * m.b2;
            if (m.b2 == 0.0) m.b2 = 2.2250738585072014e-308;  // (as for Miyamoto-Nagai above)
            break;
        }
        case GX_KIND_TRIAXIAL_HERNQUIST:
        case GX_KIND_JAFFE:
        case GX_KIND_BURKERT:
        case GX_KIND_STONE: {
            if (D.n_rad >= MAX_RAD) return GX_ERR_UNSUPPORTED;
            DevRad &r = D.rad[D.n_rad++];
            r.pad_ = 0;
            r.i1 = r.i2 = 1.0;
            r.b = 0.0;
            if (c.kind == GX_KIND_TRIAXIAL_HERNQUIST) {
                if (!(c.p[2] > 0.0 && c.p[3] > 0.0)) return GX_ERR_BADARG;
                r.profile = RAD_HERNQUIST; r.K = G * c.p[0]; r.a = c.p[1];
                r.i1 = 1.0 / (c.p[2] * c.p[2]); r.i2 = 1.0 / (c.p[3] * c.p[3]);
            } else if (c.kind == GX_KIND_JAFFE) {
                r.profile = RAD_JAFFE; r.K = G * c.p[0]; r.a = c.p[1];
            } else if (c.kind == GX_KIND_BURKERT) {
                r.profile = RAD_BURKERT; r.K = G * c.p[0] / (3.0 * log(2.0) - M_PI / 2); r.a = c.p[1]; r.b = 1.0 / c.p[1];
            } else {
                if (!(c.p[1] > 0.0 && c.p[2] > 0.0) || c.p[1] == c.p[2]) return GX_ERR_BADARG;
                r.profile = RAD_STONE; r.K = 2.0 * G * c.p[0] / (M_PI * (c.p[2] - c.p[1])); r.a = c.p[1]; r.b = c.p[2];
            }
            break;
        }
        case GX_KIND_HARMONIC: {
            if (D.n_harm >= MAX_HARM) return GX_ERR_UNSUPPORTED;
            DevHarm &h = D.harm[D.n_harm++];
            h.w2x = c.p[0] * c.p[0]; h.w2y = c.p[1] * c.p[1]; h.w2z = c.p[2] * c.p[2];
            break;
        }
        case GX_KIND_HENON_HEILES: {
            if (D.n_henon >= MAX_HENON) return GX_ERR_UNSUPPORTED;
            DevHenon &h = D.henon[D.n_henon++];
            h.k = c.p[0]; h.it2 = 1.0 / (c.p[1] * c.p[1]);
            break;
        }
        default:
            return GX_ERR_UNSUPPORTED;
        }
    }
    model = MODEL_GENERIC;
    if (D.td.n == 0 && D.n_log + D.n_iso + D.n_satoh + D.n_rad + D.n_harm + D.n_henon == 0) {
        if (D.n_mn == 1 && D.n_hern == 2 && D.n_nfw == 1 && D.n_plc == 0) model = MODEL_MW;
        if (D.n_mn == 3 && D.n_hern == 2 && D.n_nfw == 1 && D.n_plc == 0 && D.mn[0].b2 == D.mn[1].b2 &&
            D.mn[0].b2 == D.mn[2].b2)
            model = MODEL_MW2022;  // CountsMW2022 evaluates sqrt(z^2 + b^2) once: needs one b (an MN3 disk)
        if (D.n_mn == 1 && D.n_hern == 0 && D.n_nfw == 1 && D.n_plc == 1 && D.plc[0].tab != nullptr)
            model = MODEL_BOVY;  // (the integrators stage the bulge's force table in shared memory)
        // ... and the NFW force table for MW / MW2022: without it (allocation failed) they run as runtime composites
        if ((model == MODEL_MW || model == MODEL_MW2022) && use_device && D.nfw_tab == nullptr) model = MODEL_GENERIC;
#if GX_SPH_TABLE
        // the composite's combined spherical table S(r^2) (fitted and uploaded on first use of these parameters): the
        // three named models need it (else they run as runtime composites); any other composite of the four basic
        // kinds takes it when it has spherical components (CountsBasicTab), and runs without it otherwise
        if (use_device && tabs != TAB_NONE && (model != MODEL_GENERIC || D.n_hern + D.n_nfw + D.n_plc > 0)) {
            std::vector<SphComp> cs;
            for (int i = 0; i < pot->n; ++i) {
                const gx_component &c = pot->c[i];
                if (c.kind == GX_KIND_HERNQUIST || c.kind == GX_KIND_NFW) cs.push_back({c.kind, G * c.p[0], c.p[1], 0.0});
                if (c.kind == GX_KIND_POWERLAWCUTOFF) cs.push_back({c.kind, G * c.p[0], c.p[2], 1.5 - c.p[1] / 2});
            }
            D.sph_j0w = (unsigned)((1023 + sph_e_lo(cs)) << SPHW_SUB_BITS);
            D.sph_wide = sph_table_for(cs, nullptr, may_upload);
            if (D.sph_wide == nullptr) model = MODEL_GENERIC;
        }
#endif
    }
    return 0;
}

#define GX_DISPATCH_MODEL(model, CALL)                 \
    switch (model) {                                   \
    case MODEL_MW: { using C = CountsMW; CALL; } break;         \
    case MODEL_MW2022: { using C = CountsMW2022; CALL; } break; \
    case MODEL_BOVY: { using C = CountsBovy; CALL; } break;     \
    default: { using C = CountsRuntime; CALL; } break;          \
    }

// A runtime composite of the four basic kinds with constant parameters: the integrators run it through CountsBasic.
static inline bool is_basic_composite(const DevPot &D, Model model) {
    return model == MODEL_GENERIC && D.td.n == 0 &&
           D.n_log + D.n_iso + D.n_satoh + D.n_rad + D.n_harm + D.n_henon == 0;
}

static inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : GX_ERR_CUDA; }
static inline bool stream_is_capturing(void *stream) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing((cudaStream_t)stream, &cap) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return cap != cudaStreamCaptureStatusNone;
}

// ================================================================================================
// K1 bulk evaluation
// ================================================================================================

struct EvalArgs {
    const double *xyz;
    double *phi, *grad, *acc, *hess;
    long long N;
    unsigned what;
    int tma_ok;  // every tile address is 16-byte aligned: full tiles may move by cp.async.bulk (else plain loads/stores)
};

// Light outputs (Phi / grad / acc only): one thread per point, direct loads and stores; the 24-byte stride
// of a lane is absorbed by L1/L2 and this measured faster (4.3 TB/s) than staging for these modes.
template <class C>
__global__ void __launch_bounds__(256) k_potential_eval_direct(const __grid_constant__ DevPot P, const EvalArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += stride) {
        const double x = __ldg(a.xyz + 3 * i), y = __ldg(a.xyz + 3 * i + 1), z = __ldg(a.xyz + 3 * i + 2);
        if (a.what & GX_PHI) a.phi[i] = potential_value<C>(P, x, y, z);
        if (a.what & (GX_GRAD | GX_ACC)) {
            double g0, g1, g2;
            gradient<C>(P, x, y, z, g0, g1, g2);
            if (a.what & GX_GRAD) { a.grad[3 * i] = g0; a.grad[3 * i + 1] = g1; a.grad[3 * i + 2] = g2; }
            if (a.what & GX_ACC) { a.acc[3 * i] = -g0; a.acc[3 * i + 1] = -g1; a.acc[3 * i + 2] = -g2; }
        }
    }
}

// ---- TMA (bulk async copy) + mbarrier helpers: 1-D cp.async.bulk global <-> shared, sm_90+ ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GX_DONE;\n"
        "bra GX_WAIT;\n"
        "GX_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// With the Hessian (72 B/point out) the natural [N,3] / [N,3,3] layouts would make every store instruction
// touch 32 different sectors, so tiles of 256 points are staged through shared memory and moved by the TMA engine:
// one elected thread issues a 6 KB cp.async.bulk load per tile (completion on an mbarrier, the next tile's load
// is in flight while the current one is computed) and 6/6/18 KB cp.async.bulk stores of the results.
// Persistent CTAs, grid-stride over tiles; a ragged last tile takes the plain load/store path.
// The OUTPUT staging is double buffered as well, so the thread that drives the TMA engine never waits for the store
// it has just issued -- it only makes sure (wait_group.read 1) that the store of two tiles ago has left shared
// memory before that buffer is refilled (ncu on the single-buffered first version: `barrier` was the top stall).
// Dynamic shared memory, sized by the requested outputs: [2][TILE*3] inputs, then per stage the grad / acc /
// Hessian tiles.  Measured ceiling of this traffic pattern (24 B in, 96 B out per point): the same pipeline with the
// arithmetic removed, or with plain coalesced stores instead of TMA stores, runs at the same 5.8-5.9 TB/s.
constexpr int EVAL_TILE = 256;
template <class C, int TILE>
__global__ void __launch_bounds__(TILE) k_potential_eval(const __grid_constant__ DevPot P, const EvalArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ alignas(8) unsigned long long bar[2];
    const bool want_grad = (a.what & GX_GRAD) != 0, want_acc = (a.what & GX_ACC) != 0, want_h = (a.what & GX_HESS) != 0;
    const bool want_g = want_grad || want_acc;
    double *s_in = reinterpret_cast<double *>(smem_raw);                       // [2][TILE*3]
    const int per_stage = TILE * ((want_grad ? 3 : 0) + (want_acc ? 3 : 0) + (want_h ? 9 : 0));
    double *s_out = s_in + 2 * TILE * 3;                                        // [2][per_stage]
    const int off_a = want_grad ? TILE * 3 : 0, off_h = off_a + (want_acc ? TILE * 3 : 0);
    const int tid = threadIdx.x;
    const long long n_tiles = (a.N + TILE - 1) / TILE;
    auto tile_cnt = [&](long long t) { long long r = a.N - t * TILE; return (int)(r < TILE ? r : TILE); };
    auto tile_full = [&](long long t) { return a.tma_ok && tile_cnt(t) == TILE; };  // moved by the TMA engine
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles && tile_full(tile)) {
        mbar_expect_tx(&bar[0], TILE * 24);
        tma_load_1d(s_in, a.xyz + tile * TILE * 3, TILE * 24, &bar[0]);
    }
    unsigned it = 0;
    for (; tile < n_tiles; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        const long long base = tile * TILE;
        const int cnt = tile_cnt(tile);
        const bool full = tile_full(tile);
        const long long next = tile + gridDim.x;
        double *in = s_in + stage * TILE * 3, *out = s_out + stage * per_stage;
        if (tid == 0) {
            if (next < n_tiles && tile_full(next)) {  // prefetch the next tile's positions
                mbar_expect_tx(&bar[stage ^ 1], TILE * 24);
                tma_load_1d(s_in + (stage ^ 1) * TILE * 3, a.xyz + next * TILE * 3, TILE * 24, &bar[stage ^ 1]);
            }
            // the store issued two tiles ago read from `out`: it must have left shared memory (one group may stay pending)
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncthreads();  // `out` is free for everybody
        if (full) {
            mbar_wait(&bar[stage], (it >> 1) & 1);
        } else {
            for (int k = tid; k < cnt * 3; k += TILE) in[k] = __ldg(a.xyz + base * 3 + k);
            __syncthreads();
        }
        if (tid < cnt) {
            const double x = in[3 * tid], y = in[3 * tid + 1], z = in[3 * tid + 2];
            if (a.what & GX_PHI) a.phi[base + tid] = potential_value<C>(P, x, y, z);  // already coalesced
            double g[3] = {0, 0, 0};
            if (want_h) {
                double H[6];
                grad_hess<C>(P, x, y, z, g, H);
                double *h = out + off_h + 9 * tid;
                h[0] = H[0]; h[1] = H[1]; h[2] = H[2];
                h[3] = H[1]; h[4] = H[3]; h[5] = H[4];
                h[6] = H[2]; h[7] = H[4]; h[8] = H[5];
            } else if (want_g) {
                gradient<C>(P, x, y, z, g[0], g[1], g[2]);
            }
            if (want_grad) { out[3 * tid] = g[0]; out[3 * tid + 1] = g[1]; out[3 * tid + 2] = g[2]; }
            if (want_acc) { out[off_a + 3 * tid] = -g[0]; out[off_a + 3 * tid + 1] = -g[1]; out[off_a + 3 * tid + 2] = -g[2]; }
        }
        if (full) {
            fence_async_smem();  // make the generic-proxy writes above visible to the async (TMA) proxy
            __syncthreads();
            if (tid == 0) {
                if (want_grad) tma_store_1d(a.grad + base * 3, out, TILE * 24);
                if (want_acc) tma_store_1d(a.acc + base * 3, out + off_a, TILE * 24);
                if (want_h) tma_store_1d(a.hess + base * 9, out + off_h, TILE * 72);
                tma_commit();  // not waited for here: the next tile computes into the other buffer meanwhile
            }
        } else {
            __syncthreads();
            if (want_grad) for (int k = tid; k < cnt * 3; k += TILE) a.grad[base * 3 + k] = out[k];
            if (want_acc) for (int k = tid; k < cnt * 3; k += TILE) a.acc[base * 3 + k] = out[off_a + k];
            if (want_h) for (int k = tid; k < cnt * 9; k += TILE) a.hess[base * 9 + k] = out[off_h + k];
            if (tid == 0) tma_commit();  // keep one commit group per tile so that wait_group.read 1 stays aligned
        }
    }
    if (tid == 0) tma_wait_all();
}

// ================================================================================================
// K2 fixed-step integrator
// ================================================================================================

// ---- saved states in the reference layout [N,T,3] (dynamics/_src/orbit/register_dfx.py:78-82) ----
// A thread owns a particle.  Its saves k, k+1, ... are 24 bytes apart in q[N,T,3] (and in p) and the next particle's
// row is 24 T bytes away, so a store instruction of the warp touches 32 different 32-byte sectors with 8 bytes each;
// and when the saves of one particle are far apart in time, a partially written sector leaves L2 before its
// neighbours arrive (ncu, round 1, Dopri8 with 10 saves: 3.75x the algorithmic DRAM write traffic, plus the
// read-modify-write reads).  So each lane stages up to four saves in shared memory ([slot][6][lane]: conflict-free,
// private to the lane, no barrier) and writes them out when the run ENDS ON A SECTOR BOUNDARY of the output:
// 24 (T i + k + 1) = 0 (mod 32)  <=>  (T i + k + 1) mod 4 == 0.  Every flush but a row's first and last is then
// exactly three whole sectors of q and three of p, written by back-to-back stores of one lane (L2 sees whole sectors).
// Used when the caller asks for GX_LAYOUT_NT3 with T >= SAVE_STAGE_MIN_T; GX_LAYOUT_T3N and short rows store directly.
constexpr int SAVE_SLOTS = 4;  // saves per run (the last one of a run is written straight from registers)
constexpr int SAVE_STAGE_MIN_T = 4;
// Nothing but the save index k is carried between saves: the sector phase of save k is (T i + k) mod 4, the number of
// staged saves is that phase (or k itself inside the row's first, shorter run).
// (dynamic shared memory: the wide force table of the fixed-step kernels first, `off` doubles of it, then the staging)
__device__ __forceinline__ double *save_buf(int off) { return dyn_smem() + off; }
template <class A>
__device__ __forceinline__ int save_staged(const A &a, long long i, int k) {  // saves staged before save k
    const int ph = (int)(((unsigned)a.T * (unsigned)i + (unsigned)k) & 3u);
    return ph < k ? ph : k;
}
// (a rolled loop: the integrators' register allocation should not pay for this cold path)
__device__ __forceinline__ void save_flush_n(double *qd, double *pd, int n, int nv, int off) {
    const int bd = blockDim.x;
    const double *b = save_buf(off) + threadIdx.x;
#pragma unroll 1
    for (int s = 0; s < n; ++s, b += nv * bd, qd += 3, pd += 3) {
        qd[0] = b[0]; qd[1] = b[bd]; qd[2] = b[2 * bd];
        pd[0] = b[3 * bd]; pd[1] = b[4 * bd]; pd[2] = b[5 * bd];
    }
}
template <class A>
__device__ __forceinline__ void save_flush(const A &a, long long i, int kend) {  // write out what is staged before kend
    if (!a.stage) return;
    const int n = save_staged(a, i, kend);
    if (n > 0) save_flush_n(a.q + i * a.sn + 3LL * (kend - n), a.p + i * a.sn + 3LL * (kend - n), n, a.epi.nv, a.stage_off);
}
// Where save k of particle i goes: component c of q at q[c * st], of p at p[c * st] -- the output itself (direct
// stores, or the save that closes a run) or the lane's staging column.  The caller stores the six values as it
// computes them (nothing extra is live across the dense-output evaluation) and then calls save_commit().
struct SaveDst {
    double *q, *p;
    long long st;
};
template <class A>
__device__ __forceinline__ SaveDst save_dst(const A &a, long long i, int k) {
    SaveDst d;
    if (!a.stage) {
        d.q = a.q + i * a.sn + k * a.sk; d.p = a.p + i * a.sn + k * a.sk; d.st = a.sc;
    } else if ((((unsigned)a.T * (unsigned)i + (unsigned)k + 1u) & 3u) != 0u) {  // not at a sector boundary: stage
        const int bd = blockDim.x;
        d.q = save_buf(a.stage_off) + threadIdx.x + (a.epi.nv * save_staged(a, i, k)) * bd; d.p = d.q + 3 * bd; d.st = bd;
    } else {  // save k ends on a sector boundary: it goes straight out, behind the staged ones (save_commit)
        d.q = a.q + i * a.sn + 3LL * k; d.p = a.p + i * a.sn + 3LL * k; d.st = 1;
    }
    return d;
}
template <class A>
__device__ __forceinline__ void save_commit(const A &a, long long i, int k) {
    if (a.stage && (((unsigned)a.T * (unsigned)i + (unsigned)k + 1u) & 3u) == 0u) save_flush(a, i, k);
}
template <class A>
__device__ __forceinline__ void save_put(const A &a, long long i, int k, double qx, double qy, double qz, double px,
                                         double py, double pz) {
    const SaveDst d = save_dst(a, i, k);
    d.q[0] = qx; d.q[d.st] = qy; d.q[2 * d.st] = qz;
    d.p[0] = px; d.p[d.st] = py; d.p[2 * d.st] = pz;
    save_commit(a, i, k);
}
// the saves a failed particle never reached: NaN, like an unfilled diffrax buffer
template <class A>
__device__ __forceinline__ void save_fill_nan(const A &a, long long i, int k) {
    save_flush(a, i, k);
    const double NANV = __longlong_as_double(0x7ff8000000000000LL);
    double *qo = a.q + i * a.sn, *po = a.p + i * a.sn;
    for (; k < a.T; ++k) {
        qo[k * a.sk] = NANV; qo[k * a.sk + a.sc] = NANV; qo[k * a.sk + 2 * a.sc] = NANV;
        po[k * a.sk] = NANV; po[k * a.sk + a.sc] = NANV; po[k * a.sk + 2 * a.sc] = NANV;
    }
}
#ifndef GX_SAVE_STAGE
#define GX_SAVE_STAGE 1
#endif
static inline size_t save_stage_bytes(int layout, int T, int block, int nv = 6) {
    return (GX_SAVE_STAGE && layout == GX_LAYOUT_NT3 && T >= SAVE_STAGE_MIN_T) ? (size_t)block * (SAVE_SLOTS - 1) * nv * sizeof(double) : 0;
}

// ---- Orbit post-processing fused into the save path --------------------------------------------------
// Orbit.total_energy / angular_momentum (coordinates/_src/pscs/base.py:182-330) and tidal_tensor along the orbit
// (potential/_src/register_funcs.py:347-377) are functions of the saved (q, p) alone; evaluated on the values while they
// are still in registers they cost one potential (and Hessian) evaluation per save and no second pass over the saved
// states (C2: 48 GB).  They travel through the same per-lane staging as q and p -- a staged save is nv doubles: q, p,
// then L (3) and E (1), then the nine tidal-tensor entries -- and leave with the same flush: 4 saves of E are exactly
// one 32-byte sector, of L three, of the tensor nine.  Kernels are instantiated with EPI = true only for these calls;
// the default kernels do not contain any of this.
struct EpiOut {
    double *E, *L, *TT;  // [N,T], [N,T,3], [N,T,3,3] (GX_LAYOUT_NT3) / [T,N], [T,3,N], [T,9,N] (GX_LAYOUT_T3N); any may be null
    int t3n;
    int nv;  // doubles per staged save: 6, 10 (L, E) or 19 (L, E, tidal tensor)
};
// (explicit FMA forms: the fused epilogue and the stand-alone pass gx_energy_angmom give the same bits)
__device__ __forceinline__ double kinetic_plus(double phi, double vx, double vy, double vz) {
    return fma(0.5, fma(vx, vx, fma(vy, vy, vz * vz)), phi);
}
__device__ __forceinline__ void cross3(double x, double y, double z, double vx, double vy, double vz, double &lx,
                                       double &ly, double &lz) {
    lx = fma(y, vz, -(z * vy)); ly = fma(z, vx, -(x * vz)); lz = fma(x, vy, -(y * vx));
}
template <class A>
__device__ __forceinline__ void epilogue_flush(const A &a, long long i, int kend) {  // staged extras before save kend
    if (!a.stage) return;
    const int n = save_staged(a, i, kend);
    const EpiOut &e = a.epi;
    const int bd = blockDim.x;
    const double *b = save_buf(a.stage_off) + threadIdx.x + 6 * bd;
    long long r = i * a.T + (kend - n);
#pragma unroll 1
    for (int s = 0; s < n; ++s, b += e.nv * bd, ++r) {
        if (e.L) { e.L[3 * r] = b[0]; e.L[3 * r + 1] = b[bd]; e.L[3 * r + 2] = b[2 * bd]; }
        if (e.E) e.E[r] = b[3 * bd];
        if (e.TT) {
#pragma unroll
            for (int c = 0; c < 9; ++c) e.TT[9 * r + c] = b[(4 + c) * bd];
        }
    }
}
// E, L and the tidal tensor of save k of particle i, given the saved state (out of line: cold beside the step loop)
template <class C, class A>
__device__ __noinline__ void epilogue_put(const DevPot *Pp, const A *ap, long long i, int k, double qx, double qy,
                                          double qz, double px, double py, double pz) {
    const A &a = *ap;
    const EpiOut &e = a.epi;
    double v[13];
    cross3(qx, qy, qz, px, py, pz, v[0], v[1], v[2]);
    v[3] = e.E ? kinetic_plus(potential_value<C>(*Pp, qx, qy, qz), px, py, pz) : 0.0;
    if (e.TT) {
        double g[3], H[6];
        grad_hess<C>(*Pp, qx, qy, qz, g, H);
        const double tr3 = (H[0] + H[3] + H[5]) * (1.0 / 3.0);  // J - tr(J)/3 I
        v[4] = H[0] - tr3; v[5] = H[1]; v[6] = H[2];
        v[7] = H[1]; v[8] = H[3] - tr3; v[9] = H[4];
        v[10] = H[2]; v[11] = H[4]; v[12] = H[5] - tr3;
    }
    if (e.t3n) {
        const long long N = a.N;
        if (e.E) e.E[(long long)k * N + i] = v[3];
        if (e.L) { e.L[(3LL * k) * N + i] = v[0]; e.L[(3LL * k + 1) * N + i] = v[1]; e.L[(3LL * k + 2) * N + i] = v[2]; }
        if (e.TT) {
#pragma unroll
            for (int c = 0; c < 9; ++c) e.TT[(9LL * k + c) * N + i] = v[4 + c];
        }
        return;
    }
    if (a.stage && (((unsigned)a.T * (unsigned)i + (unsigned)k + 1u) & 3u) != 0u) {  // inside a run: stage
        const int bd = blockDim.x;
        double *b = save_buf(a.stage_off) + threadIdx.x + (e.nv * save_staged(a, i, k) + 6) * bd;
        b[0] = v[0]; b[bd] = v[1]; b[2 * bd] = v[2]; b[3 * bd] = v[3];
        if (e.TT) {
#pragma unroll
            for (int c = 0; c < 9; ++c) b[(4 + c) * bd] = v[4 + c];
        }
        return;
    }
    epilogue_flush(a, i, k);  // the save that closes a run (or direct stores): behind the staged ones
    const long long r = i * a.T + k;
    if (e.L) { e.L[3 * r] = v[0]; e.L[3 * r + 1] = v[1]; e.L[3 * r + 2] = v[2]; }
    if (e.E) e.E[r] = v[3];
    if (e.TT) {
#pragma unroll
        for (int c = 0; c < 9; ++c) e.TT[9 * r + c] = v[4 + c];
    }
}
// end of a particle: what is still staged, then NaN for the saves it never reached (as save_fill_nan does for q, p)
template <class A>
__device__ __noinline__ void epilogue_finish(const A *ap, long long i, int k) {
    const A &a = *ap;
    const EpiOut &e = a.epi;
    epilogue_flush(a, i, k);
    const double NANV = __longlong_as_double(0x7ff8000000000000LL);
    for (; k < a.T; ++k) {
        const long long r = i * a.T + k, N = a.N;
        if (e.E) e.E[e.t3n ? (long long)k * N + i : r] = NANV;
        if (e.L) for (int c = 0; c < 3; ++c) e.L[e.t3n ? (3LL * k + c) * N + i : 3 * r + c] = NANV;
        if (e.TT) for (int c = 0; c < 9; ++c) e.TT[e.t3n ? (9LL * k + c) * N + i : 9 * r + c] = NANV;
    }
}

template <bool EPI, class C, class A>
__device__ __forceinline__ void save_put_epi(const DevPot &P, const A &a, long long i, int k, double qx, double qy,
                                             double qz, double px, double py, double pz) {
    save_put(a, i, k, qx, qy, qz, px, py, pz);
    if constexpr (EPI) epilogue_put<C>(&P, &a, i, k, qx, qy, qz, px, py, pz);
}
template <bool EPI, class A>
__device__ __forceinline__ void save_finish(const A &a, long long i, int k) {
    if constexpr (EPI) epilogue_finish(&a, i, k);
    save_fill_nan(a, i, k);
}

struct FixedArgs {
    const double *q0, *p0, *ts;
    double *q, *p;
    int *status;
    long long N, n_steps;  // n_steps: trip count of the shared time grid (host-computed)
    long long sn, sk, sc;  // output strides (elements): particle, save, component
    double t0, t1, dt0;
    int T, hit_max_steps;
    int stage;  // saves go through the per-lane staging buffer (save_put)
    int stage_off;  // ... which starts this many doubles into dynamic shared memory (behind the wide force table)
    EpiOut epi;  // fused E / L / tidal-tensor outputs (EPI kernels); epi.nv = 6 otherwise
};

// Run-length form of the shared time grid (k_integrate_fixed_seg).  diffrax's grid t_{n+1} = fl(t_n + dt0) has a step
// h_n = t_{n+1} - t_n that is exactly representable and piecewise constant: while t stays inside one binade the sum
// rounds the same way every time, so h only changes where t crosses a power of two (and at the clipped last step).
// C1's 10^4 steps are ~30 runs.  The host, which walks the grid anyway for the trip count, passes the runs by value.
constexpr int FIXED_MAX_SEG = 120;
struct FixedSeg {
    int n_seg;
    long long cnt[FIXED_MAX_SEG];  // steps in the run
    double h[FIXED_MAX_SEG];       // their common step (in tau = dir * t)
};

static inline double clip_to_end_host(double tprev, double tnext, double t1) {
    return (tnext > t1 - 1e-10) ? t1 : tnext;
}

// Walk diffrax's ConstantStepSize grid exactly as the device does (same fp64 operations): trip count, whether
// max_steps cut it short, and (while `seg_ok`) its run-length form for k_integrate_fixed_seg.  seg_ok turns false when
// the grid needs more than FIXED_MAX_SEG runs or a step is not exactly representable (the general kernel handles both).
static void walk_time_grid(double t0, double t1, double dt0, long long max_steps, FixedSeg &sg, bool &seg_ok,
                           long long &n_steps, int &hit_max_steps) {
    const double dir = (t1 >= t0) ? 1.0 : -1.0;
    const double T0 = t0 * dir, T1 = t1 * dir, h0 = dt0 * dir;
    double tprev = T0, tnext = clip_to_end_host(T0, T0 + h0, T1), seg_t0 = 0.0;
    long long n = 0;
    int hit = 0;
    sg.n_seg = 0;
    while (tprev < T1) {
        if (max_steps >= 0 && n >= max_steps) { hit = 1; break; }
        ++n;
        if (seg_ok) {  // run-length encode h_n = tnext - tprev (exact); a run must also satisfy t_s + j h == t_j
            const double h = tnext - tprev;
            if (sg.n_seg > 0 && sg.h[sg.n_seg - 1] == h &&
                fma((double)(sg.cnt[sg.n_seg - 1] + 1), h, seg_t0) == tnext) {
                ++sg.cnt[sg.n_seg - 1];
            } else if (sg.n_seg < FIXED_MAX_SEG && tprev + h == tnext) {
                sg.h[sg.n_seg] = h; sg.cnt[sg.n_seg] = 1; seg_t0 = tprev; ++sg.n_seg;
            } else {
                seg_ok = false;
            }
        }
        tprev = tnext;
        tnext = clip_to_end_host(tprev, tprev + h0, T1);
    }
    n_steps = n;
    hit_max_steps = hit;
}

__device__ __forceinline__ double clip_to_end(double tprev, double tnext, double t1, bool keep) {
    // diffrax _clip_to_end (fp64 tolerance 1e-10)
    if (tnext > t1 - 1e-10) return keep ? t1 : tprev + 0.5 * (t1 - tprev);
    return tnext;
}

__device__ __forceinline__ bool finite3(double a, double b, double c) {
    return isfinite(a) && isfinite(b) && isfinite(c);
}
// six at once, branch-free on the integer pipe: the largest exponent field is not all ones.  (A chain of isfinite()
// compiles to one branch per component, and every branch target of the Dopri kernel's accept block then gets its own
// copy of the block's register moves: 145 moves per step in the round-2 profile.)
__device__ __forceinline__ bool finite6(double a, double b, double c, double d, double e, double f) {
    const unsigned M = 0x7ff00000u;
    const unsigned m = max(max(max((unsigned)__double2hiint(a) & M, (unsigned)__double2hiint(b) & M),
                               max((unsigned)__double2hiint(c) & M, (unsigned)__double2hiint(d) & M)),
                           max((unsigned)__double2hiint(e) & M, (unsigned)__double2hiint(f) & M));
    return m != M;
}

// SemiImplicitEuler state update y1 = y0 + f*dt as one FMA per component (GX_FUSED_UPDATE=1, default): six FP64
// instructions fewer per step (+4 % measured).  diffrax writes a multiply and an add; whether XLA contracts them is
// up to LLVM (fp-contract=fast in both of its backends), so neither form is "the" reference bit pattern -- the
// oracle keeps the two roundings, and the two forms differ by <= 0.5 ulp per step (parity bar: 1e-12 after 1e4
// steps; measured median 2e-14 either way).  GX_FUSED_UPDATE=0 restores __dmul_rn + __dadd_rn.
// The time grid (tnext = tprev + dt0 accumulated in fp64, last step clipped to t1) is identical for every
// particle; the host walks it once to get the trip count, so the device loop is a counted loop.
#ifndef GX_FIXED_MIN_BLOCKS
#define GX_FIXED_MIN_BLOCKS 1
#endif
#ifndef GX_FUSED_UPDATE
#define GX_FUSED_UPDATE 1
#endif
// Widest CTA of the kernels that hold the 132 KB table (one CTA per SM).  The run-length kernel is bound by latency --
// FP64 chains and scattered shared-memory loads in turn -- and gains from every warp it can get: 1024 threads at 64
// registers (MilkyWayPotential: no spill; MW2022 / Bovy / runtime composites: 8-32 bytes outside the hot loop) run 6.5 %
// / 1.2 % / 1.5 % faster than 640 threads at 76-90 registers; 768 and 832 threads measured no better than 640.  The
// step-by-step kernel (LeapfrogMidpoint, time-dependent parameters) would spill 40-64 bytes at 64 registers and stays
// at 640.
// 4: Estrin form of the table row in the Dopri right-hand side; 6: the same with the row fetched as soon as r^2 is known,
// before the disk terms (the call used to END on the loads' latency): Dopri8 MW 57.1 -> 53.1 ms, MW2022 68.5 -> 66.2, Bovy
// 58.7 -> 53.5, C2's shape 101.3 -> 98.9; no spill at the 168-register budget (Bovy: 48 bytes outside the stage loop);
// the same operations, the same bits.
#ifndef GX_DP_SPH_FORM
#define GX_DP_SPH_FORM 6
#endif
#ifndef GX_RHS_FACTORS
#define GX_RHS_FACTORS 1
#endif
// 3: Horner, 4: Estrin form of the table row in the fixed-step kernels
#ifndef GX_FIXED_SPH_FORM
#define GX_FIXED_SPH_FORM 3
#endif
#ifndef GX_FIXED_TAB_BLOCK
#define GX_FIXED_TAB_BLOCK 640
#endif
#ifndef GX_FIXED_SEG_BLOCK
#define GX_FIXED_SEG_BLOCK 1024
#endif
#ifndef GX_FIXED_CNT32
#define GX_FIXED_CNT32 1
#endif
#ifndef GX_FIXED_SEG
#define GX_FIXED_SEG 1
#endif
#ifndef GX_BASIC_MIX
#define GX_BASIC_MIX 0
#endif
// how a fixed-step table step evaluates its polynomial, for every batch size: 1 Horner (7 FP64, 7 deep), 2 Estrin (9, 3 deep)
#ifndef GX_SPH_FIXED_FORM
#define GX_SPH_FIXED_FORM 1
#endif
// CTAs of the kernels that hold the wide table (132 KB: one per SM) are as wide as the batch needs, up to 640 threads
// (384 with the save epilogue: its staging needs the room, and its kernels the registers)
template <class C, bool EPI = false, bool SEG = false>
__host__ __device__ constexpr int fixed_max_block() {
    return (sph_tab_fixed_ok<C>() || sph_mix_ok<C>()) ? (EPI ? 384 : (SEG ? GX_FIXED_SEG_BLOCK : GX_FIXED_TAB_BLOCK)) : 128;
}
template <class C, int SCHEME, bool FWD, bool EPI = false, bool SMALL = false>
__global__ void __launch_bounds__(fixed_max_block<C, EPI>(), GX_FIXED_MIN_BLOCKS) k_integrate_fixed(const __grid_constant__ DevPot P, const FixedArgs a) {
    // tables in shared memory: the composite's combined spherical table (MW2022, Bovy; every static model in small
    // batches, see k_integrate_fixed_seg), else the PowerLawCutoff / NFW tables of round 1 (GX_SPH_TABLE=0 builds)
    constexpr int SPHT = sph_tab_fixed_ok<C>() ? GX_FIXED_SPH_FORM : 0;  // 3: the wide format, at the start of dynamic shared memory
    constexpr bool STAGED = !SPHT && C::is_static && C::kPLC > 0;
    constexpr bool NFWT = !SPHT && nfw_tab_fixed_ok<C>();
    constexpr bool MIX = sph_mix_ok<C>() && SCHEME == GX_SCHEME_SEMI_IMPLICIT_EULER;  // (see k_integrate_fixed_seg)
    unsigned plc_base = 0, nfw_base = 0;
    if constexpr (STAGED) { plc_stage<C>(P); plc_base = plc_smem_base<C>(); }
    if constexpr (SPHT != 0 || MIX) nfw_base = sph_wide_stage(P);
    else nfw_base = nfw_stage<C, NFWT>(P);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    // integrate in tau = dir * t (diffrax flips the sign of time the same way for t1 < t0)
    const double T0 = FWD ? a.t0 : -a.t0, T1 = FWD ? a.t1 : -a.t1, h0 = FWD ? a.dt0 : -a.dt0;
    double qx = a.q0[3 * i], qy = a.q0[3 * i + 1], qz = a.q0[3 * i + 2];
    double px = a.p0[3 * i], py = a.p0[3 * i + 1], pz = a.p0[3 * i + 2];
    double mqx = qx, mqy = qy, mqz = qz, mpx = px, mpy = py, mpz = pz, tm = T0;  // LeapfrogMidpoint memory
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    int k = 0;
    auto load_ts = [&](int kk) { return (kk < a.T) ? (FWD ? __ldg(a.ts + kk) : -__ldg(a.ts + kk)) : INF; };
    double tsave = load_ts(k);
    while (tsave <= T0) {  // save times equal to t0 return y0
        save_put_epi<EPI, C>(P, a, i, k, qx, qy, qz, px, py, pz);
        ++k;
        tsave = load_ts(k);
    }
    double tprev = T0, tnext = clip_to_end(T0, T0 + h0, T1, true);
    for (long long n = 0; n < a.n_steps; ++n) {
        const double h = tnext - tprev;
        const double hs = FWD ? h : -h;  // signed step in physical time
        double nqx, nqy, nqz, npx, npy, npz, gx_, gy_, gz_;
        if (SCHEME == GX_SCHEME_SEMI_IMPLICIT_EULER) {
#if GX_FUSED_UPDATE
            nqx = fma(px, hs, qx);
            nqy = fma(py, hs, qy);
            nqz = fma(pz, hs, qz);
            if (C::is_static) {  // p1 = p0 - (fh h) x: the step is folded into the two scalar factors
                double fh, fv;
                if (MIX && sph_mix_table_step((unsigned long long)n)) gradient_factors<C, false, false, GX_FIXED_SPH_FORM>(P, nqx, nqy, nqz, fh, fv, 0u, nfw_base);
                else gradient_factors<C, STAGED, NFWT, SPHT>(P, nqx, nqy, nqz, fh, fv, plc_base, nfw_base);
                const double fhh = -fh * hs, fvh = -fv * hs;
                npx = fma(fhh, nqx, px);
                npy = fma(fhh, nqy, py);
                npz = fma(fvh, nqz, pz);
                gx_ = gy_ = gz_ = 0.0;
            } else {
                if (GX_BASIC_MIX && C::basic_tab && GX_SPH_MIX_PERIOD > 0 && P.n_mn <= 1 && !sph_mix_table_step((unsigned long long)n))
                    gradient<C, false, false, 0>(P, nqx, nqy, nqz, gx_, gy_, gz_);
                else gradient<C, STAGED, NFWT, SPHT>(P, nqx, nqy, nqz, gx_, gy_, gz_, FWD ? tprev : -tprev, nfw_base);
                npx = fma(-gx_, hs, px);
                npy = fma(-gy_, hs, py);
                npz = fma(-gz_, hs, pz);
            }
#else
            nqx = __dadd_rn(qx, __dmul_rn(px, hs));
            nqy = __dadd_rn(qy, __dmul_rn(py, hs));
            nqz = __dadd_rn(qz, __dmul_rn(pz, hs));
            gradient<C, STAGED, NFWT, SPHT>(P, nqx, nqy, nqz, gx_, gy_, gz_, FWD ? tprev : -tprev, nfw_base);
            npx = __dadd_rn(px, __dmul_rn(-gx_, hs));
            npy = __dadd_rn(py, __dmul_rn(-gy_, hs));
            npz = __dadd_rn(pz, __dmul_rn(-gz_, hs));
#endif
        } else {
            const double hm = tnext - tm;
            const double hh = FWD ? hm : -hm;
            gradient<C, STAGED, NFWT, SPHT>(P, qx, qy, qz, gx_, gy_, gz_, FWD ? tprev : -tprev, nfw_base);
            nqx = __dadd_rn(mqx, __dmul_rn(px, hh));
            nqy = __dadd_rn(mqy, __dmul_rn(py, hh));
            nqz = __dadd_rn(mqz, __dmul_rn(pz, hh));
            npx = __dadd_rn(mpx, __dmul_rn(-gx_, hh));
            npy = __dadd_rn(mpy, __dmul_rn(-gy_, hh));
            npz = __dadd_rn(mpz, __dmul_rn(-gz_, hh));
            mqx = qx; mqy = qy; mqz = qz; mpx = px; mpy = py; mpz = pz;
            tm = tprev;
        }
        if (tsave <= tnext) {  // LocalLinearInterpolation between (tprev, y) and (tnext, yn)
            do {
                const double th = (tsave - tprev) / (tnext - tprev);
                save_put_epi<EPI, C>(P, a, i, k, __dadd_rn(qx, __dmul_rn(th, __dsub_rn(nqx, qx))),
                         __dadd_rn(qy, __dmul_rn(th, __dsub_rn(nqy, qy))), __dadd_rn(qz, __dmul_rn(th, __dsub_rn(nqz, qz))),
                         __dadd_rn(px, __dmul_rn(th, __dsub_rn(npx, px))), __dadd_rn(py, __dmul_rn(th, __dsub_rn(npy, py))),
                         __dadd_rn(pz, __dmul_rn(th, __dsub_rn(npz, pz))));
                ++k;
                tsave = load_ts(k);
            } while (tsave <= tnext);
        }
        qx = nqx; qy = nqy; qz = nqz; px = npx; py = npy; pz = npz;
        tprev = tnext;
        tnext = clip_to_end(tprev, tprev + h0, T1, true);
    }
    int st = a.hit_max_steps ? GX_MAX_STEPS_REACHED : GX_OK;
    if (!(finite3(qx, qy, qz) && finite3(px, py, pz))) st = (st == GX_OK) ? GX_NONFINITE : st;
    save_finish<EPI>(a, i, k);
    if (a.status) a.status[i] = st;
}

// K2, run-length variant (static models, SemiImplicitEuler): the same arithmetic on the state as k_integrate_fixed,
// but no time arithmetic inside the hot loop.  Within a run the grid times are t_s + j h exactly (every sum is
// representable), so "which step contains the next save time" is decided once per save by an exact comparison
// (fma(j, h, t_s) is the grid time itself) instead of a DADD + 2 DSETP + FSEL per step, the step is a uniform
// constant-bank operand, and the state is updated in place.  Saves are interpolated exactly as in k_integrate_fixed
// (same theta, same operations): results are bit-identical to it.
// SMALL: the instantiation for launches of few warps per scheduler (CTAs of up to 512 threads: C1's 10^4 particles are
// one warp on half of the schedulers), which are bound by the dependent chain of ONE step: it fetches the table row as
// soon as r^2 is known, before the disk terms (gradient_factors<..., 5>).  The SAME arithmetic, bit for bit -- a
// particle's result must not depend on the batch it travels in (an earlier variant that took the table on every step in
// Estrin form for small batches was measured, 1.94 ms, and not shipped for that reason).
template <class C, bool FWD, bool SMALL = false, bool EPI = false>
__global__ void __launch_bounds__((SMALL && fixed_max_block<C, EPI, true>() > 512) ? 512 : fixed_max_block<C, EPI, true>(), GX_FIXED_MIN_BLOCKS)
k_integrate_fixed_seg(const __grid_constant__ DevPot P, const FixedArgs a, const __grid_constant__ FixedSeg sg) {
    // (runtime composites come here too when none of their parameters depends on time: same loop, gradient<C>())
    // tables in shared memory: the composite's combined spherical table (MW2022, Bovy), else the PowerLawCutoff / NFW
    // tables of round 1 (GX_SPH_TABLE=0 builds)
    constexpr int SPHF = SMALL ? 5 : GX_FIXED_SPH_FORM;  // Horner; SMALL: the row fetched early
    constexpr int SPHT = sph_tab_fixed_ok<C>() ? SPHF : 0;  // 3: the wide format, at the start of dynamic shared memory
    constexpr bool STAGED = !SPHT && C::is_static && C::kPLC > 0;
    constexpr bool NFWT = !SPHT && nfw_tab_fixed_ok<C>();
    // MIX (MilkyWayPotential in large batches): the closed forms cost issue slots (169 per warp-step, the kernel's whole
    // budget) and leave the shared-memory port idle; the table lookup costs ~40 port cycles and 50 slots fewer.  Steps
    // therefore ALTERNATE between the two evaluations of the same force, by the global index of the step (so the
    // trajectory does not depend on where the save times fall, and the general kernel takes the same sequence): both
    // resources work at the same time.
    constexpr bool MIX = sph_mix_ok<C>();
    unsigned plc_base = 0, nfw_base = 0;
    if constexpr (STAGED) { plc_stage<C>(P); plc_base = plc_smem_base<C>(); }
    if constexpr (SPHT != 0 || MIX) nfw_base = sph_wide_stage(P);
    else nfw_base = nfw_stage<C, NFWT>(P);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    unsigned gstep = 0;  // index of the step in the whole grid (MIX)
    // (a runtime composite with one disk is MilkyWayPotential's case, but its closed forms are loops with runtime trip
    //  counts: alternating measured 2.01e11 against 2.11e11 on the table alone -- GX_BASIC_MIX=1 builds only)
    const bool mixb = GX_BASIC_MIX && C::basic_tab && GX_SPH_MIX_PERIOD > 0 && P.n_mn <= 1;
    const double T0 = FWD ? a.t0 : -a.t0;
    double qx = a.q0[3 * i], qy = a.q0[3 * i + 1], qz = a.q0[3 * i + 2];
    double px = a.p0[3 * i], py = a.p0[3 * i + 1], pz = a.p0[3 * i + 2];
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    int k = 0;
    auto load_ts = [&](int kk) { return (kk < a.T) ? (FWD ? __ldg(a.ts + kk) : -__ldg(a.ts + kk)) : INF; };
    double tsave = load_ts(k);
    while (tsave <= T0) {  // save times equal to t0 return y0
        save_put_epi<EPI, C>(P, a, i, k, qx, qy, qz, px, py, pz);
        ++k;
        tsave = load_ts(k);
    }
    double tprev = T0;
    for (int s = 0; s < sg.n_seg; ++s) {
        const double h = sg.h[s];
        const double hs = FWD ? h : -h;  // signed step in physical time
        long long cnt = sg.cnt[s];
        while (cnt > 0) {
            // m = number of leading steps of this run that end before the next save time
            long long m = cnt;
            if (tsave < INF) {
                const double x = (tsave - tprev) / h;
                long long j = (x >= (double)cnt) ? cnt : ((x > 0.0) ? (long long)x : 0);
                while (j < cnt && fma((double)(j + 1), h, tprev) < tsave) ++j;
                while (j > 0 && !(fma((double)j, h, tprev) < tsave)) --j;
                m = j;
            }
#if GX_FIXED_CNT32
            for (long long left = m; left > 0;) {  // 32-bit counter in the hot loop (two instructions fewer per step)
                const int mm = left > (1LL << 30) ? (1 << 30) : (int)left;
                left -= mm;
#pragma unroll 1
                for (int n = mm; n > 0; --n) {
#else
            {
                for (long long n = 0; n < m; ++n) {
#endif
                    qx = fma(px, hs, qx);
                    qy = fma(py, hs, qy);
                    qz = fma(pz, hs, qz);
                    if constexpr (C::is_static) {
                        double fh, fv;
                        if (MIX && sph_mix_table_step(gstep)) gradient_factors<C, false, false, SPHF>(P, qx, qy, qz, fh, fv, 0u, nfw_base);
                        else gradient_factors<C, STAGED, NFWT, SPHT>(P, qx, qy, qz, fh, fv, plc_base, nfw_base);
                        if constexpr (MIX) ++gstep;
                        const double fhh = -fh * hs, fvh = -fv * hs;
                        px = fma(fhh, qx, px);
                        py = fma(fhh, qy, py);
                        pz = fma(fvh, qz, pz);
                    } else {
                        double g0, g1, g2;
                        if (mixb && !sph_mix_table_step(gstep)) gradient<C, false, false, 0>(P, qx, qy, qz, g0, g1, g2);
                        else gradient<C, false, false, SPHT>(P, qx, qy, qz, g0, g1, g2, 0.0, nfw_base);
                        if constexpr (C::basic_tab) ++gstep;
                        px = fma(-g0, hs, px);
                        py = fma(-g1, hs, py);
                        pz = fma(-g2, hs, pz);
                    }
                }
            }
            tprev = fma((double)m, h, tprev);
            cnt -= m;
            if (cnt > 0) {  // the step that contains (at least) one save time
                const double tnext = tprev + h;
                const double nqx = fma(px, hs, qx), nqy = fma(py, hs, qy), nqz = fma(pz, hs, qz);
                double npx, npy, npz;
                if constexpr (C::is_static) {
                    double fh, fv;
                    if (MIX && sph_mix_table_step(gstep)) gradient_factors<C, false, false, SPHF>(P, nqx, nqy, nqz, fh, fv, 0u, nfw_base);
                    else gradient_factors<C, STAGED, NFWT, SPHT>(P, nqx, nqy, nqz, fh, fv, plc_base, nfw_base);
                    if constexpr (MIX) ++gstep;
                    const double fhh = -fh * hs, fvh = -fv * hs;
                    npx = fma(fhh, nqx, px); npy = fma(fhh, nqy, py); npz = fma(fvh, nqz, pz);
                } else {
                    double g0, g1, g2;
                    if (mixb && !sph_mix_table_step(gstep)) gradient<C, false, false, 0>(P, nqx, nqy, nqz, g0, g1, g2);
                    else gradient<C, false, false, SPHT>(P, nqx, nqy, nqz, g0, g1, g2, 0.0, nfw_base);
                    if constexpr (C::basic_tab) ++gstep;
                    npx = fma(-g0, hs, px); npy = fma(-g1, hs, py); npz = fma(-g2, hs, pz);
                }
                while (tsave <= tnext) {  // LocalLinearInterpolation between (tprev, y) and (tnext, yn)
                    const double th = (tsave - tprev) / (tnext - tprev);
                    save_put_epi<EPI, C>(P, a, i, k, __dadd_rn(qx, __dmul_rn(th, __dsub_rn(nqx, qx))),
                             __dadd_rn(qy, __dmul_rn(th, __dsub_rn(nqy, qy))), __dadd_rn(qz, __dmul_rn(th, __dsub_rn(nqz, qz))),
                             __dadd_rn(px, __dmul_rn(th, __dsub_rn(npx, px))), __dadd_rn(py, __dmul_rn(th, __dsub_rn(npy, py))),
                             __dadd_rn(pz, __dmul_rn(th, __dsub_rn(npz, pz))));
                    ++k;
                    tsave = load_ts(k);
                }
                qx = nqx; qy = nqy; qz = nqz; px = npx; py = npy; pz = npz;
                tprev = tnext;
                --cnt;
            }
        }
    }
    int st = a.hit_max_steps ? GX_MAX_STEPS_REACHED : GX_OK;
    if (!(finite3(qx, qy, qz) && finite3(px, py, pz))) st = (st == GX_OK) ? GX_NONFINITE : st;
    save_finish<EPI>(a, i, k);
    if (a.status) a.status[i] = st;
}

// ================================================================================================
// K3 Dopri8 + PID, persistent warps with a per-lane work queue
// ================================================================================================

struct Dp8Args {
    const double *q0, *p0, *t0v, *ts;
    const int *order;
    double *q, *p;
    int *status, *n_acc, *n_tot;
    unsigned long long *ticket;
    double *rec;   // optional step records for the parallel dense output (single-orbit solves)
    int *n_rec;
    int rec_cap;
    long long N, max_steps;
    long long sn, sk, sc;
    double t0s, t1;
    double rtol, atol, pcoeff, icoeff, dcoeff, safety, factormin, factormax, dtmin, dtmax, dt0;
    int T;
    int stage;  // saves go through the per-lane staging buffer (save_put)
    int stage_off;  // ... which starts this many doubles into dynamic shared memory (behind the wide force table)
    EpiOut epi;  // fused E / L / tidal-tensor outputs (EPI kernels); epi.nv = 6 otherwise
};

template <class TB>
__host__ __device__ constexpr bool dq_row_nz(int l) {
    for (int m = 0; m < 6; ++m)
        if (TB::DQ_NZ(l, m)) return true;
    return false;
}
template <class TB>
__host__ __device__ constexpr bool db_row_nz(int l) {
    for (int m = 0; m < 6; ++m)
        if (TB::DB_NZ(l, m)) return true;
    return false;
}

// Dense-output weight of stage l at theta: theta * sum_m D(l, m) theta^m.  For every row but the first the constant term
// D(l, 0) vanishes (the interpolant's slope at theta = 0 is k_1 alone), so those rows are theta^2 * (degree-4 Horner):
// one FP64 instruction and one constant load fewer per row, 20 rows per save.  Shared by the in-kernel SaveAt path and
// k_dense_eval so that both give the same bits.
template <class TB, bool Q>
__device__ __forceinline__ double dense_weight(int l, double th, double th2) {
    const bool c0 = Q ? TB::DQ_NZ(l, 0) : TB::DB_NZ(l, 0);
    double w = Q ? TB::DQ(l, 5) : TB::DB(l, 5);
    if (c0) {
#pragma unroll
        for (int m = 4; m >= 0; --m) w = fma(w, th, Q ? TB::DQ(l, m) : TB::DB(l, m));
        return w * th;
    }
#pragma unroll
    for (int m = 4; m >= 1; --m) w = fma(w, th, Q ? TB::DQ(l, m) : TB::DB(l, m));
    return w * th2;
}

__device__ __forceinline__ double rms6(double a, double b, double c, double d, double e, double f) {
    double s = a * a;
    s = fma(b, b, s); s = fma(c, c, s); s = fma(d, d, s); s = fma(e, e, s); s = fma(f, f, s);
    return sqrt(s * (1.0 / 6.0));
}

template <class C>
__device__ __forceinline__ void accel(const DevPot &P, double x, double y, double z, double &ax, double &ay,
                                      double &az) {
    double g0, g1, g2;
    gradient<C>(P, x, y, z, g0, g1, g2);
    ax = -g0; ay = -g1; az = -g2;
}

// Out-of-line right-hand side for the Dopri8 kernel: 13 inlined copies of the gradient make the step body
// ~100 KB of SASS, far beyond the 32 KB instruction cache (ncu: "no_instruction" was the top stall), so the callee
// takes no pointer argument and finds the potential parameters itself (constant or shared memory, next paragraph).
// Where the out-of-line RHS finds the potential: template parameter IMG.
//   IMG = true   a __constant__ image (c_pot_dp8): inside the callee the parameters are constant-bank operands of the
//                FP64 instructions instead of ~30 shared-memory loads per call (measured: -6.5 % on the Dopri8 step).
//                One image per device, shared by every launch: use_const_image() below decides, without ever blocking,
//                whether a launch may use it.
//   IMG = false  the kernel copies its own __grid_constant__ parameter into shared memory once per CTA: no state
//                outside the launch.  Used whenever the image is busy with another potential on another stream, and
//                always under CUDA-graph capture.
__constant__ DevPot c_pot_dp8;
// (IMG = false kernels keep their copy of the potential in static shared memory; the force table is in dynamic memory)
template <class C>
__device__ __forceinline__ DevPot *pot_smem() {
    __shared__ DevPot sP;
    return &sP;
}
template <class C, bool IMG>
__device__ __forceinline__ const DevPot &rhs_pot() {
    if constexpr (IMG) return c_pot_dp8;
    else return *pot_smem<C>();
}
// The callees return the acceleration BY VALUE (three doubles in registers): with reference parameters the results
// went through the caller's local-memory frame (3 STL + 3 LDL + the generic-to-local address arithmetic per call).
struct Acc3 { double x, y, z; };
template <class C, bool IMG>
__device__ __noinline__ Acc3 accel_call_static(double x, double y, double z) {
    // (handing the table's shared-window address down in a register instead of re-deriving it here -- S2R + LEA + ISETP
    // + MOV per call -- was measured: the extra live register costs the Dopri8 kernel 40 bytes of spills)
    double g0, g1, g2;
    unsigned tab_base = 0;  // the kernel prologue staged the force table (sph_stage / nfw_stage)
    constexpr int SPHM = sph_tab_ok<C>() ? GX_DP_SPH_FORM : 0;  // the wide table at the start of dynamic shared memory, Estrin form
    if constexpr (SPHM != 0) tab_base = (unsigned)__cvta_generic_to_shared(dyn_smem());
    else if constexpr (nfw_tab_ok<C>()) tab_base = (unsigned)__cvta_generic_to_shared(nfw_smem<C>());
    gradient<C, (!SPHM && C::is_static && C::kPLC > 0), (!SPHM && nfw_tab_ok<C>()), SPHM>(rhs_pot<C, IMG>(), x, y, z, g0, g1,
                                                                                          g2, 0.0, tab_base);
    return Acc3{-g0, -g1, -g2};
}
// The same callee returning the two scalar force factors of an axisymmetric model, a = -(f_h x, f_h y, f_v z): it no
// longer needs q at its end (no copies of the arguments out of the way of its table loads), and the caller forms the
// products straight into the stage's registers instead of moving three returned doubles there.
struct Fac2 { double h, v; };
template <class C, bool IMG>
__device__ __noinline__ Fac2 accel_fac_static(double x, double y, double z) {
    unsigned tab_base = 0;
    constexpr int SPHM = sph_tab_ok<C>() ? GX_DP_SPH_FORM : 0;
    if constexpr (SPHM != 0) tab_base = (unsigned)__cvta_generic_to_shared(dyn_smem());
    else if constexpr (nfw_tab_ok<C>()) tab_base = (unsigned)__cvta_generic_to_shared(nfw_smem<C>());
    double fh, fv;
    gradient_factors<C, (!SPHM && C::is_static && C::kPLC > 0), (!SPHM && nfw_tab_ok<C>()), SPHM>(rhs_pot<C, IMG>(), x, y, z, fh,
                                                                                                  fv, 0, tab_base);
    return Fac2{fh, fv};
}
// runtime composites may be time dependent (LinearParameter): the callee also receives the physical time
template <class C, bool IMG>
__device__ __noinline__ Acc3 accel_call_timed(double t, double x, double y, double z) {
    double g0, g1, g2;
    gradient<C, false>(rhs_pot<C, IMG>(), x, y, z, g0, g1, g2, t);
    return Acc3{-g0, -g1, -g2};
}
template <class C, bool IMG>
__device__ __forceinline__ void accel_call(double x, double y, double z, double &ax, double &ay, double &az, double t) {
    if constexpr (GX_RHS_FACTORS && (C::is_static || C::basic_only)) {
        const Fac2 f = accel_fac_static<C, IMG>(x, y, z);
        ax = -__dmul_rn(f.h, x); ay = -__dmul_rn(f.h, y); az = -__dmul_rn(f.v, z);  // (the products gradient<>() forms)
        return;
    }
    Acc3 a;
    if constexpr (C::is_static || C::basic_only) a = accel_call_static<C, IMG>(x, y, z);
    else a = accel_call_timed<C, IMG>(t, x, y, z);
    ax = a.x; ay = a.y; az = a.z;
}

// Hairer-Norsett-Wanner initial step as restated by diffrax (_select_initial_step); inv_order = 1 / error order.
// Works in tau = dir*t: f = dir * (p, a).
template <class C, bool IMG>
__device__ double select_initial_step(const DevPot &P, double dir, double tau0, const double y[6], const double a0[3],
                                      double rtol, double atol, double inv_order) {
    double f0[6] = {y[3] * dir, y[4] * dir, y[5] * dir, a0[0] * dir, a0[1] * dir, a0[2] * dir};
    double sc[6], v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sc[i] = atol + fabs(y[i]) * rtol;
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = y[i] / sc[i];
    double d0 = rms6(v[0], v[1], v[2], v[3], v[4], v[5]);
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = f0[i] / sc[i];
    double d1 = rms6(v[0], v[1], v[2], v[3], v[4], v[5]);
    bool cond = (d0 < 1e-5) || (d1 < 1e-5);
    double h0 = cond ? 1e-6 : 0.01 * (d0 / d1);
    double y1[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y1[i] = y[i] + h0 * f0[i];
    double a1x, a1y, a1z;
    accel_call<C, IMG>(y1[0], y1[1], y1[2], a1x, a1y, a1z, dir * (tau0 + h0));
    double f1[6] = {y1[3] * dir, y1[4] * dir, y1[5] * dir, a1x * dir, a1y * dir, a1z * dir};
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = (f1[i] - f0[i]) / sc[i];
    double d2 = rms6(v[0], v[1], v[2], v[3], v[4], v[5]) / h0;
    double maxd = fmax(d1, d2);
    double h1 = (maxd <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / maxd, inv_order);
    return fmin(100.0 * h0, h1);
}

constexpr int REC_DOUBLES = GX_DENSE_RECORD_DOUBLES;  // tprev, tnext, hd, q0[3], p0[3], a[14][3]

#ifndef GX_DP8_MIN_BLOCKS
#define GX_DP8_MIN_BLOCKS 3
#endif
// The 14 x 3 stage accelerations are thread-private arrays; because the right-hand side is an out-of-line call
// they live in the thread's local-memory frame (L1-resident, 336 B/thread) rather than in registers.  Measured
// alternatives on B200 (MW2022, rtol = atol = 1e-10, 3e5 particles): everything inlined with the stages in
// registers 161 ms (I-cache bound: 105 KB of SASS); stages in shared memory 151-205 ms; this version 121 ms.
// TB = TabDp8 (diffrax.Dopri8) or TabDp5 (diffrax.Dopri5): same kernel, tableau resolved at compile time.
#ifndef GX_DENSE_COEF
#define GX_DENSE_COEF 0
#endif
#ifndef GX_DP8_PIPELINE
#define GX_DP8_PIPELINE 1
#endif
template <class C, class TB, bool IMG, bool EPI = false>
__global__ void __launch_bounds__((sph_tab_ok<C>() ? 384 : 128), (sph_tab_ok<C>() ? 1 : GX_DP8_MIN_BLOCKS))
k_integrate_dopri8(const __grid_constant__ DevPot P, const Dp8Args a) {
    constexpr int NS = TB::NS;
    if constexpr (!IMG) {  // stage the potential parameters in shared memory for accel_call()
        const double *src = reinterpret_cast<const double *>(&P);
        double *dst = reinterpret_cast<double *>(pot_smem<C>());
        for (int w = threadIdx.x; w < (int)(sizeof(DevPot) / sizeof(double)); w += blockDim.x) dst[w] = src[w];
        __syncthreads();
    }
    if constexpr (sph_tab_ok<C>()) {
        (void)sph_wide_stage(P);  // the combined spherical table (132 KB, one CTA of 384 threads per SM), read by accel_call()
    } else {
        plc_stage<C>(P);  // (GX_SPH_TABLE=0) the PowerLawCutoff table
        (void)nfw_stage<C, nfw_tab_ok<C>()>(P);  // ... and the NFW force table
    }
    const unsigned FULL = 0xffffffffu;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const bool simple_i = (TB::ORDER == 8) && (a.pcoeff == 0.0 && a.dcoeff == 0.0 && a.icoeff == 1.0);
    bool have = false, exhausted = false;
    long long idx = 0;
    double q0x = 0, q0y = 0, q0z = 0, p0x = 0, p0y = 0, p0z = 0;  // state at tprev
    double fsx = 0.0, fsy = 0.0, fsz = 0.0;  // acceleration at (tprev, q0): the FSAL value, the only stage carried over
#define AX(l) rx[l]
#define AY(l) ry[l]
#define AZ(l) rz_[l]
    double dir = 1.0, T1 = 0, tprev = 0, tnext = 0, tsave = INF;
    double prev_inv = 1.0, prev_prev_inv = 1.0;
    bool at_dtmin = false;
    int k = 0, nacc = 0, ntot = 0, st = GX_OK;

    // Every lane passes the same points of the loop in every iteration (no `continue`): the warp votes below need it.
    //   1. a lane whose particle is finished (or failed) writes its tail and frees itself;
    //   2. free lanes pull the next particle from the global ticket;
    //   3. lanes with a particle attempt one step and run the controller (a rejected lane simply retries next time);
    //   4. SaveAt: all lanes whose accepted step contains a save time evaluate the dense output TOGETHER, once per
    //      round of a warp-uniform loop (__any_sync) -- left to the compiler's reconvergence, the same code ran
    //      with 6 of 32 lanes active and the dense output was a quarter of C2 (ncu, round 2);
    //   5. accepted lanes advance.
    for (;;) {
        // ---------------- finished (or failed) particle: write remaining saves, counters, free the lane
        if (have && (!(tprev < T1) || st != GX_OK || (a.max_steps >= 0 && ntot >= a.max_steps))) {
            if (tprev < T1 && st == GX_OK) st = GX_MAX_STEPS_REACHED;
            save_finish<EPI>(a, idx, k);  // (writes out what is still staged; NaN for saves never reached)
            k = a.T;
            if (a.status) a.status[idx] = st;
            if (a.n_acc) a.n_acc[idx] = nacc;
            if (a.n_tot) a.n_tot[idx] = ntot;
            if (a.rec && a.n_rec) *a.n_rec = nacc < a.rec_cap ? nacc : a.rec_cap;
            have = false;
        }
        // ---------------- refill idle lanes from the global ticket counter
        if (!have && !exhausted) {
            unsigned long long tk = atomicAdd(a.ticket, 1ULL);
            if (tk >= (unsigned long long)a.N) {
                exhausted = true;
            } else {
                idx = a.order ? (long long)a.order[tk] : (long long)tk;
                const double t0 = a.t0v ? a.t0v[idx] : a.t0s;
                dir = (a.t1 >= t0) ? 1.0 : -1.0;
                const double T0 = t0 * dir;
                T1 = a.t1 * dir;
                q0x = a.q0[3 * idx]; q0y = a.q0[3 * idx + 1]; q0z = a.q0[3 * idx + 2];
                p0x = a.p0[3 * idx]; p0y = a.p0[3 * idx + 1]; p0z = a.p0[3 * idx + 2];
                k = 0; nacc = 0; ntot = 0; st = GX_OK;
                prev_inv = prev_prev_inv = 1.0;
                at_dtmin = false;
                tsave = (k < a.T) ? __ldg(a.ts + k) * dir : INF;
                while (tsave <= T0) {
                    save_put_epi<EPI, C>(P, a, idx, k, q0x, q0y, q0z, p0x, p0y, p0z);
                    ++k;
                    tsave = (k < a.T) ? __ldg(a.ts + k) * dir : INF;
                }
                accel_call<C, IMG>(q0x, q0y, q0z, fsx, fsy, fsz, dir * T0);
                // PIDController.init: heuristic when dt0 is None (exponent 1/(error_order + 1), Hairer II.4 -- the
                // choice that reproduces the reference's 8-digit OrbitSolver doctests), then clamp to [dtmin, dtmax]
                double h;
                if (a.dt0 > 0.0) {
                    h = a.dt0;
                } else {
                    const double y[6] = {q0x, q0y, q0z, p0x, p0y, p0z};
                    const double a0[3] = {fsx, fsy, fsz};
                    h = select_initial_step<C, IMG>(P, dir, T0, y, a0, a.rtol, a.atol, 1.0 / (TB::ORDER + 1));
                }
                if (a.dtmax > 0.0) h = fmin(h, a.dtmax);
                if (a.dtmin > 0.0) { at_dtmin = h <= a.dtmin; h = fmax(h, a.dtmin); }
                tprev = T0;
                tnext = clip_to_end(T0, T0 + h, T1, true);
                have = true;
            }
        }
        if (__all_sync(FULL, !have)) break;
        const bool run = have && (tprev < T1) && st == GX_OK && !(a.max_steps >= 0 && ntot >= a.max_steps);
        const double h = tnext - tprev;
        const double hd = h * dir;  // signed step in physical time
        const double hd2 = hd * hd;
        bool keep = false;
        double dt = 0.0, inv = 1.0, q1x = 0, q1y = 0, q1z = 0, p1x = 0, p1y = 0, p1z = 0;
        // stage accelerations of THIS attempt (dead at the end of the iteration: declared here so that nothing but the
        // FSAL value is loop-carried and the register allocator need not keep 42 doubles in place across the back edge)
        double rx[NS], ry[NS], rz_[NS];
        AX(0) = fsx; AY(0) = fsy; AZ(0) = fsz;
        if (run) {
        // ---------------- one attempted step of size h from (tprev, y0)
        double sx = 0, sy = 0, sz = 0;
#if GX_DP8_PIPELINE
        // Software-pipelined across the out-of-line call: the part of stage i+1's sum that does not need a_i is formed
        // BEFORE the call that computes a_i (ptxas does not move work across a call), so it runs in the shadow of the
        // callee's first instructions and only one FMA per component is left between two calls.
        // (n = q0 + CN[i] hd p0 + hd^2 sum_{l < i-1} AA(i, l) a_l: everything of the coming stage point but a_{i-1}'s term)
        double nx = fma(TB::CN(1) * hd, p0x, q0x), ny = fma(TB::CN(1) * hd, p0y, q0y), nz = fma(TB::CN(1) * hd, p0z, q0z);
#pragma unroll
        for (int i = 1; i < NS; ++i) {
            double xi = nx, yi = ny, zi = nz;
            if (TB::AA_NZ(i, i - 1)) {
                const double c = hd2 * TB::AA(i, i - 1);
                xi = fma(c, AX(i - 1), xi);
                yi = fma(c, AY(i - 1), yi);
                zi = fma(c, AZ(i - 1), zi);
            }
            if (i + 1 < NS) {
                double tx = 0, ty = 0, tz = 0;
#pragma unroll
                for (int l = 0; l < i; ++l) {
                    if (TB::AA_NZ(i + 1, l)) {
                        tx = fma(TB::AA(i + 1, l), AX(l), tx);
                        ty = fma(TB::AA(i + 1, l), AY(l), ty);
                        tz = fma(TB::AA(i + 1, l), AZ(l), tz);
                    }
                }
                const double chn = TB::CN(i + 1) * hd;
                nx = fma(hd2, tx, fma(chn, p0x, q0x));
                ny = fma(hd2, ty, fma(chn, p0y, q0y));
                nz = fma(hd2, tz, fma(chn, p0z, q0z));
            }
            { double t0_, t1_, t2_; accel_call<C, IMG>(xi, yi, zi, t0_, t1_, t2_, fma(TB::CN(i), hd, dir * tprev)); AX(i) = t0_; AY(i) = t1_; AZ(i) = t2_; }
            if (i == NS - 1) { sx = xi; sy = yi; sz = zi; }  // FSAL: the last stage sits at q1
        }
#else
#pragma unroll
        for (int i = 1; i < NS; ++i) {
            // q_i = q0 + CN[i] hd p0 + hd^2 sum_{l<i} TB::AA(i, l) a_l
            sx = 0; sy = 0; sz = 0;
#pragma unroll
            for (int l = 0; l < i; ++l) {
                if (TB::AA_NZ(i, l)) {
                    sx = fma(TB::AA(i, l), AX(l), sx);
                    sy = fma(TB::AA(i, l), AY(l), sy);
                    sz = fma(TB::AA(i, l), AZ(l), sz);
                }
            }
            const double ch = TB::CN(i) * hd;
            const double xi = fma(hd2, sx, fma(ch, p0x, q0x));
            const double yi = fma(hd2, sy, fma(ch, p0y, q0y));
            const double zi = fma(hd2, sz, fma(ch, p0z, q0z));
            { double t0_, t1_, t2_; accel_call<C, IMG>(xi, yi, zi, t0_, t1_, t2_, fma(TB::CN(i), hd, dir * tprev)); AX(i) = t0_; AY(i) = t1_; AZ(i) = t2_; }
            if (i == NS - 1) { sx = xi; sy = yi; sz = zi; }  // FSAL: the last stage sits at q1
        }
#endif
        q1x = sx; q1y = sy; q1z = sz;
        double bx = 0, by = 0, bz = 0, epx = 0, epy = 0, epz = 0, eqx = 0, eqy = 0, eqz = 0;
#pragma unroll
        for (int l = 0; l < NS; ++l) {
            if (TB::B_NZ(l)) { bx = fma(TB::B(l), AX(l), bx); by = fma(TB::B(l), AY(l), by); bz = fma(TB::B(l), AZ(l), bz); }
            if (TB::E_NZ(l)) { epx = fma(TB::E(l), AX(l), epx); epy = fma(TB::E(l), AY(l), epy); epz = fma(TB::E(l), AZ(l), epz); }
            if (TB::EA_NZ(l)) { eqx = fma(TB::EA(l), AX(l), eqx); eqy = fma(TB::EA(l), AY(l), eqy); eqz = fma(TB::EA(l), AZ(l), eqz); }
        }
        p1x = fma(hd, bx, p0x); p1y = fma(hd, by, p0y); p1z = fma(hd, bz, p0z);
        ++ntot;

        // ---------------- PID controller (diffrax PIDController.adapt_step_size)
        const bool nan1 = isnan(q1x) || isnan(q1y) || isnan(q1z) || isnan(p1x) || isnan(p1y) || isnan(p1z);
        double e0 = hd2 * eqx, e1 = hd2 * eqy, e2 = hd2 * eqz, e3 = hd * epx, e4 = hd * epy, e5 = hd * epz;
        e0 *= rcp_fast(a.atol + fmax(fabs(q0x), fabs(nan1 ? q0x : q1x)) * a.rtol);
        e1 *= rcp_fast(a.atol + fmax(fabs(q0y), fabs(nan1 ? q0y : q1y)) * a.rtol);
        e2 *= rcp_fast(a.atol + fmax(fabs(q0z), fabs(nan1 ? q0z : q1z)) * a.rtol);
        e3 *= rcp_fast(a.atol + fmax(fabs(p0x), fabs(nan1 ? p0x : p1x)) * a.rtol);
        e4 *= rcp_fast(a.atol + fmax(fabs(p0y), fabs(nan1 ? p0y : p1y)) * a.rtol);
        e5 *= rcp_fast(a.atol + fmax(fabs(p0z), fabs(nan1 ? p0z : p1z)) * a.rtol);
        double ms = e0 * e0;  // mean square of the scaled error; serr = sqrt(ms)
        ms = fma(e1, e1, ms); ms = fma(e2, e2, ms); ms = fma(e3, e3, ms); ms = fma(e4, e4, ms); ms = fma(e5, e5, ms);
        ms *= (1.0 / 6.0);
        const bool bad = !(ms <= 1.7976931348623157e308);  // NaN or inf error estimate -> reject, shrink
        keep = !bad && ms < 1.0;
        if (a.dtmin > 0.0) keep = keep || at_dtmin;
        double factor;
        if (simple_i) {
            // I-controller (diffrax default pcoeff = dcoeff = 0, icoeff = 1):
            // factor = safety * serr^(-1/8) = safety * ms^(-1/16), by one rsqrt and three square roots
            double y = rsqrt_fast(ms);             // ms^(-1/2)
            y = y * rsqrt_fast(y);                 // ms^(-1/4)
            y = y * rsqrt_fast(y);                 // ms^(-1/8)
            y = y * rsqrt_fast(y);                 // ms^(-1/16)
            factor = bad ? 0.0 : ((ms > 0.0) ? a.safety * y : a.factormax);
        } else {
            const double serr = bad ? INF : sqrt(ms);
            inv = 1.0 / serr;
            const double c1 = (a.icoeff + a.pcoeff + a.dcoeff) * (1.0 / TB::ORDER);
            const double c2 = -(a.pcoeff + 2.0 * a.dcoeff) * (1.0 / TB::ORDER);
            const double c3 = a.dcoeff * (1.0 / TB::ORDER);
            factor = a.safety;
            if (c1 != 0.0) factor *= pow(inv, c1);
            if (c2 != 0.0) factor *= pow(prev_inv, c2);
            if (c3 != 0.0) factor *= pow(prev_prev_inv, c3);
        }
        const double fmin_ = keep ? 1.0 : a.factormin;
        factor = fmin(fmax(factor, fmin_), a.factormax);
        dt = h * factor;
        if (inv == 0.0 || isinf(inv)) { inv = 1.0; prev_inv = 1.0; }
        if (a.dtmax > 0.0) dt = fmin(dt, a.dtmax);
        if (a.dtmin > 0.0) { at_dtmin = dt <= a.dtmin; dt = fmax(dt, a.dtmin); }

        if (keep) {
            if (a.rec) {  // record the accepted step; k_dense_eval turns the records into saves in parallel
                if (nacc < a.rec_cap) {
                    double *r = a.rec + (long long)nacc * REC_DOUBLES;
                    r[0] = tprev; r[1] = tnext; r[2] = hd;
                    r[3] = q0x; r[4] = q0y; r[5] = q0z; r[6] = p0x; r[7] = p0y; r[8] = p0z;
#pragma unroll
                    for (int l = 0; l < NS; ++l) { r[9 + 3 * l] = AX(l); r[10 + 3 * l] = AY(l); r[11 + 3 * l] = AZ(l); }
                } else {
                    st = GX_MAX_STEPS_REACHED;  // record buffer exhausted
                }
            }
        }
        }  // if (run)

        // ---------------- SaveAt(ts): degree-6 continuous extension on the accepted step, warp-converged
        bool want = keep && (tsave <= tnext);
#if GX_DENSE_COEF
        // Coefficient form.  A save evaluates q(theta) = q0 + theta hd p0 + hd^2 sum_l wq_l(theta) a_l (and p likewise)
        // with 26 weight polynomials w_l(theta) = theta sum_m D(l, m) theta^m: 210 FP64 + ~100 constant loads per save,
        // executed once per round of the warp-uniform loop below by the ~5 lanes that want a save -- and the warp runs
        // 2.65 rounds per step at C2's 1000 saves (ncu, round 2: a third of the kernel's instructions).  Summing over
        // the stages FIRST, c_m = sum_l D(l, m) a_l for the six powers, is done once per step and leaves 36 FP64 per
        // save (six Horner chains).  The stage accelerations go through the thread's local memory and a ROLLED loop
        // over the stages on this path: unrolled, the 36 coefficients beside the 42 accelerations spilled all over the
        // step; as an out-of-line function, the second callee made ptxas save registers inside the right-hand side.
        if (__any_sync(FULL, want)) {
            double cq[3][6], cp[3][6];
            double inv_h = 0.0;
            // what the step still needs after the saves (the new state, the FSAL acceleration, the controller's output)
            // waits in local memory meanwhile: explicit, so that the allocator does not spill the step's long-lived
            // values all through the stage loop to make room for the coefficients here
            double stash[NS + 8];
            int zo;
            asm volatile("mov.u32 %0, 0;" : "=r"(zo));  // (opaque index: keeps the array in local memory)
            stash[zo + 0] = q1x; stash[zo + 1] = q1y; stash[zo + 2] = q1z; stash[zo + 3] = p1x; stash[zo + 4] = p1y;
            stash[zo + 5] = p1z; stash[zo + 6] = dt; stash[zo + 7] = inv; stash[zo + 8] = prev_inv; stash[zo + 9] = prev_prev_inv;
            stash[zo + 10] = AX(NS - 1); stash[zo + 11] = AY(NS - 1); stash[zo + 12] = AZ(NS - 1);
            if (want) {
                double al[NS][3];
#pragma unroll
                for (int l = 0; l < NS; ++l) { al[l][0] = AX(l); al[l][1] = AY(l); al[l][2] = AZ(l); }
                inv_h = rcp_fast(tnext - tprev);
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int m = 0; m < 6; ++m) { cq[c][m] = 0.0; cp[c][m] = 0.0; }
#pragma unroll 1
                for (int l = 0; l < NS; ++l) {
                    const double a0 = al[l][0], a1 = al[l][1], a2 = al[l][2];
#pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        const double dq = TB::DQ(l, m), db = TB::DB(l, m);
                        cq[0][m] = fma(dq, a0, cq[0][m]); cq[1][m] = fma(dq, a1, cq[1][m]); cq[2][m] = fma(dq, a2, cq[2][m]);
                        cp[0][m] = fma(db, a0, cp[0][m]); cp[1][m] = fma(db, a1, cp[1][m]); cp[2][m] = fma(db, a2, cp[2][m]);
                    }
                }
                // the step folded in: q = q0 + theta (hd p0 + hd^2 sum ...), p = p0 + theta hd sum ...
#pragma unroll
                for (int m = 0; m < 6; ++m) {
                    cq[0][m] *= hd2; cq[1][m] *= hd2; cq[2][m] *= hd2;
                    cp[0][m] *= hd; cp[1][m] *= hd; cp[2][m] *= hd;
                }
                cq[0][0] = fma(hd, p0x, cq[0][0]); cq[1][0] = fma(hd, p0y, cq[1][0]); cq[2][0] = fma(hd, p0z, cq[2][0]);
            }
            while (__any_sync(FULL, want)) {
                if (want) {
                    const double th = (tsave - tprev) * inv_h;
                    double vq[3], vp[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        double wq = cq[c][5], wp = cp[c][5];
#pragma unroll
                        for (int m = 4; m >= 0; --m) { wq = fma(wq, th, cq[c][m]); wp = fma(wp, th, cp[c][m]); }
                        vq[c] = wq; vp[c] = wp;
                    }
                    if constexpr (EPI) {
                        save_put_epi<EPI, C>(P, a, idx, k, fma(th, vq[0], q0x), fma(th, vq[1], q0y), fma(th, vq[2], q0z),
                                             fma(th, vp[0], p0x), fma(th, vp[1], p0y), fma(th, vp[2], p0z));
                    } else {
                        const SaveDst d = save_dst(a, idx, k);
                        d.q[0] = fma(th, vq[0], q0x);
                        d.q[d.st] = fma(th, vq[1], q0y);
                        d.q[2 * d.st] = fma(th, vq[2], q0z);
                        d.p[0] = fma(th, vp[0], p0x);
                        d.p[d.st] = fma(th, vp[1], p0y);
                        d.p[2 * d.st] = fma(th, vp[2], p0z);
                        save_commit(a, idx, k);
                    }
                    ++k;
                    tsave = (k < a.T) ? __ldg(a.ts + k) * dir : INF;
                    want = tsave <= tnext;
                }
            }
            q1x = stash[zo + 0]; q1y = stash[zo + 1]; q1z = stash[zo + 2]; p1x = stash[zo + 3]; p1y = stash[zo + 4];
            p1z = stash[zo + 5]; dt = stash[zo + 6]; inv = stash[zo + 7]; prev_inv = stash[zo + 8]; prev_prev_inv = stash[zo + 9];
            AX(NS - 1) = stash[zo + 10]; AY(NS - 1) = stash[zo + 11]; AZ(NS - 1) = stash[zo + 12];
        }
#else
        while (__any_sync(FULL, want)) {
            if (want) {
                const double th = (tsave - tprev) / (tnext - tprev), th2 = th * th;
                double wqx = 0, wqy = 0, wqz = 0, wpx = 0, wpy = 0, wpz = 0;
#pragma unroll
                for (int l = 0; l < NS; ++l) {
                    if (dq_row_nz<TB>(l)) {
                        const double w = dense_weight<TB, true>(l, th, th2);
                        wqx = fma(w, AX(l), wqx); wqy = fma(w, AY(l), wqy); wqz = fma(w, AZ(l), wqz);
                    }
                    if (db_row_nz<TB>(l)) {
                        const double w = dense_weight<TB, false>(l, th, th2);
                        wpx = fma(w, AX(l), wpx); wpy = fma(w, AY(l), wpy); wpz = fma(w, AZ(l), wpz);
                    }
                }
                const double thh = th * hd;
                if constexpr (EPI) {
                    save_put_epi<EPI, C>(P, a, idx, k, fma(hd2, wqx, fma(thh, p0x, q0x)), fma(hd2, wqy, fma(thh, p0y, q0y)),
                                         fma(hd2, wqz, fma(thh, p0z, q0z)), fma(hd, wpx, p0x), fma(hd, wpy, p0y),
                                         fma(hd, wpz, p0z));
                } else {
                    const SaveDst d = save_dst(a, idx, k);
                    d.q[0] = fma(hd2, wqx, fma(thh, p0x, q0x));
                    d.q[d.st] = fma(hd2, wqy, fma(thh, p0y, q0y));
                    d.q[2 * d.st] = fma(hd2, wqz, fma(thh, p0z, q0z));
                    d.p[0] = fma(hd, wpx, p0x);
                    d.p[d.st] = fma(hd, wpy, p0y);
                    d.p[2 * d.st] = fma(hd, wpz, p0z);
                    save_commit(a, idx, k);
                }
                ++k;
                tsave = (k < a.T) ? __ldg(a.ts + k) * dir : INF;
                want = tsave <= tnext;
            }
        }
#endif
        if (keep) {
            q0x = q1x; q0y = q1y; q0z = q1z; p0x = p1x; p0y = p1y; p0z = p1z;
            fsx = AX(NS - 1); fsy = AY(NS - 1); fsz = AZ(NS - 1);
            prev_prev_inv = prev_inv;
            prev_inv = inv;
            tprev = tnext;
            ++nacc;
            if (!finite6(q0x, q0y, q0z, p0x, p0y, p0z)) st = GX_NONFINITE;
        }
        if (run) {
            if (tprev > T1) tprev = T1;
            tnext = clip_to_end(tprev, tprev + dt, T1, keep);
        }
    }
#undef AX
#undef AY
#undef AZ
}

// ================================================================================================
// K3j: the reference's JOINT batch semantics -- one shared adaptive step for the whole batch
// ================================================================================================
// evaluate_orbit(pot, w0[N,6], t) / OrbitSolver.solve(field, (q[N,3], p[N,3]), t0, t1) with scalar times hand the whole
// batch to ONE diffeqsolve (dynamics/_src/orbit/solver.py:774-803, legacy/integrator.py:288-298): a single 6N-dimensional
// ODE, one step size, the error norm an RMS over all 6N numbers, one accept / reject decision per attempt.  K3 controls
// the step per particle (what the reference does under vmap); this kernel restates the joint form for callers that need
// the reference's numbers for that call form: a persistent cooperative grid, every thread walks its share of the
// particles through the 13 stages of the attempt (state in a global ping-pong buffer, stage loop rolled -- this is the
// semantic mode, not the fast one), the squared scaled errors are reduced in a FIXED order (per thread, per CTA tree,
// then the CTA partials by every CTA the same way: run-to-run reproducible), one grid barrier per attempt, and every
// thread then takes the same controller decision from the same number.  SaveAt values are evaluated speculatively
// while the stages are in registers (a rejected attempt's saves are overwritten by the step that finally contains them).
namespace cgx = cooperative_groups;

struct JointArgs {
    const double *q0, *p0, *ts;
    double *q, *p;
    int *status, *n_acc, *n_tot;
    double *ws;  // state[2][9][N] (q, p, FSAL acceleration; SoA), then 2 x JOINT_PART x gridDim partial sums
    long long N, max_steps;
    long long sn, sk, sc;
    double t0, t1;
    double rtol, atol, pcoeff, icoeff, dcoeff, safety, factormin, factormax, dtmin, dtmax, dt0;
    int T;
};
constexpr int JOINT_PART = 3;       // values per reduction (sums of squares / non-finite count)
constexpr int JOINT_MAX_GRID = 2048;

// Sum over the grid of JOINT_PART per-thread values, identical in every thread, in a fixed order.  `slot` alternates so
// that a CTA still reading the partials of one reduction cannot see the next one's.
__device__ __forceinline__ void joint_reduce(cgx::grid_group &grid, double *part, int slot, double v[JOINT_PART]) {
    __shared__ double sh[JOINT_PART][128];
    __shared__ double tot[JOINT_PART];
    const int tid = threadIdx.x, bd = blockDim.x;
#pragma unroll
    for (int c = 0; c < JOINT_PART; ++c) sh[c][tid] = v[c];
    __syncthreads();
    for (int w = bd >> 1; w > 0; w >>= 1) {
        if (tid < w) {
#pragma unroll
            for (int c = 0; c < JOINT_PART; ++c) sh[c][tid] += sh[c][tid + w];
        }
        __syncthreads();
    }
    double *mine = part + (size_t)slot * JOINT_PART * JOINT_MAX_GRID;
    if (tid < JOINT_PART) mine[tid * JOINT_MAX_GRID + blockIdx.x] = sh[tid][0];
    __threadfence();
    grid.sync();
    // every CTA adds the partials the same way: thread t takes blocks t, t + bd, ... then the same tree as above
#pragma unroll
    for (int c = 0; c < JOINT_PART; ++c) {
        double acc = 0.0;
        for (int b = tid; b < (int)gridDim.x; b += bd) acc += __ldcg(mine + c * JOINT_MAX_GRID + b);
        sh[c][tid] = acc;
    }
    __syncthreads();
    for (int w = bd >> 1; w > 0; w >>= 1) {
        if (tid < w) {
#pragma unroll
            for (int c = 0; c < JOINT_PART; ++c) sh[c][tid] += sh[c][tid + w];
        }
        __syncthreads();
    }
    if (tid < JOINT_PART) tot[tid] = sh[tid][0];
    __syncthreads();
#pragma unroll
    for (int c = 0; c < JOINT_PART; ++c) v[c] = tot[c];
    __syncthreads();
}

template <class C, class TB>
__global__ void __launch_bounds__(128) k_integrate_joint(const __grid_constant__ DevPot P, const JointArgs a) {
    constexpr int NS = TB::NS;
    cgx::grid_group grid = cgx::this_grid();
    const long long N = a.N, gt = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    double *S[2] = {a.ws, a.ws + 9 * N};
    double *part = a.ws + 18 * N;
    const double INF = __longlong_as_double(0x7ff0000000000000LL), NANV = __longlong_as_double(0x7ff8000000000000LL);
    const double dir = (a.t1 >= a.t0) ? 1.0 : -1.0, T0 = a.t0 * dir, T1 = a.t1 * dir;
    const double n6 = 6.0 * (double)N;
    int slot = 0, cur = 0, k = 0;
    auto acc_at = [&](double x, double y, double z, double t, double &ax, double &ay, double &az) {
        double g0, g1, g2;
        gradient<C>(P, x, y, z, g0, g1, g2, t);
        ax = -g0; ay = -g1; az = -g2;
    };
    auto out_q = [&](long long i, int kk, int c) -> double & { return a.q[i * a.sn + kk * a.sk + c * a.sc]; };
    auto out_p = [&](long long i, int kk, int c) -> double & { return a.p[i * a.sn + kk * a.sk + c * a.sc]; };
    auto ts_at = [&](int kk) { return (kk < a.T) ? __ldg(a.ts + kk) * dir : INF; };

    // ---- t0: FSAL acceleration, saves at t0, and the initial step (PIDController.init / _select_initial_step on the
    //      joint state: the norms are RMS over all 6N numbers)
    int k0 = 0;
    while (ts_at(k0) <= T0) ++k0;
    double v[JOINT_PART] = {0.0, 0.0, 0.0};
    for (long long i = gt; i < N; i += stride) {
        double y[6] = {a.q0[3 * i], a.q0[3 * i + 1], a.q0[3 * i + 2], a.p0[3 * i], a.p0[3 * i + 1], a.p0[3 * i + 2]};
        double f[3];
        acc_at(y[0], y[1], y[2], dir * T0, f[0], f[1], f[2]);
#pragma unroll
        for (int c = 0; c < 6; ++c) S[0][c * N + i] = y[c];
#pragma unroll
        for (int c = 0; c < 3; ++c) S[0][(6 + c) * N + i] = f[c];
        for (int kk = 0; kk < k0; ++kk)
            for (int c = 0; c < 3; ++c) { out_q(i, kk, c) = y[c]; out_p(i, kk, c) = y[3 + c]; }
        const double f0[6] = {y[3] * dir, y[4] * dir, y[5] * dir, f[0] * dir, f[1] * dir, f[2] * dir};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const double sc = a.atol + fabs(y[c]) * a.rtol;
            const double u = y[c] / sc, w = f0[c] / sc;
            v[0] = fma(u, u, v[0]);
            v[1] = fma(w, w, v[1]);
        }
    }
    k = k0;
    double h;
    if (a.dt0 > 0.0) {
        h = a.dt0;
    } else {
        joint_reduce(grid, part, slot, v); slot ^= 1;
        const double d0 = sqrt(v[0] / n6), d1 = sqrt(v[1] / n6);
        const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        v[0] = v[1] = v[2] = 0.0;
        for (long long i = gt; i < N; i += stride) {
            double y[6], f[3], f1[3];
#pragma unroll
            for (int c = 0; c < 6; ++c) y[c] = S[0][c * N + i];
#pragma unroll
            for (int c = 0; c < 3; ++c) f[c] = S[0][(6 + c) * N + i];
            const double f0[6] = {y[3] * dir, y[4] * dir, y[5] * dir, f[0] * dir, f[1] * dir, f[2] * dir};
            double y1[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) y1[c] = y[c] + h0 * f0[c];
            acc_at(y1[0], y1[1], y1[2], dir * (T0 + h0), f1[0], f1[1], f1[2]);
            const double g1[6] = {y1[3] * dir, y1[4] * dir, y1[5] * dir, f1[0] * dir, f1[1] * dir, f1[2] * dir};
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const double w = (g1[c] - f0[c]) / (a.atol + fabs(y[c]) * a.rtol);
                v[0] = fma(w, w, v[0]);
            }
        }
        joint_reduce(grid, part, slot, v); slot ^= 1;
        const double d2 = sqrt(v[0] / n6) / h0;
        const double maxd = fmax(d1, d2);
        const double h1 = (maxd <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / maxd, 1.0 / (TB::ORDER + 1));
        h = fmin(100.0 * h0, h1);
    }
    bool at_dtmin = false;
    if (a.dtmax > 0.0) h = fmin(h, a.dtmax);
    if (a.dtmin > 0.0) { at_dtmin = h <= a.dtmin; h = fmax(h, a.dtmin); }
    double tprev = T0, tnext = clip_to_end(T0, T0 + h, T1, true);
    double prev_inv = 1.0, prev_prev_inv = 1.0;
    long long nacc = 0, ntot = 0;
    int st = GX_OK;
    grid.sync();  // (dt0 given: the state written above must be visible before the first attempt reads it -- own
                  //  particles only, but keep the phases aligned)

    while (tprev < T1 && st == GX_OK) {
        if (a.max_steps >= 0 && ntot >= a.max_steps) { st = GX_MAX_STEPS_REACHED; break; }
        h = tnext - tprev;
        const double hd = h * dir, hd2 = hd * hd;
        int kend = k;  // saves k .. kend-1 lie in (tprev, tnext]
        while (ts_at(kend) <= tnext) ++kend;
        v[0] = v[1] = v[2] = 0.0;
        const double *Sc = S[cur];
        double *Sn = S[cur ^ 1];
        for (long long i = gt; i < N; i += stride) {
            const double q0x = Sc[i], q0y = Sc[N + i], q0z = Sc[2 * N + i];
            const double p0x = Sc[3 * N + i], p0y = Sc[4 * N + i], p0z = Sc[5 * N + i];
            double ax[NS], ay[NS], az[NS];
            ax[0] = Sc[6 * N + i]; ay[0] = Sc[7 * N + i]; az[0] = Sc[8 * N + i];
            double xi = q0x, yi = q0y, zi = q0z;
#pragma unroll 1
            for (int s = 1; s < NS; ++s) {
                double sx = 0, sy = 0, sz = 0;
                for (int l = 0; l < s; ++l) {
                    const double c = TB::AA(s, l);
                    sx = fma(c, ax[l], sx); sy = fma(c, ay[l], sy); sz = fma(c, az[l], sz);
                }
                const double ch = TB::CN(s) * hd;
                xi = fma(hd2, sx, fma(ch, p0x, q0x));
                yi = fma(hd2, sy, fma(ch, p0y, q0y));
                zi = fma(hd2, sz, fma(ch, p0z, q0z));
                acc_at(xi, yi, zi, fma(TB::CN(s), hd, dir * tprev), ax[s], ay[s], az[s]);
            }
            double bx = 0, by = 0, bz = 0, epx = 0, epy = 0, epz = 0, eqx = 0, eqy = 0, eqz = 0;
            for (int l = 0; l < NS; ++l) {
                bx = fma(TB::B(l), ax[l], bx); by = fma(TB::B(l), ay[l], by); bz = fma(TB::B(l), az[l], bz);
                epx = fma(TB::E(l), ax[l], epx); epy = fma(TB::E(l), ay[l], epy); epz = fma(TB::E(l), az[l], epz);
                eqx = fma(TB::EA(l), ax[l], eqx); eqy = fma(TB::EA(l), ay[l], eqy); eqz = fma(TB::EA(l), az[l], eqz);
            }
            const double y0[6] = {q0x, q0y, q0z, p0x, p0y, p0z};
            const double y1[6] = {xi, yi, zi, fma(hd, bx, p0x), fma(hd, by, p0y), fma(hd, bz, p0z)};
            const double er[6] = {hd2 * eqx, hd2 * eqy, hd2 * eqz, hd * epx, hd * epy, hd * epz};
            bool fin = true;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const double e = er[c] / (a.atol + fmax(fabs(y0[c]), fabs(y1[c])) * a.rtol);
                v[0] = fma(e, e, v[0]);
                fin = fin && isfinite(y1[c]);
                Sn[c * N + i] = y1[c];
            }
            if (!fin) v[1] += 1.0;
            Sn[6 * N + i] = ax[NS - 1]; Sn[7 * N + i] = ay[NS - 1]; Sn[8 * N + i] = az[NS - 1];
            // SaveAt(ts) on this attempt (dense output; a save time equal to tnext gets theta = 1 like the rest)
            for (int kk = k; kk < kend; ++kk) {
                const double th = (ts_at(kk) - tprev) / (tnext - tprev);
                double wqx = 0, wqy = 0, wqz = 0, wpx = 0, wpy = 0, wpz = 0;
                for (int l = 0; l < NS; ++l) {
                    double wq = TB::DQ(l, 5), wp = TB::DB(l, 5);
                    for (int m = 4; m >= 0; --m) { wq = fma(wq, th, TB::DQ(l, m)); wp = fma(wp, th, TB::DB(l, m)); }
                    wq *= th; wp *= th;
                    wqx = fma(wq, ax[l], wqx); wqy = fma(wq, ay[l], wqy); wqz = fma(wq, az[l], wqz);
                    wpx = fma(wp, ax[l], wpx); wpy = fma(wp, ay[l], wpy); wpz = fma(wp, az[l], wpz);
                }
                const double thh = th * hd;
                out_q(i, kk, 0) = fma(hd2, wqx, fma(thh, p0x, q0x));
                out_q(i, kk, 1) = fma(hd2, wqy, fma(thh, p0y, q0y));
                out_q(i, kk, 2) = fma(hd2, wqz, fma(thh, p0z, q0z));
                out_p(i, kk, 0) = fma(hd, wpx, p0x);
                out_p(i, kk, 1) = fma(hd, wpy, p0y);
                out_p(i, kk, 2) = fma(hd, wpz, p0z);
            }
        }
        joint_reduce(grid, part, slot, v); slot ^= 1;
        ++ntot;
        // ---- PID controller on the joint error norm (diffrax PIDController.adapt_step_size), the same in every thread
        const double ms = v[0] / n6;
        const bool bad = !(ms <= 1.7976931348623157e308);
        bool keep = !bad && ms < 1.0;
        if (a.dtmin > 0.0) keep = keep || at_dtmin;
        const double serr = bad ? INF : sqrt(ms);
        double inv = 1.0 / serr;
        const double c1 = (a.icoeff + a.pcoeff + a.dcoeff) * (1.0 / TB::ORDER);
        const double c2 = -(a.pcoeff + 2.0 * a.dcoeff) * (1.0 / TB::ORDER);
        const double c3 = a.dcoeff * (1.0 / TB::ORDER);
        double factor = a.safety;
        if (bad) {
            factor = 0.0;
        } else if (!(ms > 0.0)) {
            factor = a.factormax;
        } else {
            if (c1 != 0.0) factor *= pow(inv, c1);
            if (c2 != 0.0) factor *= pow(prev_inv, c2);
            if (c3 != 0.0) factor *= pow(prev_prev_inv, c3);
        }
        factor = fmin(fmax(factor, keep ? 1.0 : a.factormin), a.factormax);
        double dt = h * factor;
        if (inv == 0.0 || isinf(inv)) { inv = 1.0; prev_inv = 1.0; }
        if (a.dtmax > 0.0) dt = fmin(dt, a.dtmax);
        if (a.dtmin > 0.0) { at_dtmin = dt <= a.dtmin; dt = fmax(dt, a.dtmin); }
        if (keep) {
            cur ^= 1;
            prev_prev_inv = prev_inv;
            prev_inv = inv;
            tprev = tnext;
            ++nacc;
            k = kend;
            if (v[1] > 0.0) st = GX_NONFINITE;
        }
        if (tprev > T1) tprev = T1;
        tnext = clip_to_end(tprev, tprev + dt, T1, keep);
    }
    // saves never reached: NaN, like an unfilled diffrax buffer
    for (long long i = gt; i < N; i += stride)
        for (int kk = k; kk < a.T; ++kk)
            for (int c = 0; c < 3; ++c) { out_q(i, kk, c) = NANV; out_p(i, kk, c) = NANV; }
    if (gt == 0) {
        if (a.status) *a.status = st;
        if (a.n_acc) *a.n_acc = (int)nacc;
        if (a.n_tot) *a.n_tot = (int)ntot;
    }
}

// Dense output in parallel: one thread per save time, binary search over the recorded accepted steps, then the
// same degree-6 continuous extension as the in-kernel SaveAt path.  Used for single orbits with many saves (the
// progenitor orbit of a mock stream: 5e5 saves on ~10^3 steps), where a serial in-kernel evaluation would dominate.
template <class TB>
__global__ void __launch_bounds__(256) k_dense_eval(const double *__restrict__ rec, const int *__restrict__ n_rec_p,
                                                    double t0, double t1, const double *__restrict__ ts, long long M,
                                                    double *__restrict__ q, double *__restrict__ p) {
    constexpr int NS = TB::NS;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int n_rec = *n_rec_p;
    const double dir = (t1 >= t0) ? 1.0 : -1.0;
    const double tau = ts[i] * dir, T0 = t0 * dir;
    const double NANV = __longlong_as_double(0x7ff8000000000000LL);
    double *qo = q + 3 * i, *po = p + 3 * i;
    if (n_rec <= 0) {
        qo[0] = qo[1] = qo[2] = po[0] = po[1] = po[2] = NANV;
        return;
    }
    if (tau <= T0) {  // ts == t0 returns y0
        const double *r = rec;
        qo[0] = r[3]; qo[1] = r[4]; qo[2] = r[5]; po[0] = r[6]; po[1] = r[7]; po[2] = r[8];
        return;
    }
    // first record with tnext >= tau  (records are contiguous in time: tnext_j == tprev_{j+1})
    int lo = 0, hi = n_rec - 1;
    if (tau > rec[(long long)hi * REC_DOUBLES + 1]) {  // beyond the recorded range (failed / truncated solve)
        qo[0] = qo[1] = qo[2] = po[0] = po[1] = po[2] = NANV;
        return;
    }
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rec[(long long)mid * REC_DOUBLES + 1] >= tau) hi = mid; else lo = mid + 1;
    }
    const double *r = rec + (long long)lo * REC_DOUBLES;
    const double tprev = r[0], tnext = r[1], hd = r[2], hd2 = hd * hd;
    const double th = (tau - tprev) / (tnext - tprev), th2 = th * th;
    double wqx = 0, wqy = 0, wqz = 0, wpx = 0, wpy = 0, wpz = 0;
#pragma unroll
    for (int l = 0; l < NS; ++l) {
        const double ax = r[9 + 3 * l], ay = r[10 + 3 * l], az = r[11 + 3 * l];
        if (dq_row_nz<TB>(l)) {
            const double w = dense_weight<TB, true>(l, th, th2);
            wqx = fma(w, ax, wqx); wqy = fma(w, ay, wqy); wqz = fma(w, az, wqz);
        }
        if (db_row_nz<TB>(l)) {
            const double w = dense_weight<TB, false>(l, th, th2);
            wpx = fma(w, ax, wpx); wpy = fma(w, ay, wpy); wpz = fma(w, az, wpz);
        }
    }
    const double thh = th * hd;
    qo[0] = fma(hd2, wqx, fma(thh, r[6], r[3]));
    qo[1] = fma(hd2, wqy, fma(thh, r[7], r[4]));
    qo[2] = fma(hd2, wqz, fma(thh, r[8], r[5]));
    po[0] = fma(hd, wpx, r[6]);
    po[1] = fma(hd, wpy, r[7]);
    po[2] = fma(hd, wpz, r[8]);
}

// ================================================================================================
// K4 stream release (Fardal+15 / Chen+24)
// ================================================================================================

struct ReleaseArgs {
    const double *xq, *xp, *mass, *draws;
    double *ql, *pl, *qt, *pt;
    long long M;
    double G;
    int df;
    const double *t_rel;  // release times (time-dependent potentials: each particle sees the potential of its own time)
};

template <class C>
__global__ void __launch_bounds__(128) k_stream_release(const __grid_constant__ DevPot P, const ReleaseArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.M) return;
    const double x[3] = {a.xq[3 * i], a.xq[3 * i + 1], a.xq[3 * i + 2]};
    const double v[3] = {a.xp[3 * i], a.xp[3 * i + 1], a.xp[3 * i + 2]};
    const double r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const double L[3] = {x[1] * v[2] - x[2] * v[1], x[2] * v[0] - x[0] * v[2], x[0] * v[1] - x[1] * v[0]};
    // omega = |x cross v| / r^2     (register_api.py:77-88)
    const double om[3] = {L[0] / (r * r), L[1] / (r * r), L[2] / (r * r)};
    const double omega = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    // d2Phi/dr2 = rhat . H . rhat  (register_funcs.py:442-457); r_t = cbrt(G m / (omega^2 - d2Phi/dr2))
    double H[6];
    double gdummy[3];
    bool frozen = false;
    if constexpr (!C::is_static && !C::basic_only) {
        if (P.td.n > 0) {  // LinearParameter composite: the potential at this particle's release time
            FrozenPot F;
            freeze_td(P.td, a.t_rel[i], F);
            grad_hess<C>(F, x[0], x[1], x[2], gdummy, H);
            frozen = true;
        }
    }
    if (!frozen) grad_hess<C>(P, x[0], x[1], x[2], gdummy, H);
    const double rh[3] = {x[0] / r, x[1] / r, x[2] / r};
    const double d2 = rh[0] * (H[0] * rh[0] + H[1] * rh[1] + H[2] * rh[2]) +
                      rh[1] * (H[1] * rh[0] + H[3] * rh[1] + H[4] * rh[2]) +
                      rh[2] * (H[2] * rh[0] + H[4] * rh[1] + H[5] * rh[2]);
    const double m = a.mass[i];
    const double rt = cbrt(a.G * m / (omega * omega - d2));
    const double Ln = sqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
    const double zh[3] = {L[0] / Ln, L[1] / Ln, L[2] / Ln};
    const double vr = v[0] * rh[0] + v[1] * rh[1] + v[2] * rh[2];
    double ph[3] = {v[0] - vr * rh[0], v[1] - vr * rh[1], v[2] - vr * rh[2]};
    const double pn = sqrt(ph[0] * ph[0] + ph[1] * ph[1] + ph[2] * ph[2]);
    ph[0] /= pn; ph[1] /= pn; ph[2] /= pn;
    if (a.df == GX_DF_FARDAL15) {
        const double vc = omega * rt;
        const double kr = 2.0 + a.draws[0 * a.M + i] * 0.5;
        const double kvphi = kr * (0.3 + a.draws[1 * a.M + i] * 0.5);
        const double kz = 0.0 + a.draws[2 * a.M + i] * 0.5;
        const double kvz = 0.0 + a.draws[3 * a.M + i] * 0.5;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a.qt[3 * i + c] = x[c] + rt * (kr * rh[c] + kz * zh[c]);
            a.pt[3 * i + c] = v[c] + vc * (kvphi * ph[c] + kvz * zh[c]);
            a.ql[3 * i + c] = x[c] - rt * (kr * rh[c] - kz * zh[c]);
            a.pl[3 * i + c] = v[c] - vc * (kvphi * ph[c] - kvz * zh[c]);
        }
    } else {
        const double D2R = 0.017453292519943295;
        const double *pv = a.draws + 6 * i;
        const double Dr = pv[0] * rt;
        const double vesc = sqrt(2.0 * a.G * m / Dr);
        const double Dv = pv[3] * vesc;
        double sp, cp, st_, ct, sa, ca, sb, cb;
        sincos(pv[1] * D2R, &sp, &cp);
        sincos(pv[2] * D2R, &st_, &ct);
        sincos(pv[4] * D2R, &sa, &ca);
        sincos(pv[5] * D2R, &sb, &cb);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double dx = (Dr * ct * cp) * rh[c], dy = (Dr * ct * sp) * ph[c], dz = (Dr * st_) * zh[c];
            const double ex = (Dv * cb * ca) * rh[c], ey = (Dv * cb * sa) * ph[c], ez = (Dv * sb) * zh[c];
            a.qt[3 * i + c] = x[c] + dx + dy + dz;
            a.pt[3 * i + c] = v[c] + ex + ey + ez;
            a.ql[3 * i + c] = x[c] - dx - dy + dz;
            a.pl[3 * i + c] = v[c] - ex - ey + ez;
        }
    }
}

// ================================================================================================
// diagnostics, probes, FP64 peak microbenchmark
// ================================================================================================

template <class C>
__global__ void __launch_bounds__(256) k_energy_angmom(const __grid_constant__ DevPot P, const double *q,
                                                       const double *p, long long N, double *E, double *L,
                                                       const double *tv, double ts, long long t_period) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
    const double vx = p[3 * i], vy = p[3 * i + 1], vz = p[3 * i + 2];
    if (E) {
        double phi;
        bool frozen = false;
        if constexpr (!C::is_static && !C::basic_only) {
            if (P.td.n > 0) {  // LinearParameter composite: Phi(q, t) at the state's own time (tv[i mod t_period] or ts)
                FrozenPot F;
                freeze_td(P.td, tv ? tv[i % t_period] : ts, F);
                phi = potential_value<C>(F, x, y, z);
                frozen = true;
            }
        }
        if (!frozen) phi = potential_value<C>(P, x, y, z);
        E[i] = kinetic_plus(phi, vx, vy, vz);
    }
    if (L) cross3(x, y, z, vx, vy, vz, L[3 * i], L[3 * i + 1], L[3 * i + 2]);
}

__global__ void k_bench_dfma(long long iters, double *sink) {
    // 8 independent chains per thread; x <- x*m + c keeps every value bounded.
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    const double m = 0.999999, c = 1e-6;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, m, c); x1 = fma(x1, m, c); x2 = fma(x2, m, c); x3 = fma(x3, m, c);
            x4 = fma(x4, m, c); x5 = fma(x5, m, c); x6 = fma(x6, m, c); x7 = fma(x7, m, c);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

// Same, but every DFMA reads three different registers (x <- x*y + z with y, z loop-carried), the operand pattern
// of real code: measures what the register file can feed the FP64 pipe.
__global__ void k_bench_dfma3(long long iters, double *sink) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    double y0 = 0.999999 + x0 * 1e-9, y1 = y0 - 1e-9, y2 = y0 - 2e-9, y3 = y0 - 3e-9;
    double z0 = 1e-6 + x0 * 1e-12, z1 = z0 * 1.1, z2 = z0 * 1.2, z3 = z0 * 1.3;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, y0, z0); x1 = fma(x1, y1, z1); x2 = fma(x2, y2, z2); x3 = fma(x3, y3, z3);
            x4 = fma(x4, y1, z0); x5 = fma(x5, y2, z1); x6 = fma(x6, y3, z2); x7 = fma(x7, y0, z3);
        }
        // keep y, z loop-carried so they stay in registers and are not folded
        y0 = fma(y0, 1.0, 0.0 * x0); z0 = fma(z0, 1.0, 0.0 * x1);
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) sink[0] = s + y0 + y1 + y2 + y3 + z0 + z1 + z2 + z3;
}

__global__ void k_debug_math(int op, const GammaTab *gt, const double *tab, const double *x, long long N, double *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double v = x[i];
    double r;
    switch (op) {
    case 0: r = rcp_fast(v); break;
    case 1: r = rsqrt_fast(v); break;
    case 2: r = log1p_pos(v); break;
    case 3: r = gammainc_P(*gt, v, nullptr); break;
    case 4: r = nfw_menc_shape(v); break;
    case 5:  // NFW force table F(s); NaN outside the tabulated range
        if (!poly_table_eval<false, NFW_E_LO, NFW_NINT>(tab, v, r, nullptr)) r = __longlong_as_double(0x7ff8000000000000LL);
        break;
    case 6:  // PowerLawCutoff table G(s) for exponent a; NaN outside the tabulated range
        if (!plc_table_eval_at<false>(tab, v, r, nullptr)) r = __longlong_as_double(0x7ff8000000000000LL);
        break;
    case 7: {  // ... and its derivative dG/ds
        double g;
        if (!plc_table_eval_at<false>(tab, v, g, &r)) r = __longlong_as_double(0x7ff8000000000000LL);
        break;
    }
    default: r = 0.0;
    }
    out[i] = r;
}

}  // namespace gx

// ================================================================================================
// C ABI
// ================================================================================================
using namespace gx;

// launch with `dyn` bytes of dynamic shared memory on top of the kernel's static tables (together they may exceed the
// 48 KB a kernel gets without opting in; the attribute is per device, so it is set on every such launch)
template <auto Kern, class... Args>
static inline void launch_dyn(int grid, int block, size_t dyn, cudaStream_t s, const Args &...args) {
    if (dyn) cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    Kern<<<grid, block, dyn, s>>>(args...);
}

// gx_strict.cu (compiled with -fmad=false): the reference-order kernels behind GX_SCHEME_STRICT
int gx_strict_integrate_fixed(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0,
                              double t1, double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps,
                              int32_t layout, double *q, double *p, int32_t *status, void *stream);

int gx_strict_integrate_adaptive(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                                 const double *p0, int64_t N, const double *t0, double t0_scalar, double t1,
                                 const double *ts, int32_t T, int64_t max_steps, int32_t layout, double *q, double *p,
                                 int32_t *status, int32_t *n_accepted, int32_t *n_attempted, void *stream);

extern "C" {

int gx_version(void) { return GX_VERSION; }

const char *gx_strerror(int code) {
    switch (code) {
    case 0: return "ok";
    case GX_ERR_BADARG: return "bad argument";
    case GX_ERR_UNSUPPORTED: return "unsupported potential component or parameter";
    case GX_ERR_CUDA: return "CUDA error";
    default: return "unknown error";
    }
}

int64_t gx_workspace_bytes(void) { return 256; }

static inline int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }


static void out_strides(int layout, long long N, int T, long long &sn, long long &sk, long long &sc) {
    if (layout == GX_LAYOUT_T3N) { sn = 1; sk = 3 * N; sc = N; }
    else { sn = 3LL * T; sk = 3; sc = 1; }
}

int gx_potential_eval(const gx_potential *pot, const double *xyz, double t, int64_t N, uint32_t what, double *phi,
                      double *grad, double *acc, double *hess, void *stream) {
    DevPot D; Model model;
    int rc = build_devpot(pot, D, model, true, TD_FREEZE, t);  // LinearParameter: the potential at time t
    if (rc) return rc;
    if (N < 0) return GX_ERR_BADARG;
    if (N == 0) return 0;
    if (!xyz) return GX_ERR_BADARG;
    if (((what & GX_PHI) && !phi) || ((what & GX_GRAD) && !grad) || ((what & GX_ACC) && !acc) ||
        ((what & GX_HESS) && !hess))
        return GX_ERR_BADARG;
    // cp.async.bulk needs 16-byte aligned global addresses; a tile starts 256 * 24 (or 72) bytes after the previous one,
    // so the bases decide.  Odd-row views of an [N,3] array (xyz[1:]) are only 8-byte aligned: plain path for those.
    const uintptr_t al = (uintptr_t)xyz | ((what & GX_GRAD) ? (uintptr_t)grad : 0) | ((what & GX_ACC) ? (uintptr_t)acc : 0) |
                         ((what & GX_HESS) ? (uintptr_t)hess : 0);
    if (al & 7) return GX_ERR_BADARG;
    EvalArgs a{xyz, phi, grad, acc, hess, (long long)N, what, (al & 15) == 0};
    const int block = EVAL_TILE;
    long long want = (N + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    if (what & GX_HESS) {
        // persistent CTAs: exactly the resident set (dynamic shared memory below; ask for the full carveout)
        int dev = 0, sms = 148, per_sm = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int outs = ((what & GX_GRAD) ? 3 : 0) + ((what & GX_ACC) ? 3 : 0) + 9;
        const size_t dyn = (size_t)EVAL_TILE * 8 * (2 * 3 + 2 * outs);  // 61-74 KB: 3 resident CTAs per SM
#define GX_LAUNCH_EVAL(C_)                                                                                    \
    do {                                                                                                      \
        auto kern = k_potential_eval<C_, EVAL_TILE>;                                                          \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);                    \
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);                      \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, dyn);                             \
        if (per_sm < 1) per_sm = 1;                                                                           \
        long long resident = (long long)per_sm * sms;                                                         \
        int grid = (int)(want < resident ? want : resident);                                                  \
        kern<<<grid, block, dyn, s>>>(D, a);                                                                  \
    } while (0)
        GX_DISPATCH_MODEL(model, GX_LAUNCH_EVAL(C));
#undef GX_LAUNCH_EVAL
    } else {
        int grid = (int)(want < 148LL * 32 ? want : 148LL * 32);
        GX_DISPATCH_MODEL(model, (k_potential_eval_direct<C><<<grid, block, 0, s>>>(D, a)));
    }
    return cuda_rc(cudaGetLastError());
}

}  // extern "C" (reopened below)

static int epi_setup(const gx_orbit_epilogue *epi, int layout, EpiOut &e) {
    e = EpiOut{nullptr, nullptr, nullptr, layout == GX_LAYOUT_T3N, 6};
    if (!epi) return 0;
    e.E = epi->energy; e.L = epi->angmom; e.TT = epi->tidal;
    if (!e.E && !e.L && !e.TT) return 0;
    e.nv = e.TT ? 19 : 10;
    return 1;
}

template <bool EPI>
static void launch_fixed(Model model, const DevPot &D, bool seg_ok, bool small, int scheme, bool fwd, int grid, int block,
                         size_t dyn, cudaStream_t s, const FixedArgs &a, const FixedSeg &sg) {
    if (seg_ok) {
        switch (model) {
#define GX_SEG_STATIC(C_)                                                                                     \
    do {                                                                                                      \
        if (small) {                                                                                          \
            if (fwd) launch_dyn<k_integrate_fixed_seg<C_, true, true, EPI>>(grid, block, dyn, s, D, a, sg);   \
            else launch_dyn<k_integrate_fixed_seg<C_, false, true, EPI>>(grid, block, dyn, s, D, a, sg);      \
        } else {                                                                                              \
            if (fwd) launch_dyn<k_integrate_fixed_seg<C_, true, false, EPI>>(grid, block, dyn, s, D, a, sg);  \
            else launch_dyn<k_integrate_fixed_seg<C_, false, false, EPI>>(grid, block, dyn, s, D, a, sg);     \
        }                                                                                                     \
    } while (0)
        case MODEL_MW: GX_SEG_STATIC(CountsMW); break;
        case MODEL_MW2022: GX_SEG_STATIC(CountsMW2022); break;
        case MODEL_BOVY: GX_SEG_STATIC(CountsBovy); break;
        default:  // runtime composite, no time-dependent parameter
            if (is_basic_composite(D, model)) {
                if (D.sph_wide) {  // spherical components in the combined table
                    GX_SEG_STATIC(CountsBasicTab);
                } else {
                    if (fwd) launch_dyn<k_integrate_fixed_seg<CountsBasic, true, false, EPI>>(grid, block, dyn, s, D, a, sg);
                    else launch_dyn<k_integrate_fixed_seg<CountsBasic, false, false, EPI>>(grid, block, dyn, s, D, a, sg);
                }
            } else {
                if (fwd) launch_dyn<k_integrate_fixed_seg<CountsRuntime, true, false, EPI>>(grid, block, dyn, s, D, a, sg);
                else launch_dyn<k_integrate_fixed_seg<CountsRuntime, false, false, EPI>>(grid, block, dyn, s, D, a, sg);
            }
            break;
        }
#undef GX_SEG_STATIC
        return;
    }
    // the step-by-step kernel (LeapfrogMidpoint, time-dependent parameters, > 120 runs, GX_SCHEME_GENERAL_KERNEL); the
    // static models take the same small-batch variant as above, so both kernels keep giving the same bits
#define GX_GEN(C_, SCHEME_, SMALL_)                                                                                \
    do {                                                                                                           \
        if (fwd) launch_dyn<k_integrate_fixed<C_, SCHEME_, true, EPI, SMALL_>>(grid, block, dyn, s, D, a);         \
        else launch_dyn<k_integrate_fixed<C_, SCHEME_, false, EPI, SMALL_>>(grid, block, dyn, s, D, a);            \
    } while (0)
#define GX_GEN_STATIC(C_)                                                                                          \
    do {                                                                                                           \
        if (scheme == GX_SCHEME_SEMI_IMPLICIT_EULER) {                                                             \
            GX_GEN(C_, GX_SCHEME_SEMI_IMPLICIT_EULER, false);                                                      \
        } else {                                                                                                   \
            GX_GEN(C_, GX_SCHEME_LEAPFROG_MIDPOINT, false);                                                        \
        }                                                                                                          \
    } while (0)
    switch (model) {
    case MODEL_MW: GX_GEN_STATIC(CountsMW); break;
    case MODEL_MW2022: GX_GEN_STATIC(CountsMW2022); break;
    case MODEL_BOVY: GX_GEN_STATIC(CountsBovy); break;
    default:
        if (is_basic_composite(D, model) && D.sph_wide) {  // (the same arithmetic as the run-length kernel's)
            if (scheme == GX_SCHEME_SEMI_IMPLICIT_EULER) GX_GEN(CountsBasicTab, GX_SCHEME_SEMI_IMPLICIT_EULER, false);
            else GX_GEN(CountsBasicTab, GX_SCHEME_LEAPFROG_MIDPOINT, false);
        } else if (scheme == GX_SCHEME_SEMI_IMPLICIT_EULER) GX_GEN(CountsRuntime, GX_SCHEME_SEMI_IMPLICIT_EULER, false);
        else GX_GEN(CountsRuntime, GX_SCHEME_LEAPFROG_MIDPOINT, false);
        break;
    }
#undef GX_GEN_STATIC
#undef GX_GEN
}

static int fixed_impl(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0, double t1,
                      double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps, int32_t layout,
                      double *q, double *p, int32_t *status, const gx_orbit_epilogue *epi, void *stream) {
    DevPot D; Model model;
    int rc = build_devpot(pot, D, model, true, TD_INTEGRATE, 0.0, !stream_is_capturing(stream), TAB_WIDE);
    if (rc) return rc;
    if (N < 0 || T < 0 || (N > 0 && (!q0 || !p0)) || (N > 0 && T > 0 && (!ts || !q || !p))) return GX_ERR_BADARG;
    const bool general_kernel = (scheme & GX_SCHEME_GENERAL_KERNEL) != 0;
    const bool strict = (scheme & GX_SCHEME_STRICT) != 0;
    scheme &= ~(GX_SCHEME_GENERAL_KERNEL | GX_SCHEME_STRICT);
    if (scheme != GX_SCHEME_SEMI_IMPLICIT_EULER && scheme != GX_SCHEME_LEAPFROG_MIDPOINT) return GX_ERR_BADARG;
    if (layout != GX_LAYOUT_NT3 && layout != GX_LAYOUT_T3N) return GX_ERR_BADARG;
    const double dir = (t1 >= t0) ? 1.0 : -1.0;
    if (t1 != t0 && !(dt0 * dir > 0.0)) return GX_ERR_BADARG;  // ConstantStepSize needs dt0 in the direction of t1
    FixedArgs a;
    const bool with_epi = epi_setup(epi, layout, a.epi) != 0;
    if (with_epi && (strict || D.td.n > 0)) return GX_ERR_UNSUPPORTED;  // (E needs Phi(q, t): static potentials only)
    if (N == 0) return 0;
    if (strict)
        return gx_strict_integrate_fixed(pot, q0, p0, N, t0, t1, dt0, ts, T, scheme, max_steps, layout, q, p, status,
                                         stream);
    FixedSeg sg;
    bool seg_ok = GX_FIXED_SEG && GX_FUSED_UPDATE && !general_kernel && scheme == GX_SCHEME_SEMI_IMPLICIT_EULER &&
                  (model != MODEL_GENERIC || D.td.n == 0);
    a.q0 = q0; a.p0 = p0; a.ts = ts; a.q = q; a.p = p; a.status = status;
    a.N = N; a.t0 = t0; a.t1 = t1; a.dt0 = dt0; a.T = T;
    out_strides(layout, N, T, a.sn, a.sk, a.sc);
    walk_time_grid(t0, t1, dt0, max_steps, sg, seg_ok, a.n_steps, a.hit_max_steps);
    // CTA width.  Kernels without a table: narrow CTAs for small batches, so that the particles spread over all SMs and
    // schedulers.  Kernels that hold the 132 KB wide table: ONE CTA per SM, as wide as the batch needs (a multiple of 32
    // between 128 and what the staging of the saves leaves room for, at most 1024 / 640 / 384 threads for the run-length
    // kernel / the step-by-step kernel / the kernels with the save epilogue) -- 10^4 particles are 79 CTAs of four warps,
    // one per scheduler; 1.2e6 particles are 1024-thread CTAs in 8 full waves of 148.
    const bool tabled = GX_SPH_TABLE && (model != MODEL_GENERIC || (is_basic_composite(D, model) && D.sph_wide));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int block;
    size_t dyn;
    a.stage_off = 0;
    if (tabled) {
        const size_t per_thread = save_stage_bytes(layout, T, 1, a.epi.nv);  // staging bytes per lane (0: direct stores)
        const size_t room = (size_t)226 * 1024 - SPHW_BYTES;
        const int cap = with_epi ? 384 : (seg_ok ? GX_FIXED_SEG_BLOCK : GX_FIXED_TAB_BLOCK);
        int maxb = per_thread ? (int)((room / per_thread) / 32 * 32) : cap;
        if (maxb > cap) maxb = cap;
        if (maxb < 32) return GX_ERR_UNSUPPORTED;
        // full waves: the fewest waves of `sms` CTAs that hold the batch, then the narrowest CTA that fills them
        const long long per_sm = (N + sms - 1) / sms;
        const long long waves = (per_sm + maxb - 1) / maxb;
        const long long want = ((per_sm + waves - 1) / waves + 31) / 32 * 32;
        block = (int)(want < 128 ? 128 : (want > maxb ? maxb : want));
        if (block > maxb) block = maxb;
        dyn = SPHW_BYTES + (size_t)block * per_thread;
        a.stage = per_thread != 0;
        a.stage_off = SPHW_BYTES / 8;
    } else {
        block = (N >= 148LL * 128 * 4) ? 128 : ((N >= 148LL * 64 * 2) ? 64 : 32);
        dyn = save_stage_bytes(layout, T, block, a.epi.nv);  // per-lane staging of the saves (0: direct stores)
        a.stage = dyn != 0;
    }
    const int grid = grid_for(N, block);
    cudaStream_t s = (cudaStream_t)stream;
    const bool fwd = dir > 0;
    // latency-bound launches (at most 4 warps per scheduler): the instantiation that fetches the table row early
    const bool small = tabled && seg_ok && block <= 512;
    if (with_epi) launch_fixed<true>(model, D, seg_ok, small, scheme, fwd, grid, block, dyn, s, a, sg);
    else launch_fixed<false>(model, D, seg_ok, small, scheme, fwd, grid, block, dyn, s, a, sg);
    return cuda_rc(cudaGetLastError());
}

extern "C" {

int gx_integrate_fixed(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0, double t1,
                       double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps, int32_t layout,
                       double *q, double *p, int32_t *status, void *stream) {
    return fixed_impl(pot, q0, p0, N, t0, t1, dt0, ts, T, scheme, max_steps, layout, q, p, status, nullptr, stream);
}

int gx_integrate_fixed_epilogue(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0,
                                double t1, double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps,
                                int32_t layout, double *q, double *p, int32_t *status, const gx_orbit_epilogue *epi,
                                void *stream) {
    return fixed_impl(pot, q0, p0, N, t0, t1, dt0, ts, T, scheme, max_steps, layout, q, p, status, epi, stream);
}

}  // extern "C"

// Shared implementation of every adaptive entry point: `solver` picks the tableau, `rec` (optional) switches the
// kernel from saving to recording accepted steps.
// The __constant__ image of the potential (c_pot_dp8) is shared by every launch on a device, so a launch may only read
// it when nothing can change it underneath, and may only change it when nobody else can still be reading it.  All of
// that is decided here WITHOUT blocking the host and without making any stream wait for work it did not depend on:
//   * the image already holds this potential: the launch's stream waits for the copy that wrote it (an event; a no-op
//     on the stream that issued the copy) and the kernel reads the image;
//   * it holds another potential and every launch that read it is finished, or was issued on this very stream (the new
//     copy is then ordered behind it): the image is rewritten in stream order, the kernel reads it;
//   * it holds another potential that a launch on ANOTHER stream may still be reading (cudaEventQuery: not ready):
//     the image is left alone and this launch runs the IMG = false kernel, which carries the potential in its own
//     kernel parameter (shared-memory copy per CTA; ~6 % slower, no state outside the launch);
//   * the stream is being captured into a CUDA graph: always the IMG = false kernel (a graph may be replayed at any
//     time, concurrently with anything).
// The mutex is held from the decision until the launch's last-use event is recorded, so two host threads cannot
// interleave "decide" and "launch".
namespace {
struct ImgUser { cudaStream_t stream; cudaEvent_t ev; };
struct ImgState {
    bool valid = false;
    DevPot content;
    cudaEvent_t staged = nullptr;
    cudaStream_t staged_on = nullptr;
    std::vector<ImgUser> users;
};
constexpr int IMG_MAXDEV = 64;
std::mutex g_img_mtx;
ImgState g_img[IMG_MAXDEV];
}  // namespace

// 1: launch the IMG = true kernel; 0: launch the IMG = false kernel; < 0: error.  Caller holds g_img_mtx.
static int use_const_image(const DevPot &D, cudaStream_t s, int dev) {
    if (dev < 0 || dev >= IMG_MAXDEV) return 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) return GX_ERR_CUDA;
    if (cap != cudaStreamCaptureStatusNone) return 0;
    ImgState &I = g_img[dev];
    if (!I.staged && cudaEventCreateWithFlags(&I.staged, cudaEventDisableTiming) != cudaSuccess) return GX_ERR_CUDA;
    if (I.valid && memcmp(&I.content, &D, sizeof D) == 0) {
        if (I.staged_on != s && cudaStreamWaitEvent(s, I.staged, 0) != cudaSuccess) return GX_ERR_CUDA;
        return 1;
    }
    bool busy = false;
    if (I.valid && I.staged_on != s && cudaEventQuery(I.staged) != cudaSuccess) busy = true;
    for (const ImgUser &u : I.users)
        if (u.stream != s && cudaEventQuery(u.ev) != cudaSuccess) busy = true;
    (void)cudaGetLastError();  // cudaErrorNotReady from the queries is an answer, not an error
    if (busy) return 0;
    memcpy(&I.content, &D, sizeof D);
    I.valid = false;
    if (cudaMemcpyToSymbolAsync(c_pot_dp8, &I.content, sizeof D, 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return GX_ERR_CUDA;
    if (cudaEventRecord(I.staged, s) != cudaSuccess) return GX_ERR_CUDA;
    I.staged_on = s;
    I.valid = true;
    for (size_t k = 0; k < I.users.size();) {  // everything on other streams has finished (checked above): retire it
        if (I.users[k].stream != s) { cudaEventDestroy(I.users[k].ev); I.users[k] = I.users.back(); I.users.pop_back(); }
        else ++k;
    }
    return 1;
}
// after the launch of an IMG = true kernel on `s`: remember that the image is in use there.  Caller holds g_img_mtx.
static int const_image_used(cudaStream_t s, int dev) {
    ImgState &I = g_img[dev];
    for (ImgUser &u : I.users)
        if (u.stream == s) return cuda_rc(cudaEventRecord(u.ev, s));
    ImgUser u{s, nullptr};
    if (cudaEventCreateWithFlags(&u.ev, cudaEventDisableTiming) != cudaSuccess) return GX_ERR_CUDA;
    I.users.push_back(u);
    return cuda_rc(cudaEventRecord(u.ev, s));
}

static int adaptive_impl(int solver, double *rec, int *n_rec, int rec_cap, const gx_potential *pot, const gx_pid *pid,
                         const double *q0, const double *p0, int64_t N, const double *t0, double t0_scalar, double t1,
                         const double *ts, int32_t T, int64_t max_steps, const int32_t *order, int32_t layout,
                         double *q, double *p, int32_t *status, int32_t *n_accepted, int32_t *n_attempted,
                         void *workspace, void *stream, const gx_orbit_epilogue *epi = nullptr) {
    const bool strict = (solver & GX_SOLVER_STRICT) != 0;
    solver &= ~GX_SOLVER_STRICT;
    if (solver != GX_SOLVER_DOPRI8 && solver != GX_SOLVER_DOPRI5) return GX_ERR_UNSUPPORTED;
    DevPot D; Model model;
    int rc = build_devpot(pot, D, model, !strict, TD_INTEGRATE, 0.0, !stream_is_capturing(stream), TAB_WIDE);
    if (rc) return rc;
    if (!pid || N < 0 || T < 0 || (N > 0 && (!q0 || !p0)) || (N > 0 && T > 0 && (!ts || !q || !p)) || !workspace)
        return GX_ERR_BADARG;
    if (!(pid->rtol >= 0.0) || !(pid->atol >= 0.0) || (pid->rtol == 0.0 && pid->atol == 0.0)) return GX_ERR_BADARG;
    if (layout != GX_LAYOUT_NT3 && layout != GX_LAYOUT_T3N) return GX_ERR_BADARG;
    Dp8Args a;
    const bool with_epi = epi_setup(epi, layout, a.epi) != 0;
    if (with_epi && (strict || rec || D.td.n > 0)) return GX_ERR_UNSUPPORTED;  // (E needs Phi(q, t): static potentials)
    if (N == 0) return 0;
    if (strict) {
        if (rec) return GX_ERR_UNSUPPORTED;
        return gx_strict_integrate_adaptive(solver, pot, pid, q0, p0, N, t0, t0_scalar, t1, ts, T, max_steps, layout, q, p,
                                            status, n_accepted, n_attempted, stream);
    }
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, 256, s);
    if (e != cudaSuccess) return GX_ERR_CUDA;
    a.q0 = q0; a.p0 = p0; a.t0v = t0; a.ts = ts; a.order = order; a.q = q; a.p = p;
    a.status = status; a.n_acc = n_accepted; a.n_tot = n_attempted;
    a.ticket = (unsigned long long *)workspace;
    a.rec = rec; a.n_rec = n_rec; a.rec_cap = rec_cap;
    a.N = N; a.max_steps = max_steps; a.t0s = t0_scalar; a.t1 = t1;
    a.rtol = pid->rtol; a.atol = pid->atol;
    a.pcoeff = pid->pcoeff; a.icoeff = pid->icoeff; a.dcoeff = pid->dcoeff;
    a.safety = pid->safety; a.factormin = pid->factormin; a.factormax = pid->factormax;
    a.dtmin = pid->dtmin; a.dtmax = pid->dtmax;
    a.dt0 = (pid->dt0 > 0.0) ? pid->dt0 : -1.0;
    a.T = T;
    out_strides(layout, N, T, a.sn, a.sk, a.sc);
    // persistent launch: resident CTAs only (occupancy query), never more lanes than particles; small batches
    // use narrow CTAs so that the few particles spread over all SMs.
    int per_sm = 1, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // Kernels that hold the 132 KB wide table: ONE CTA per SM, up to 384 threads (168 registers each), as wide as the batch
    // needs and the staging of the saves leaves room for; the others as before.
    const bool tabled = GX_SPH_TABLE && (model != MODEL_GENERIC || (is_basic_composite(D, model) && D.sph_wide));
    int block;
    size_t dyn;
    a.stage_off = 0;
    if (tabled) {
        const size_t per_thread = save_stage_bytes(layout, T, 1, a.epi.nv);
        const size_t room = (size_t)226 * 1024 - SPHW_BYTES;
        int maxb = per_thread ? (int)((room / per_thread) / 32 * 32) : 384;
        if (maxb > 384) maxb = 384;
        if (maxb < 32) return GX_ERR_UNSUPPORTED;
        long long want_b = ((N + sms - 1) / sms + 31) / 32 * 32;
        block = (int)(want_b < 32 ? 32 : (want_b > maxb ? maxb : want_b));
        dyn = SPHW_BYTES + (size_t)block * per_thread;
        a.stage = per_thread != 0;
        a.stage_off = SPHW_BYTES / 8;
    } else {
        block = (N >= 148LL * 128) ? 128 : ((N >= 148LL * 64) ? 64 : 32);
        dyn = save_stage_bytes(layout, T, block, a.epi.nv);  // per-lane staging of the saves (0: direct stores)
        a.stage = dyn != 0;
    }
#define GX_LAUNCH_DP8_(C_, IMG_)                                                                              \
    do {                                                                                                      \
        auto kern = with_epi ? ((solver == GX_SOLVER_DOPRI5) ? k_integrate_dopri8<C_, TabDp5, IMG_, true>     \
                                                             : k_integrate_dopri8<C_, TabDp8, IMG_, true>)    \
                             : ((solver == GX_SOLVER_DOPRI5) ? k_integrate_dopri8<C_, TabDp5, IMG_>           \
                                                             : k_integrate_dopri8<C_, TabDp8, IMG_>);         \
        if (dyn) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);           \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, dyn);                             \
        if (per_sm < 1) per_sm = 1;                                                                           \
        long long want = (N + block - 1) / block, resident = (long long)per_sm * sms;                         \
        int grid = (int)(want < resident ? want : resident);                                                  \
        kern<<<grid, block, dyn, s>>>(D, a);                                                                  \
    } while (0)
#define GX_LAUNCH_DP8(C_)                                                                                     \
    do {                                                                                                      \
        if (img) GX_LAUNCH_DP8_(C_, true);                                                                    \
        else GX_LAUNCH_DP8_(C_, false);                                                                       \
    } while (0)
    std::lock_guard<std::mutex> img_lock(g_img_mtx);
    const int img = use_const_image(D, s, dev);
    if (img < 0) return img;
    if (is_basic_composite(D, model)) {
        if (D.sph_wide) GX_LAUNCH_DP8(CountsBasicTab);
        else GX_LAUNCH_DP8(CountsBasic);
    } else GX_DISPATCH_MODEL(model, GX_LAUNCH_DP8(C));
    rc = cuda_rc(cudaGetLastError());
    if (rc == 0 && img) rc = const_image_used(s, dev);
    return rc;
#undef GX_LAUNCH_DP8_
#undef GX_LAUNCH_DP8
}

extern "C" {

int gx_integrate_adaptive(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                          const double *p0, int64_t N, const double *t0, double t0_scalar, double t1, const double *ts,
                          int32_t T, int64_t max_steps, const int32_t *order, int32_t layout, double *q, double *p,
                          int32_t *status, int32_t *n_accepted, int32_t *n_attempted, void *workspace, void *stream) {
    return adaptive_impl(solver, nullptr, nullptr, 0, pot, pid, q0, p0, N, t0, t0_scalar, t1, ts, T, max_steps, order,
                         layout, q, p, status, n_accepted, n_attempted, workspace, stream);
}

int gx_integrate_adaptive_epilogue(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                                   const double *p0, int64_t N, const double *t0, double t0_scalar, double t1,
                                   const double *ts, int32_t T, int64_t max_steps, const int32_t *order, int32_t layout,
                                   double *q, double *p, int32_t *status, int32_t *n_accepted, int32_t *n_attempted,
                                   void *workspace, const gx_orbit_epilogue *epi, void *stream) {
    return adaptive_impl(solver, nullptr, nullptr, 0, pot, pid, q0, p0, N, t0, t0_scalar, t1, ts, T, max_steps, order,
                         layout, q, p, status, n_accepted, n_attempted, workspace, stream, epi);
}

int64_t gx_joint_workspace_bytes(int64_t N) {
    return N < 0 ? 0 : (int64_t)sizeof(double) * (18 * N + 2 * JOINT_PART * JOINT_MAX_GRID);
}

int gx_integrate_adaptive_joint(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                                const double *p0, int64_t N, double t0, double t1, const double *ts, int32_t T,
                                int64_t max_steps, int32_t layout, double *q, double *p, int32_t *status,
                                int32_t *n_accepted, int32_t *n_attempted, void *workspace, void *stream) {
    if (solver != GX_SOLVER_DOPRI8 && solver != GX_SOLVER_DOPRI5) return GX_ERR_UNSUPPORTED;
    DevPot D; Model model;
    int rc = build_devpot(pot, D, model, true, TD_INTEGRATE, 0.0, !stream_is_capturing(stream));
    if (rc) return rc;
    if (!pid || N < 0 || T < 0 || (N > 0 && (!q0 || !p0)) || (N > 0 && T > 0 && (!ts || !q || !p)) || !workspace)
        return GX_ERR_BADARG;
    if (!(pid->rtol >= 0.0) || !(pid->atol >= 0.0) || (pid->rtol == 0.0 && pid->atol == 0.0)) return GX_ERR_BADARG;
    if (layout != GX_LAYOUT_NT3 && layout != GX_LAYOUT_T3N) return GX_ERR_BADARG;
    if (N == 0) return 0;
    int dev = 0, sms = 148, coop = 0, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return GX_ERR_UNSUPPORTED;
    JointArgs a;
    a.q0 = q0; a.p0 = p0; a.ts = ts; a.q = q; a.p = p;
    a.status = status; a.n_acc = n_accepted; a.n_tot = n_attempted;
    a.ws = (double *)workspace;
    a.N = N; a.max_steps = max_steps; a.t0 = t0; a.t1 = t1;
    a.rtol = pid->rtol; a.atol = pid->atol;
    a.pcoeff = pid->pcoeff; a.icoeff = pid->icoeff; a.dcoeff = pid->dcoeff;
    a.safety = pid->safety; a.factormin = pid->factormin; a.factormax = pid->factormax;
    a.dtmin = pid->dtmin; a.dtmax = pid->dtmax;
    a.dt0 = (pid->dt0 > 0.0) ? pid->dt0 : -1.0;
    a.T = T;
    out_strides(layout, N, T, a.sn, a.sk, a.sc);
    const int block = (N >= 148LL * 128) ? 128 : ((N >= 148LL * 64) ? 64 : 32);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t err = cudaSuccess;
#define GX_LAUNCH_JOINT(C_)                                                                                   \
    do {                                                                                                      \
        void *kern = (solver == GX_SOLVER_DOPRI5) ? (void *)k_integrate_joint<C_, TabDp5>                     \
                                                  : (void *)k_integrate_joint<C_, TabDp8>;                    \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, 0);                               \
        if (per_sm < 1) per_sm = 1;                                                                           \
        long long want = (N + block - 1) / block, resident = (long long)per_sm * sms;                         \
        if (resident > JOINT_MAX_GRID) resident = JOINT_MAX_GRID;                                             \
        int grid = (int)(want < resident ? want : resident);                                                  \
        void *params[2] = {(void *)&D, (void *)&a};                                                           \
        err = cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(block), params, 0, s);                       \
    } while (0)
    if (is_basic_composite(D, model)) GX_LAUNCH_JOINT(CountsBasic);
    else GX_DISPATCH_MODEL(model, GX_LAUNCH_JOINT(C));
#undef GX_LAUNCH_JOINT
    if (err != cudaSuccess) return GX_ERR_CUDA;
    return cuda_rc(cudaGetLastError());
}

int gx_integrate_dopri8(const gx_potential *pot, const gx_pid *pid, const double *q0, const double *p0, int64_t N,
                        const double *t0, double t0_scalar, double t1, const double *ts, int32_t T, int64_t max_steps,
                        const int32_t *order, int32_t layout, double *q, double *p, int32_t *status,
                        int32_t *n_accepted, int32_t *n_attempted, void *workspace, void *stream) {
    return adaptive_impl(GX_SOLVER_DOPRI8, nullptr, nullptr, 0, pot, pid, q0, p0, N, t0, t0_scalar, t1, ts, T,
                         max_steps, order, layout, q, p, status, n_accepted, n_attempted, workspace, stream);
}

int gx_integrate_adaptive_record(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                                 const double *p0, double t0, double t1, int64_t max_steps, double *rec,
                                 int32_t rec_capacity, int32_t *n_rec, int32_t *status, int32_t *n_accepted,
                                 int32_t *n_attempted, void *workspace, void *stream) {
    if ((solver & ~GX_SOLVER_STRICT) != GX_SOLVER_DOPRI8 && (solver & ~GX_SOLVER_STRICT) != GX_SOLVER_DOPRI5) return GX_ERR_UNSUPPORTED;
    if (!rec || !n_rec || rec_capacity <= 0) return GX_ERR_BADARG;
    if (cudaMemsetAsync(n_rec, 0, sizeof(int32_t), (cudaStream_t)stream) != cudaSuccess) return GX_ERR_CUDA;
    return adaptive_impl(solver, rec, n_rec, rec_capacity, pot, pid, q0, p0, 1, nullptr, t0, t1, nullptr, 0, max_steps,
                         nullptr, GX_LAYOUT_NT3, nullptr, nullptr, status, n_accepted, n_attempted, workspace, stream);
}

int gx_integrate_dopri8_record(const gx_potential *pot, const gx_pid *pid, const double *q0, const double *p0,
                               double t0, double t1, int64_t max_steps, double *rec, int32_t rec_capacity,
                               int32_t *n_rec, int32_t *status, int32_t *n_accepted, int32_t *n_attempted,
                               void *workspace, void *stream) {
    return gx_integrate_adaptive_record(GX_SOLVER_DOPRI8, pot, pid, q0, p0, t0, t1, max_steps, rec, rec_capacity, n_rec,
                                        status, n_accepted, n_attempted, workspace, stream);
}

int gx_dense_eval_solver(int32_t solver, const double *rec, const int32_t *n_rec, double t0, double t1,
                         const double *ts, int64_t M, double *q, double *p, void *stream) {
    if (solver != GX_SOLVER_DOPRI8 && solver != GX_SOLVER_DOPRI5) return GX_ERR_UNSUPPORTED;
    if (M < 0 || (M > 0 && (!rec || !n_rec || !ts || !q || !p))) return GX_ERR_BADARG;
    if (M == 0) return 0;
    if (solver == GX_SOLVER_DOPRI5)
        k_dense_eval<TabDp5><<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>(rec, n_rec, t0, t1, ts, (long long)M, q, p);
    else
        k_dense_eval<TabDp8><<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>(rec, n_rec, t0, t1, ts, (long long)M, q, p);
    return cuda_rc(cudaGetLastError());
}

int gx_dense_eval(const double *rec, const int32_t *n_rec, double t0, double t1, const double *ts, int64_t M,
                  double *q, double *p, void *stream) {
    return gx_dense_eval_solver(GX_SOLVER_DOPRI8, rec, n_rec, t0, t1, ts, M, q, p, stream);
}

int gx_stream_release(const gx_potential *pot, int32_t df, const double *prog_q, const double *prog_p,
                      const double *prog_mass, const double *draws, int64_t M, double *q_lead, double *p_lead,
                      double *q_trail, double *p_trail, void *stream) {
    return gx_stream_release_t(pot, df, prog_q, prog_p, prog_mass, nullptr, draws, M, q_lead, p_lead, q_trail, p_trail,
                               stream);
}

int gx_stream_release_t(const gx_potential *pot, int32_t df, const double *prog_q, const double *prog_p,
                        const double *prog_mass, const double *t_release, const double *draws, int64_t M,
                        double *q_lead, double *p_lead, double *q_trail, double *p_trail, void *stream) {
    DevPot D; Model model;
    int rc = build_devpot(pot, D, model, true, t_release ? TD_INTEGRATE : TD_REJECT);
    if (rc) return rc;
    if (df != GX_DF_FARDAL15 && df != GX_DF_CHEN24) return GX_ERR_BADARG;
    if (M < 0 || (M > 0 && (!prog_q || !prog_p || !prog_mass || !draws || !q_lead || !p_lead || !q_trail || !p_trail)))
        return GX_ERR_BADARG;
    if (M == 0) return 0;
    ReleaseArgs a{prog_q, prog_p, prog_mass, draws, q_lead, p_lead, q_trail, p_trail, (long long)M, pot->G, df, t_release};
    cudaStream_t s = (cudaStream_t)stream;
    GX_DISPATCH_MODEL(model, (k_stream_release<C><<<grid_for(M, 128), 128, 0, s>>>(D, a)));
    return cuda_rc(cudaGetLastError());
}

int gx_energy_angmom(const gx_potential *pot, const double *q, const double *p, int64_t N, double *energy,
                     double *angmom, void *stream) {
    DevPot D; Model model;
    int rc = build_devpot(pot, D, model);
    if (rc) return rc;
    if (N < 0 || (N > 0 && (!q || !p))) return GX_ERR_BADARG;
    if (N == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    GX_DISPATCH_MODEL(model, (k_energy_angmom<C><<<grid_for(N, 256), 256, 0, s>>>(D, q, p, (long long)N, energy, angmom,
                                                                                 nullptr, 0.0, 1LL)));
    return cuda_rc(cudaGetLastError());
}

int gx_energy_angmom_t(const gx_potential *pot, const double *q, const double *p, int64_t N, const double *t,
                       int64_t t_period, double t_scalar, double *energy, double *angmom, void *stream) {
    DevPot D; Model model;
    int rc = build_devpot(pot, D, model, true, TD_INTEGRATE);
    if (rc) return rc;
    if (N < 0 || (N > 0 && (!q || !p)) || (t && t_period <= 0)) return GX_ERR_BADARG;
    if (N == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    GX_DISPATCH_MODEL(model, (k_energy_angmom<C><<<grid_for(N, 256), 256, 0, s>>>(D, q, p, (long long)N, energy, angmom, t,
                                                                                 t_scalar, (long long)(t ? t_period : 1))));
    return cuda_rc(cudaGetLastError());
}

int gx_bench_dfma(int32_t blocks, int32_t threads, int64_t iters, double *sink, int64_t *fma_per_thread,
                  void *stream) {
    if (blocks == 0 || threads <= 0 || threads > 1024 || !sink) return GX_ERR_BADARG;
    if (iters < 0) return GX_ERR_BADARG;
    // blocks < 0 selects the three-register-operand variant (same instruction count)
    if (blocks < 0) k_bench_dfma3<<<-blocks, threads, 0, (cudaStream_t)stream>>>((long long)iters, sink);
    else k_bench_dfma<<<blocks, threads, 0, (cudaStream_t)stream>>>((long long)iters, sink);
    if (fma_per_thread) *fma_per_thread = iters * 8 * 16;
    return cuda_rc(cudaGetLastError());
}

// jax.random.normal (threefry2x32, partitionable, float64); see include/galax_b200.h and galax_b200/jaxrandom.py.
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__global__ void __launch_bounds__(256) k_jax_normal(uint32_t k0, uint32_t k1, long long n, double *out) {
    const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t x0 = (uint32_t)((unsigned long long)i >> 32) + ks[0], x1 = (uint32_t)i + ks[1];
#pragma unroll
        for (int g = 0; g < 5; ++g) {
            const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                x0 += x1;
                x1 = rotl32(x1, R[g & 1][j]) ^ x0;
            }
            x0 += ks[(g + 1) % 3];
            x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
        }
        const unsigned long long bits = ((unsigned long long)x0 << 32) | x1;
        const double fl = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ULL)) - 1.0;
        const double lo = -0.99999999999999988898;  // nextafter(-1, 0)
        const double u = fmax(lo, __dadd_rn(__dmul_rn(fl, 1.0 - lo), lo));  // uniform(minval=lo, maxval=1)
        out[i] = 1.41421356237309514547 * erfinv(u);
    }
}

int gx_jax_normal(uint32_t key_hi, uint32_t key_lo, int64_t n, double *out, void *stream) {
    if (n < 0 || (n > 0 && !out)) return GX_ERR_BADARG;
    if (n == 0) return 0;
    const long long want = (n + 255) / 256;
    k_jax_normal<<<(int)(want < 148LL * 64 ? want : 148LL * 64), 256, 0, (cudaStream_t)stream>>>(key_hi, key_lo, (long long)n, out);
    return cuda_rc(cudaGetLastError());
}

__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t &o0,
                                                      uint32_t &o1) {
    const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
    const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    for (int g = 0; g < 5; ++g) {
        for (int j = 0; j < 4; ++j) {
            x0 += x1;
            x1 = ((x1 << R[g & 1][j]) | (x1 >> (32 - R[g & 1][j]))) ^ x0;
        }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
    o0 = x0;
    o1 = x1;
}

// draws[j][i] = normal(split(subkey_i, 4)[j], ()): key_j = threefry(subkey, counter j), bits = threefry(key_j, counter 0)
__global__ void __launch_bounds__(256) k_jax_fardal_per_key(const uint2 *subkeys, long long M, double *draws) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint2 sk = subkeys[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t kj0, kj1, b0, b1;
        threefry2x32(sk.x, sk.y, 0u, (uint32_t)j, kj0, kj1);
        threefry2x32(kj0, kj1, 0u, 0u, b0, b1);
        const unsigned long long bits = ((unsigned long long)b0 << 32) | b1;
        const double fl = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ULL)) - 1.0;
        const double lo = -0.99999999999999988898;
        const double u = fmax(lo, __dadd_rn(__dmul_rn(fl, 1.0 - lo), lo));
        draws[(long long)j * M + i] = 1.41421356237309514547 * erfinv(u);
    }
}

// the key chain itself: key, subkey = split(key) -- split(key)[c] = threefry(key, counter c) -- M times in a row.  A hash
// chain has no parallelism, so ONE device thread walks it (two independent threefry evaluations per link, ~0.15 us a
// link): slower than a host core, but the entry stays enqueue-only (no host buffer that must outlive an asynchronous
// copy, no allocation, no synchronisation) and can be captured into a graph.
__global__ void k_jax_key_chain(uint32_t k0, uint32_t k1, long long M, uint2 *subkeys) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t a = k0, b = k1;
    for (long long i = 0; i < M; ++i) {
        uint32_t s0, s1, n0, n1;
        threefry2x32(a, b, 0u, 1u, s0, s1);
        threefry2x32(a, b, 0u, 0u, n0, n1);
        subkeys[i] = make_uint2(s0, s1);
        a = n0;
        b = n1;
    }
}

int64_t gx_jax_fardal_chain_workspace_bytes(int64_t M) { return M > 0 ? M * (int64_t)sizeof(uint2) : 0; }

int gx_jax_fardal_chain(uint32_t key_hi, uint32_t key_lo, int64_t M, double *draws, void *workspace, void *stream) {
    if (M < 0 || (M > 0 && (!draws || !workspace))) return GX_ERR_BADARG;
    if (M == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    uint2 *d = (uint2 *)workspace;
    k_jax_key_chain<<<1, 32, 0, s>>>(key_hi, key_lo, (long long)M, d);
    k_jax_fardal_per_key<<<grid_for(M, 256), 256, 0, s>>>(d, (long long)M, draws);
    return cuda_rc(cudaGetLastError());
}

int gx_fixed_time_grid(double t0, double t1, double dt0, int64_t max_steps, int64_t *n_steps, int32_t *hit_max_steps,
                       int32_t *n_runs, int64_t *run_count, double *run_step, int32_t run_capacity) {
    const double dir = (t1 >= t0) ? 1.0 : -1.0;
    if (t1 != t0 && !(dt0 * dir > 0.0)) return GX_ERR_BADARG;
    FixedSeg sg;
    bool seg_ok = true;
    long long n = 0;
    int hit = 0;
    walk_time_grid(t0, t1, dt0, max_steps, sg, seg_ok, n, hit);
    if (n_steps) *n_steps = n;
    if (hit_max_steps) *hit_max_steps = hit;
    if (n_runs) *n_runs = seg_ok ? sg.n_seg : -1;
    if (seg_ok && run_count && run_step) {
        if (run_capacity < sg.n_seg) return GX_ERR_BADARG;
        for (int k = 0; k < sg.n_seg; ++k) { run_count[k] = sg.cnt[k]; run_step[k] = sg.h[k]; }
    }
    return 0;
}

int gx_force_table(int32_t which, double a, double *coef, int64_t capacity, int32_t *n_intervals, int32_t *degree,
                   int32_t *e_lo, int32_t *sub_bits, double *max_rel_err) {
    // host only: the same long-double Chebyshev fits nfw_table() / plc_table_for(a) upload to the device
    if (which != 0 && which != 1) return GX_ERR_BADARG;
    if (which == 1 && !(a > 0.0)) return GX_ERR_BADARG;
    const int nint = which == 0 ? NFW_NINT : PLC_NINT, lo = which == 0 ? NFW_E_LO : PLC_E_LO;
    if (n_intervals) *n_intervals = nint;
    if (degree) *degree = PLC_DEG;
    if (e_lo) *e_lo = lo;
    if (sub_bits) *sub_bits = PLC_SUB_BITS;
    if (!coef && !max_rel_err) return 0;
    if (coef && capacity < (int64_t)nint * (PLC_DEG + 1)) return GX_ERR_BADARG;
    std::vector<double> row(PLC_DEG + 1);
    double worst = 0.0;
    for (int j = 0; j < nint; ++j) {
        const int e = lo + j / PLC_SUB, sub = j % PLC_SUB;
        const long double base = ldexpl(1.0L, e);
        const long double s0 = base * (1.0L + sub / (long double)PLC_SUB), s1 = base * (1.0L + (sub + 1) / (long double)PLC_SUB);
        if (which == 0) fit_interval(nfw_F_ld, s0, s1, row.data(), &worst);
        else plc_fit_interval((long double)a, s0, s1, row.data(), &worst);
        if (coef) memcpy(coef + (size_t)j * (PLC_DEG + 1), row.data(), sizeof(double) * (PLC_DEG + 1));
    }
    if (max_rel_err) *max_rel_err = worst;
    return 0;
}

int gx_spherical_force_table(const gx_potential *pot, double *coef, int64_t capacity, int32_t *n_intervals,
                             int32_t *degree, int32_t *e_lo, int32_t *sub_bits, double *max_rel_err) {
    if (!pot || pot->n < 0 || pot->n > GX_MAX_COMPONENTS) return GX_ERR_BADARG;
    std::vector<SphComp> cs;
    for (int i = 0; i < pot->n; ++i) {
        const gx_component &c = pot->c[i];
        if (c.kind == GX_KIND_HERNQUIST || c.kind == GX_KIND_NFW) cs.push_back({c.kind, pot->G * c.p[0], c.p[1], 0.0});
        if (c.kind == GX_KIND_POWERLAWCUTOFF) cs.push_back({c.kind, pot->G * c.p[0], c.p[2], 1.5 - c.p[1] / 2});
    }
    if (cs.empty()) return GX_ERR_UNSUPPORTED;
    if (n_intervals) *n_intervals = SPHW_NINT;
    if (degree) *degree = SPHW_DEG;
    if (e_lo) *e_lo = sph_e_lo(cs);
    if (sub_bits) *sub_bits = SPHW_SUB_BITS;
    if (!coef && !max_rel_err) return 0;
    if (coef && capacity < (int64_t)SPHW_NINT * SPHW_ROW) return GX_ERR_BADARG;
    std::vector<double> tmp;
    if (!coef) { tmp.resize((size_t)SPHW_NINT * SPHW_ROW); coef = tmp.data(); }
    const double worst = sph_table_fit<SPHW_SUB_BITS, SPHW_ROW>(cs, coef);
    if (max_rel_err) *max_rel_err = worst;
    return 0;
}

int gx_debug_math(int32_t op, double a, const double *x, int64_t N, double *out, void *stream) {
    if (N < 0 || (N > 0 && (!x || !out))) return GX_ERR_BADARG;
    if (N == 0) return 0;
    GammaTab *gt = nullptr;
    if (op == 3) {  // debug-only: a small device allocation for the table
        GammaTab h;
        fill_gamma_tab(h, a);
        if (cudaMalloc(&gt, sizeof h) != cudaSuccess) return GX_ERR_CUDA;
        cudaMemcpyAsync(gt, &h, sizeof h, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    }
    const double *tab = nullptr;
    if (op == 5) tab = nfw_table();
    if (op == 6 || op == 7) tab = plc_table_for(a);
    if (op >= 5 && op <= 7 && tab == nullptr) return GX_ERR_UNSUPPORTED;  // the table could not be built to 1e-14
    k_debug_math<<<grid_for(N, 256), 256, 0, (cudaStream_t)stream>>>(op, gt, tab, x, (long long)N, out);
    int rc = cuda_rc(cudaGetLastError());
    if (gt) { cudaStreamSynchronize((cudaStream_t)stream); cudaFree(gt); }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// host-buffer entries
// ------------------------------------------------------------------------------------------------

struct DevBuf {
    void *p = nullptr;
    cudaError_t alloc(size_t bytes) { return bytes ? cudaMalloc(&p, bytes) : cudaSuccess; }
    ~DevBuf() { if (p) cudaFree(p); }
};

#define GX_CK(expr) do { if ((expr) != cudaSuccess) return GX_ERR_CUDA; } while (0)

int gx_host_potential_eval(const gx_potential *pot, const double *xyz, double t, int64_t N, uint32_t what,
                           double *phi, double *grad, double *acc, double *hess) {
    if (N < 0) return GX_ERR_BADARG;
    DevBuf dx, dphi, dg, da, dh;
    const size_t n = (size_t)N;
    GX_CK(dx.alloc(n * 24));
    if (what & GX_PHI) GX_CK(dphi.alloc(n * 8));
    if (what & GX_GRAD) GX_CK(dg.alloc(n * 24));
    if (what & GX_ACC) GX_CK(da.alloc(n * 24));
    if (what & GX_HESS) GX_CK(dh.alloc(n * 72));
    if (n) GX_CK(cudaMemcpy(dx.p, xyz, n * 24, cudaMemcpyHostToDevice));
    int rc = gx_potential_eval(pot, (const double *)dx.p, t, N, what, (double *)dphi.p, (double *)dg.p,
                               (double *)da.p, (double *)dh.p, nullptr);
    if (rc) return rc;
    if (n && (what & GX_PHI)) GX_CK(cudaMemcpy(phi, dphi.p, n * 8, cudaMemcpyDeviceToHost));
    if (n && (what & GX_GRAD)) GX_CK(cudaMemcpy(grad, dg.p, n * 24, cudaMemcpyDeviceToHost));
    if (n && (what & GX_ACC)) GX_CK(cudaMemcpy(acc, da.p, n * 24, cudaMemcpyDeviceToHost));
    if (n && (what & GX_HESS)) GX_CK(cudaMemcpy(hess, dh.p, n * 72, cudaMemcpyDeviceToHost));
    GX_CK(cudaDeviceSynchronize());
    return 0;
}

int gx_host_integrate_fixed(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0,
                            double t1, double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps,
                            double *q, double *p, int32_t *status) {
    if (N < 0 || T < 0) return GX_ERR_BADARG;
    const size_t n = (size_t)N, out = n * (size_t)T * 24;
    DevBuf dq0, dp0, dts, dq, dp, dst;
    GX_CK(dq0.alloc(n * 24)); GX_CK(dp0.alloc(n * 24)); GX_CK(dts.alloc((size_t)T * 8));
    GX_CK(dq.alloc(out)); GX_CK(dp.alloc(out)); GX_CK(dst.alloc(n * 4));
    if (n) { GX_CK(cudaMemcpy(dq0.p, q0, n * 24, cudaMemcpyHostToDevice)); GX_CK(cudaMemcpy(dp0.p, p0, n * 24, cudaMemcpyHostToDevice)); }
    if (T) GX_CK(cudaMemcpy(dts.p, ts, (size_t)T * 8, cudaMemcpyHostToDevice));
    int rc = gx_integrate_fixed(pot, (const double *)dq0.p, (const double *)dp0.p, N, t0, t1, dt0,
                                (const double *)dts.p, T, scheme, max_steps, GX_LAYOUT_NT3, (double *)dq.p,
                                (double *)dp.p, (int32_t *)dst.p, nullptr);
    if (rc) return rc;
    if (out) { GX_CK(cudaMemcpy(q, dq.p, out, cudaMemcpyDeviceToHost)); GX_CK(cudaMemcpy(p, dp.p, out, cudaMemcpyDeviceToHost)); }
    if (n && status) GX_CK(cudaMemcpy(status, dst.p, n * 4, cudaMemcpyDeviceToHost));
    GX_CK(cudaDeviceSynchronize());
    return 0;
}

int gx_host_integrate_dopri8(const gx_potential *pot, const gx_pid *pid, const double *q0, const double *p0,
                             int64_t N, const double *t0, double t0_scalar, double t1, const double *ts, int32_t T,
                             int64_t max_steps, double *q, double *p, int32_t *status, int32_t *n_accepted,
                             int32_t *n_attempted) {
    if (N < 0 || T < 0) return GX_ERR_BADARG;
    const size_t n = (size_t)N, out = n * (size_t)T * 24;
    DevBuf dq0, dp0, dt0, dts, dq, dp, dst, dna, dnt, dws;
    GX_CK(dq0.alloc(n * 24)); GX_CK(dp0.alloc(n * 24)); GX_CK(dts.alloc((size_t)T * 8));
    if (t0) GX_CK(dt0.alloc(n * 8));
    GX_CK(dq.alloc(out)); GX_CK(dp.alloc(out)); GX_CK(dst.alloc(n * 4)); GX_CK(dna.alloc(n * 4)); GX_CK(dnt.alloc(n * 4));
    GX_CK(dws.alloc(256));
    if (n) { GX_CK(cudaMemcpy(dq0.p, q0, n * 24, cudaMemcpyHostToDevice)); GX_CK(cudaMemcpy(dp0.p, p0, n * 24, cudaMemcpyHostToDevice)); }
    if (n && t0) GX_CK(cudaMemcpy(dt0.p, t0, n * 8, cudaMemcpyHostToDevice));
    if (T) GX_CK(cudaMemcpy(dts.p, ts, (size_t)T * 8, cudaMemcpyHostToDevice));
    int rc = gx_integrate_dopri8(pot, pid, (const double *)dq0.p, (const double *)dp0.p, N,
                                 t0 ? (const double *)dt0.p : nullptr, t0_scalar, t1, (const double *)dts.p, T,
                                 max_steps, nullptr, GX_LAYOUT_NT3, (double *)dq.p, (double *)dp.p, (int32_t *)dst.p,
                                 (int32_t *)dna.p, (int32_t *)dnt.p, dws.p, nullptr);
    if (rc) return rc;
    if (out) { GX_CK(cudaMemcpy(q, dq.p, out, cudaMemcpyDeviceToHost)); GX_CK(cudaMemcpy(p, dp.p, out, cudaMemcpyDeviceToHost)); }
    if (n && status) GX_CK(cudaMemcpy(status, dst.p, n * 4, cudaMemcpyDeviceToHost));
    if (n && n_accepted) GX_CK(cudaMemcpy(n_accepted, dna.p, n * 4, cudaMemcpyDeviceToHost));
    if (n && n_attempted) GX_CK(cudaMemcpy(n_attempted, dnt.p, n * 4, cudaMemcpyDeviceToHost));
    GX_CK(cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"
