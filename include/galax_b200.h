/* galax_b200.h -- C ABI of the B200-native galax hot path (libgalax_b200.so).
 *
 * The reference (GalacticDynamics/galax) is pure Python on JAX: it has no FFI for this path today.
 * Each entry point below replaces one piece of the reference's Python/JAX stack; the comment above it
 * cites the reference interface it stands in for (paths relative to /root/reference/src/galax/).
 * INTEGRATION.md shows the ctypes / XLA-FFI binding a galax maintainer would add.
 *
 * Conventions
 *   - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - `gx_*` entries take DEVICE pointers to contiguous row-major fp64 and only ENQUEUE work on
 *     `stream` (no synchronisation, no allocation); `gx_host_*` entries take HOST pointers, copy
 *     in/out and synchronise before returning;
 *   - return value: 0 on success, <0 on an argument / CUDA error (gx_strerror); per-particle
 *     outcomes (max_steps reached, non-finite state) are reported in the `status` array;
 *   - units: whatever unit system the potential parameters are expressed in (galax: kpc, Myr, Msun);
 *   - re-entrant and safe to call from several host threads / on several streams: the potential is passed by value
 *     into each launch (kernel-parameter constant bank).  Process-wide state: (1) an append-only, mutex-guarded cache
 *     of immutable force tables (NFW 30 KB, PowerLawCutoff 35 KB per exponent, and for the three named Milky-Way
 *     models and every composite of the four basic kinds 132 KB per distinct set of spherical-component parameters, at most 4096 sets; device memory, fitted in 12-25 ms of host time by the first integrator call that needs it,
 *     allocated and uploaded on FIRST use -- the one allocation an enqueue-only entry can make, so warm an entry up
 *     once with a potential before capturing it into a graph);
 *     (2) for the adaptive integrators, a per-device __constant__ copy of the potential that a launch reads only when
 *     no launch on another stream can still be reading a different one -- otherwise, and always under stream capture,
 *     the launch carries the potential itself (same results, a few per cent slower).  No entry ever blocks the host
 *     or synchronises a stream with work it does not depend on;
 *   - alignment: every pointer must be aligned to its element type (8 bytes for fp64); gx_potential_eval is fastest
 *     when xyz / grad / acc / hess are 16-byte aligned (TMA bulk copies) and takes a plain load/store path otherwise.
 */
#ifndef GALAX_B200_H
#define GALAX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GX_VERSION 100

/* component kinds: potential/_src/builtin/{miyamotonagai,hernquist,nfw/base,powerlawcutoff}.py */
#define GX_KIND_MIYAMOTO_NAGAI 0 /* p = (m_tot, a, b)       */
#define GX_KIND_HERNQUIST 1      /* p = (m_tot, r_s)        (r_s = 0: Kepler) */
#define GX_KIND_NFW 2            /* p = (m, r_s)            */
#define GX_KIND_POWERLAWCUTOFF 3 /* p = (m_tot, alpha, r_c) */
/* beyond the three named Milky-Way models; served by the runtime-count kernels */
#define GX_KIND_LOGARITHMIC 4 /* builtin/logarithmic.py: p = (v_c, r_s, q1, q2, q3, phi[rad]); q = 1, phi = 0: spherical */
#define GX_KIND_ISOCHRONE 5   /* builtin/isochrone.py:  p = (m_tot, r_s) */
#define GX_KIND_SATOH 6       /* builtin/satoh.py:      p = (m_tot, a, b) */
#define GX_KIND_TRIAXIAL_HERNQUIST 7 /* builtin/hernquist.py:89-176: p = (m_tot, r_s, q1, q2) */
#define GX_KIND_JAFFE 8       /* builtin/jaffe.py:      p = (m_tot, r_s) */
#define GX_KIND_BURKERT 9     /* builtin/burkert.py:    p = (m, r_s) */
#define GX_KIND_STONE 10      /* builtin/stoneostriker15.py: p = (m_tot, r_c, r_h), r_c != r_h */
#define GX_KIND_HARMONIC 11   /* builtin/example.py:23-100  HarmonicOscillatorPotential: p = (omega_x, omega_y, omega_z) */
#define GX_KIND_HENON_HEILES 12 /* builtin/example.py:107-176 HenonHeilesPotential: p = (coeff, timescale) */
#define GX_MAX_COMPONENTS 14

/* Parameters may depend linearly on time, p_k(t) = p[k] + dp[k] * t: galax's LinearParameter
 * (potential/_src/params/core.py:25-110, slope * (t - point_time) + point_value with the offset folded into p[k]).
 * dp == 0 everywhere: a static potential.  Time-dependent composites are supported by gx_potential_eval (frozen at
 * its t) and by the integrators for the kinds MIYAMOTO_NAGAI, HERNQUIST, NFW, ISOCHRONE, SATOH, TRIAXIAL_HERNQUIST
 * and JAFFE (each right-hand side is evaluated with the parameters of its own stage time), by gx_stream_release_t (each
 * stripping time sees the potential of its own time) and gx_energy_angmom_t; the entries without a time argument,
 * gx_stream_release and gx_energy_angmom, return GX_ERR_UNSUPPORTED for them. */
typedef struct {
    int32_t kind;
    int32_t reserved; /* summation group: consecutive components with the same non-zero value were ONE reference
                         component (the three Miyamoto-Nagai terms of an MN3 disk) and GX_SCHEME_STRICT sums them first,
                         as the reference does; 0 = a component of its own.  Ignored by the default kernels. */
    double p[8];
    double dp[8];
} gx_component;

/* A composite potential = sum of components (potential/_src/base_multi.py:39-82).  MN3 disks are passed as
 * their three Miyamoto-Nagai components (builtin/mn3.py:90-119).  G = pot.constants["G"].value. */
typedef struct {
    int32_t n;
    int32_t reserved;
    double G;
    gx_component c[GX_MAX_COMPONENTS];
} gx_potential;

/* `what` bit mask for gx_potential_eval */
#define GX_PHI 1u  /* potential/_src/base.py:113-136  _potential            -> phi[N]      */
#define GX_GRAD 2u /* potential/_src/base.py:170-179  _gradient             -> grad[N,3]   */
#define GX_ACC 4u  /* potential/_src/register_funcs.py:327-340 acceleration -> acc[N,3]    */
#define GX_HESS 8u /* potential/_src/base.py:230-239  _hessian              -> hess[N,3,3] */

/* per-particle status */
#define GX_OK 0
#define GX_MAX_STEPS_REACHED 1 /* diffrax RESULTS.max_steps_reached */
#define GX_NONFINITE 2         /* state became inf/nan              */

/* error codes */
#define GX_ERR_BADARG (-1)
#define GX_ERR_UNSUPPORTED (-2)
#define GX_ERR_CUDA (-3)

/* fixed-step schemes */
#define GX_SCHEME_SEMI_IMPLICIT_EULER 0 /* diffrax.SemiImplicitEuler (the reference's "leapfrog") */
#define GX_SCHEME_LEAPFROG_MIDPOINT 1   /* diffrax.LeapfrogMidpoint                               */
/* or-ed into `scheme`: use the general fixed-step kernel (time arithmetic every step) where the run-length kernel
 * would be chosen.  Results are bit-identical; for tests and A/B timing. */
#define GX_SCHEME_GENERAL_KERNEL 0x100
/* or-ed into `scheme`: reference-order arithmetic (galax_b200/csrc/gx_strict.cu) -- every component evaluated and
 * summed in composite order as potential/_src/base_multi.py:48-55 does (gx_component.reserved = summation group, see
 * below), IEEE division / square root, no FMA contraction, portable fdlibm log1p / exp / log
 * (include/gx_portable_math.h), the update as a multiply and an add.  Bit for bit what a plain C program computes on
 * an IEEE-754 CPU; several times slower than the default kernels.  Static composites of MIYAMOTO_NAGAI, HERNQUIST,
 * NFW and POWERLAWCUTOFF components only (GX_ERR_UNSUPPORTED otherwise). */
#define GX_SCHEME_STRICT 0x200

/* output layout of saved states, element (particle n, save k, component c) */
#define GX_LAYOUT_NT3 0 /* [N,T,3]  the reference's (*batch, T, 3), orbit/register_dfx.py:80-82 */
#define GX_LAYOUT_T3N 1 /* [T,3,N]  structure-of-arrays: fully coalesced stores                 */

/* Step-size controller = diffrax.PIDController as galax constructs it
 * (dynamics/_src/legacy/integrator.py:37-39,161-168; dynamics/_src/orbit/solver.py:121-141). */
typedef struct {
    double rtol, atol;
    double pcoeff, icoeff, dcoeff;       /* defaults 0, 1, 0 */
    double safety, factormin, factormax; /* defaults 0.9, 0.2, 10 */
    double dtmin, dtmax;                 /* <= 0: unset */
    int32_t force_dtmin;
    int32_t reserved;
    double dt0; /* <= 0 or NaN: Hairer initial-step heuristic (diffrax dt0=None) */
} gx_pid;

const char *gx_strerror(int code);
int gx_version(void);

/* Bulk evaluation.  Replaces pot.potential/gradient/acceleration/hessian on (N,3) arrays:
 * potential/_src/register_funcs.py:33-98,276-288,327-340 -> AbstractCompositePotential._gradient/_hessian
 * (base_multi.py:48-82).  Unrequested outputs may be NULL.  Parameters with a rate (LinearParameter, gx_component.dp)
 * are evaluated at `t`; constant parameters ignore it. */
int gx_potential_eval(const gx_potential *pot, const double *xyz, double t, int64_t N, uint32_t what, double *phi,
                      double *grad, double *acc, double *hess, void *stream);

/* Fixed-step integration.  Replaces
 *   OrbitSolver(solver=dfx.SemiImplicitEuler(), stepsize_controller=dfx.ConstantStepSize())
 *       .solve(HamiltonianField(pot), (q0, p0), t0, t1, dt0=dt0, max_steps=..., saveat=ts)
 * (dynamics/_src/orbit/field_hamiltonian.py:256-301, dynamics/_src/orbit/solver.py:774-803 -> diffrax.diffeqsolve).
 * q0,p0: [N,3]; ts: [T] device array of save times inside [t0,t1] (ascending in the direction of integration);
 * q,p: saved states in `layout`; status: [N] int32 (may be NULL); max_steps < 0 = unbounded.
 * SemiImplicitEuler with constant parameters runs on the run-length form of the time grid (gx_fixed_time_grid below):
 * same results bit for bit as the step-by-step kernel, which GX_SCHEME_GENERAL_KERNEL selects. */
int gx_integrate_fixed(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0, double t1,
                       double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps, int32_t layout,
                       double *q, double *p, int32_t *status, void *stream);

/* Host only (no CUDA call): the time grid gx_integrate_fixed walks for (t0, t1, dt0, max_steps) -- diffrax's
 * ConstantStepSize grid t_{n+1} = fl(t_n + dt0) with the last step clipped to t1 (_clip_to_end, 1e-10) -- as a trip
 * count and in run-length form: run k is run_count[k] steps of the exactly representable size run_step[k] (in the
 * direction of integration; t_s + j * run_step[k] reproduces the grid times exactly inside a run).  This is what the
 * fixed-step kernel for MilkyWayPotential / MilkyWayPotential2022 / BovyMWPotential2014 consumes; n_runs = -1 means the
 * grid has more than 120 runs (or a step that is not exactly representable) and the general kernel is used.
 * Any output pointer may be NULL. */
int gx_fixed_time_grid(double t0, double t1, double dt0, int64_t max_steps, int64_t *n_steps, int32_t *hit_max_steps,
                       int32_t *n_runs, int64_t *run_count, double *run_step, int32_t run_capacity);

/* Adaptive integration, per-particle step control.  Replaces
 *   OrbitSolver(dfx.Dopri8(), stepsize_controller=dfx.PIDController(rtol, atol)).solve(lstrat.VMap, field, ...)
 * and the per-particle solves of Integrator/evaluate_orbit (dynamics/_src/legacy/integrator.py:179-245,449-527;
 * dynamics/_src/legacy/funcs.py:186-213).
 * t0: [N] device array of per-particle start times, or NULL to use the scalar t0_scalar.
 * order: optional [N] int32 processing order (e.g. sorted by expected step count); NULL = identity.
 * n_accepted / n_attempted: [N] int32 step counters (may be NULL).
 * workspace: device buffer of gx_workspace_bytes() bytes (work-queue ticket), zeroed by the call. */
int gx_integrate_dopri8(const gx_potential *pot, const gx_pid *pid, const double *q0, const double *p0, int64_t N,
                        const double *t0, double t0_scalar, double t1, const double *ts, int32_t T, int64_t max_steps,
                        const int32_t *order, int32_t layout, double *q, double *p, int32_t *status,
                        int32_t *n_accepted, int32_t *n_attempted, void *workspace, void *stream);
int64_t gx_workspace_bytes(void);

/* Single orbit with many save times (the progenitor orbit of MockStreamGenerator.run, mockstream_generator.py:239:
 * one adaptive solve saved at all M stripping times).  gx_integrate_dopri8_record integrates ONE particle and
 * records every accepted step (GX_DENSE_RECORD_DOUBLES doubles each) instead of saving; gx_dense_eval then evaluates
 * the solver's dense output at M times in parallel.  Numerically identical to gx_integrate_dopri8 with saveat = ts.
 * rec: device buffer of rec_capacity * GX_DENSE_RECORD_DOUBLES doubles; n_rec: device int32 (records written);
 * status is GX_MAX_STEPS_REACHED when the buffer was too small. */
#define GX_DENSE_RECORD_DOUBLES 51
int gx_integrate_dopri8_record(const gx_potential *pot, const gx_pid *pid, const double *q0, const double *p0,
                               double t0, double t1, int64_t max_steps, double *rec, int32_t rec_capacity,
                               int32_t *n_rec, int32_t *status, int32_t *n_accepted, int32_t *n_attempted,
                               void *workspace, void *stream);
int gx_dense_eval(const double *rec, const int32_t *n_rec, double t0, double t1, const double *ts, int64_t M,
                  double *q, double *p, void *stream);

/* The same three entries with the explicit Runge-Kutta pair selectable: GX_SOLVER_DOPRI8 = diffrax.Dopri8 (galax's
 * default everywhere except) GX_SOLVER_DOPRI5 = diffrax.Dopri5, the default of the experimental StreamSimulator
 * (dynamics/_src/experimental/stream.py:32-41).  gx_integrate_dopri8* are these with solver = GX_SOLVER_DOPRI8. */
#define GX_SOLVER_DOPRI8 8
#define GX_SOLVER_DOPRI5 5
/* or-ed into `solver` of gx_integrate_adaptive: reference-order arithmetic, the adaptive counterpart of GX_SCHEME_STRICT
 * (galax_b200/csrc/gx_strict.cu) -- the generic six-component Runge-Kutta form with k_i = f(.) h, components summed in
 * composite order, IEEE division / square root, no FMA contraction, portable log1p / exp / log / pow, one thread per
 * particle.  Results AND the accept / reject sequence are bit for bit what a plain C program computes on an IEEE-754
 * CPU (tests/test_gpu_strict.py: every particle of a 4096-particle C2-shaped run equals oracle/galax_oracle.c at
 * all 1000 saves).  Same component kinds as GX_SCHEME_STRICT; `order`, `workspace` are ignored; not for *_record. */
#define GX_SOLVER_STRICT 0x200
int gx_integrate_adaptive(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                          const double *p0, int64_t N, const double *t0, double t0_scalar, double t1, const double *ts,
                          int32_t T, int64_t max_steps, const int32_t *order, int32_t layout, double *q, double *p,
                          int32_t *status, int32_t *n_accepted, int32_t *n_attempted, void *workspace, void *stream);
int gx_integrate_adaptive_record(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                                 const double *p0, double t0, double t1, int64_t max_steps, double *rec,
                                 int32_t rec_capacity, int32_t *n_rec, int32_t *status, int32_t *n_accepted,
                                 int32_t *n_attempted, void *workspace, void *stream);
int gx_dense_eval_solver(int32_t solver, const double *rec, const int32_t *n_rec, double t0, double t1,
                         const double *ts, int64_t M, double *q, double *p, void *stream);

/* The reference's JOINT batch semantics: ONE adaptive solve of the 6N-dimensional system with one shared step.  This is
 * what evaluate_orbit(pot, w0[N,6], t) / OrbitSolver.solve(field, (q[N,3], p[N,3]), t0, t1) do for scalar times
 * (dynamics/_src/orbit/solver.py:774-803, dynamics/_src/legacy/integrator.py:288-298 -> one diffrax.diffeqsolve on the
 * (N,3) pytree: the error norm is the RMS over all 6N numbers, every particle takes the same steps), whereas
 * gx_integrate_adaptive controls the step per particle (the reference under vmap / lstrat.VMap / batched t0).  The two
 * agree to the tolerance, not to rounding; this entry gives the reference's numbers for that call form.  One
 * cooperative launch (the grid must be resident: at most 2048 CTAs walk the batch), reductions in a fixed order, so
 * results are reproducible run to run (a cooperative launch: enqueue-only like the other entries, but do not count on
 * capturing it into a CUDA graph).  status / n_accepted / n_attempted: ONE int32 each (the solve is one ODE).
 * workspace: gx_joint_workspace_bytes(N) bytes.  Any potential the per-particle entry accepts, LinearParameter included. */
int64_t gx_joint_workspace_bytes(int64_t N);
int gx_integrate_adaptive_joint(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                                const double *p0, int64_t N, double t0, double t1, const double *ts, int32_t T,
                                int64_t max_steps, int32_t layout, double *q, double *p, int32_t *status,
                                int32_t *n_accepted, int32_t *n_attempted, void *workspace, void *stream);

/* Stream release (distribution function).  Replaces FardalStreamDF._sample / ChenStreamDF._sample given the random
 * draws (dynamics/_src/legacy/mockstream/df/fardal15.py:49-94, df/chen24.py:61-137; tidal radius
 * dynamics/_src/cluster/radius.py:198-215; omega dynamics/_src/register_api.py:77-88).
 * prog_q, prog_p: [M,3] progenitor orbit at the stripping times; prog_mass: [M];
 * draws: Fardal [4,M] standard normals (kr, kvphi, kz, kvz); Chen [M,6] multivariate-normal samples.
 * outputs: [M,3] each. */
#define GX_DF_FARDAL15 0
#define GX_DF_CHEN24 1
int gx_stream_release(const gx_potential *pot, int32_t df, const double *prog_q, const double *prog_p,
                      const double *prog_mass, const double *draws, int64_t M, double *q_lead, double *p_lead,
                      double *q_trail, double *p_trail, void *stream);

/* The same with the release times: every stripping time sees the potential of ITS OWN time, as FardalStreamDF._sample /
 * tidal_radius(pot, x, v, mass=..., t=t) do (df/fardal15.py:49-94, cluster/radius.py:198-215).  Needed for composites
 * with LinearParameter rates (the kinds the integrators accept); t_release: device [M].  t_release = NULL is
 * gx_stream_release (static potentials only). */
int gx_stream_release_t(const gx_potential *pot, int32_t df, const double *prog_q, const double *prog_p,
                        const double *prog_mass, const double *t_release, const double *draws, int64_t M,
                        double *q_lead, double *p_lead, double *q_trail, double *p_trail, void *stream);

/* Derived diagnostics on device: E = |p|^2/2 + Phi(q) and L = q x p for [N,3] states (used for the energy-drift
 * report; coordinates/_src/pscs/base.py total_energy / angular_momentum). */
int gx_energy_angmom(const gx_potential *pot, const double *q, const double *p, int64_t N, double *energy,
                     double *angmom, void *stream);
/* ... with the time of each state, for LinearParameter composites: E_i = |p_i|^2/2 + Phi(q_i, t[i mod t_period]) (an
 * orbit batch [B,T,3] flattened to N = B T states passes its T save times and t_period = T), or Phi(., t_scalar) when
 * t is NULL. */
int gx_energy_angmom_t(const gx_potential *pot, const double *q, const double *p, int64_t N, const double *t,
                       int64_t t_period, double t_scalar, double *energy, double *angmom, void *stream);

/* Orbit post-processing fused into the integrators: the specific total energy E = |p|^2/2 + Phi(q)
 * (coordinates/_src/pscs/base.py:231-283 total_energy = kinetic + potential_energy), the angular momentum L = q x p
 * (:285-330) and the tidal tensor J - tr(J)/3 I of the potential's Hessian J (potential/_src/register_funcs.py:347-377)
 * AT EVERY SAVED STATE, computed while the state is in registers and written with the same sector-aligned staging as q
 * and p -- no second pass over the saved orbit.  Each pointer may be NULL (not wanted).  Shapes follow `layout`:
 * GX_LAYOUT_NT3: energy [N,T], angmom [N,T,3], tidal [N,T,3,3]; GX_LAYOUT_T3N: [T,N], [T,3,N], [T,9,N].  Saves a
 * particle never reached are NaN, as in q and p.  Static potentials only (GX_ERR_UNSUPPORTED for LinearParameter
 * composites and for the reference-order GX_SCHEME_STRICT / GX_SOLVER_STRICT kernels). */
typedef struct {
    double *energy;
    double *angmom;
    double *tidal;
} gx_orbit_epilogue;
int gx_integrate_fixed_epilogue(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0,
                                double t1, double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps,
                                int32_t layout, double *q, double *p, int32_t *status, const gx_orbit_epilogue *epi,
                                void *stream);
int gx_integrate_adaptive_epilogue(int32_t solver, const gx_potential *pot, const gx_pid *pid, const double *q0,
                                   const double *p0, int64_t N, const double *t0, double t0_scalar, double t1,
                                   const double *ts, int32_t T, int64_t max_steps, const int32_t *order, int32_t layout,
                                   double *q, double *p, int32_t *status, int32_t *n_accepted, int32_t *n_attempted,
                                   void *workspace, const gx_orbit_epilogue *epi, void *stream);

/* ---- host-buffer convenience entries (allocate, copy in, run, copy out, synchronise) ---- */
int gx_host_potential_eval(const gx_potential *pot, const double *xyz, double t, int64_t N, uint32_t what,
                           double *phi, double *grad, double *acc, double *hess);
int gx_host_integrate_fixed(const gx_potential *pot, const double *q0, const double *p0, int64_t N, double t0,
                            double t1, double dt0, const double *ts, int32_t T, int32_t scheme, int64_t max_steps,
                            double *q, double *p, int32_t *status);
int gx_host_integrate_dopri8(const gx_potential *pot, const gx_pid *pid, const double *q0, const double *p0,
                             int64_t N, const double *t0, double t0_scalar, double t1, const double *ts, int32_t T,
                             int64_t max_steps, double *q, double *p, int32_t *status, int32_t *n_accepted,
                             int32_t *n_attempted);

/* ---- measurement helpers ---- */
/* FP64 FMA-pipe peak: every thread runs `iters` rounds of 8 independent DFMA chains.  Returns the number of
 * DFMA instructions issued per thread (iters*8*unroll) via *fma_per_thread; time it with events on `stream`.
 * blocks > 0: x <- x*m + c with m, c constants (one register operand; the textbook peak);
 * blocks < 0: |blocks| CTAs of x <- x*y + z with three distinct register operands (what real code looks like). */
int gx_bench_dfma(int32_t blocks, int32_t threads, int64_t iters, double *sink, int64_t *fma_per_thread, void *stream);
/* jax.random.normal(key, (n,), float64) with jax's default threefry2x32 generator in partitionable mode (jax >= 0.5;
 * the reference pins jax 0.8.0): element i = sqrt(2) erfinv(u_i), u_i from the 64 random bits threefry2x32(key, i).
 * key = raw key data (hi, lo) as returned by jax.random.key_data.  Replaces the draws of
 * legacy/mockstream/df/fardal15.py:61,81-84 (four such calls on jr.split(key, 4)) and df/chen24.py:89-91 (one call of
 * shape (M, 6), then the SVD factor) so that a seeded stream needs no host-side random numbers.  Integer part exact;
 * erfinv is CUDA's (XLA's differs at the 1e-16 level). */
int gx_jax_normal(uint32_t key_hi, uint32_t key_lo, int64_t n, double *out, void *stream);
/* The draws of the experimental StreamSimulator.init scan (experimental/stream.py:212-226, experimental/df.py:146-163):
 * `for i in range(M): key, subkey = jr.split(key); Fardal2015DF.sample(subkey, ...)`, each sample being four scalar
 * normals `jr.normal(k_j, ())` on `jr.split(subkey, 4)`.  The key chain is inherently sequential: one device thread
 * walks it (~0.15 us per link) into `workspace`, then the 4 M normals are computed in parallel.  Enqueue-only like every
 * gx_* entry.  draws: device [4][M]; workspace: device buffer of gx_jax_fardal_chain_workspace_bytes(M) bytes. */
int64_t gx_jax_fardal_chain_workspace_bytes(int64_t M);
int gx_jax_fardal_chain(uint32_t key_hi, uint32_t key_lo, int64_t M, double *draws, void *workspace, void *stream);
/* Host only (no CUDA call): the piecewise-polynomial force tables the integrators stage in shared memory, as fitted on
 * the host -- which = 0: NFW F(s) = (ln(1+s) - s/(1+s)) / s^3; which = 1: PowerLawCutoff G(s) = P(a, s^2) / s^3 for
 * the exponent a = 3/2 - alpha/2.  Rows of (degree + 1) monomial coefficients in t in [-1, 1) per interval,
 * 2^sub_bits intervals per octave of s starting at 2^e_lo; max_rel_err is the fit's own check against the long-double
 * function.  coef may be NULL (only the layout / the error are wanted). */
int gx_force_table(int32_t which, double a, double *coef, int64_t capacity, int32_t *n_intervals, int32_t *degree,
                   int32_t *e_lo, int32_t *sub_bits, double *max_rel_err);
/* Host only (no CUDA call): the COMBINED spherical force table of a composite, S(u) = sum over its Hernquist / NFW /
 * PowerLawCutoff components of Phi_i'(r)/r as a function of u = r^2 -- what the integrators of the three named
 * Milky-Way models -- and of any other composite of the four basic kinds -- look up instead of evaluating the spherical
 * components: 128 intervals per octave of u over 22 octaves placed by the scale radii, degree 5, rows of 6 doubles
 * (48 bytes = three 16-byte loads per lookup: those kernels are bound by the shared-memory port), 132 KB held in shared
 * memory by one CTA per SM.  Fitted per potential by the first integrator call that needs it, cached per device.
 * coef holds the monomial coefficients in t in [-1, 1) across each interval, as fitted; on the device row j's
 * coefficient k is stored times 2^(-k (e_j - sub_bits - 1)) (e_j = the octave's exponent), i.e. as a polynomial in
 * u - centre_j, an exact difference: the same value bit for bit, three integer instructions fewer per lookup.
 * GX_ERR_UNSUPPORTED if the potential has no spherical component of these kinds. */
int gx_spherical_force_table(const gx_potential *pot, double *coef, int64_t capacity, int32_t *n_intervals,
                             int32_t *degree, int32_t *e_lo, int32_t *sub_bits, double *max_rel_err);
/* elementwise math probes for the tests: op 0 rcp, 1 rsqrt, 2 log1p, 3 gammainc_P(a, x), 4 NFW shape ln(1+s) - s/(1+s),
 * 5 NFW force table F(s) = shape / s^3, 6 / 7 PowerLawCutoff table G(s) = P(a, s^2) / s^3 and dG/ds (NaN outside the
 * tabulated range) */
int gx_debug_math(int32_t op, double a, const double *x, int64_t N, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GALAX_B200_H */
