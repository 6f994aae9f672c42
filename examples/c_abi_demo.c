/* Plain-C caller of libgalax_b200.so: no CUDA headers, no Python, host buffers in and out.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lgalax_b200 -lgalax_b200 -Wl,-rpath,$PWD/galax_b200 -lm -o c_abi_demo
 *
 * MilkyWayPotential (builtin/milkyway.py:203-236): the reference's known-answer point [1, 2, 3] kpc
 * (tests/unit/potential/builtin/test_milkywaypotential.py:41-60), then 16 orbits through the fixed-step and the
 * Dopri8 integrators.  Prints one line per result; tests/test_gpu_c_abi.py compares them with the Python mirror. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "galax_b200.h"

static gx_potential milky_way(void) {
    gx_potential P;
    memset(&P, 0, sizeof P);
    P.G = 4.498502151469553e-12; /* pot.constants["G"].value in galactic units */
    P.n = 4;
    P.c[0].kind = GX_KIND_MIYAMOTO_NAGAI; P.c[0].p[0] = 6.8e10;  P.c[0].p[1] = 3.0;   P.c[0].p[2] = 0.28; /* disk */
    P.c[1].kind = GX_KIND_NFW;            P.c[1].p[0] = 5.4e11;  P.c[1].p[1] = 15.62;                      /* halo */
    P.c[2].kind = GX_KIND_HERNQUIST;      P.c[2].p[0] = 5e9;     P.c[2].p[1] = 1.0;                        /* bulge */
    P.c[3].kind = GX_KIND_HERNQUIST;      P.c[3].p[0] = 1.71e9;  P.c[3].p[1] = 0.07;                       /* nucleus */
    return P;
}

int main(void) {
    gx_potential P = milky_way();
    printf("version %d\n", gx_version());

    const double x[3] = {1.0, 2.0, 3.0};
    double phi, grad[3], hess[9];
    int rc = gx_host_potential_eval(&P, x, 0.0, 1, GX_PHI | GX_GRAD | GX_HESS, &phi, grad, NULL, hess);
    if (rc) { printf("gx_host_potential_eval: %s\n", gx_strerror(rc)); return 1; }
    printf("phi %.12e\n", phi);
    printf("grad %.12e %.12e %.12e\n", grad[0], grad[1], grad[2]);
    printf("hess_diag %.12e %.12e %.12e\n", hess[0], hess[4], hess[8]);

    enum { N = 16 };
    double q0[N][3], p0[N][3], q[N][3], p[N][3];
    int32_t status[N];
    for (int i = 0; i < N; ++i) {
        double r = 4.0 + i, ang = 0.37 * i;
        q0[i][0] = r * cos(ang); q0[i][1] = r * sin(ang); q0[i][2] = 0.1 * i - 0.5;
        p0[i][0] = -0.2 * sin(ang); p0[i][1] = 0.2 * cos(ang); p0[i][2] = 0.01 * i;
    }
    const double t1 = 1000.0;
    rc = gx_host_integrate_fixed(&P, &q0[0][0], &p0[0][0], N, 0.0, t1, 0.1, &t1, 1, GX_SCHEME_SEMI_IMPLICIT_EULER, -1,
                                 &q[0][0], &p[0][0], status);
    if (rc) { printf("gx_host_integrate_fixed: %s\n", gx_strerror(rc)); return 1; }
    for (int i = 0; i < N; i += 5) printf("sie %d %d %.12e %.12e %.12e\n", i, status[i], q[i][0], q[i][1], q[i][2]);

    gx_pid pid;
    memset(&pid, 0, sizeof pid);
    pid.rtol = pid.atol = 1e-8; /* OrbitSolver defaults (orbit/solver.py:121-141) */
    pid.icoeff = 1.0; pid.safety = 0.9; pid.factormin = 0.2; pid.factormax = 10.0;
    pid.dtmin = -1.0; pid.dtmax = -1.0; pid.force_dtmin = 1; pid.dt0 = -1.0;
    int32_t nacc[N], ntot[N];
    rc = gx_host_integrate_dopri8(&P, &pid, &q0[0][0], &p0[0][0], N, NULL, 0.0, t1, &t1, 1, 65536, &q[0][0], &p[0][0],
                                  status, nacc, ntot);
    if (rc) { printf("gx_host_integrate_dopri8: %s\n", gx_strerror(rc)); return 1; }
    for (int i = 0; i < N; i += 5)
        printf("dopri8 %d %d %d %d %.10e %.10e %.10e\n", i, status[i], nacc[i], ntot[i], q[i][0], q[i][1], q[i][2]);
    return 0;
}
