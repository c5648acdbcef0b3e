/*
 * bratu2d_newton.c — a plain C consumer of libariadne_b200.so (no Python, no torch, no Julia).
 *
 * Solves the 2-D Bratu problem (tensor extension of examples/bratu.jl:14-24 on the grid of examples/heat_2D.jl) with
 *   newton_krylov!(bratu!, u0, (dx, dy, lambda), res; algo = :gmres)              src/Ariadne.jl:288-372
 * through the same entry points the Julia `ccall` wrapper binds, and prints the Newton history.
 *
 *   cc -O2 -Iinclude examples/c/bratu2d_newton.c -o bratu2d_newton \
 *      -Lnewtonkrylov.jl_b200 -lariadne_b200 -Wl,-rpath,$PWD/newtonkrylov.jl_b200 -lm
 *   ./bratu2d_newton [N=256] [lambda=3.5]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ariadne_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ < 0) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ak_last_error());    \
            return 1;                                                                \
        }                                                                            \
    } while (0)

static void print_step(void* user, const double* u_dev, const double* res_dev, double n_res) {
    int* k = (int*)user; /* callback(u, res, n_res): src/Ariadne.jl:304,351 */
    (void)u_dev;
    (void)res_dev;
    printf("  newton step %2d  ||F|| = %.12e\n", (*k)++, n_res);
}

int main(int argc, char** argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 256;
    const double lambda = argc > 2 ? atof(argv[2]) : 3.5;
    const int64_t n = N * N;
    const double dx = 1.0 / (double)(N + 1);
    const double pi = 3.14159265358979323846;

    if (ak_abi_version() != AK_ABI_VERSION) {
        fprintf(stderr, "header / library ABI mismatch\n");
        return 1;
    }
    ak_ctx* ctx = NULL;
    CHECK(ak_ctx_create(0, &ctx));

    double* u0 = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t j = 0; j < N; ++j)
        for (int64_t i = 0; i < N; ++i) u0[j * N + i] = sin(pi * dx * (double)(i + 1)) * sin(pi * dx * (double)(j + 1));

    double *u = NULL, *res = NULL, *coef = NULL;
    CHECK(ak_malloc(ctx, n, &u));
    CHECK(ak_malloc(ctx, n, &res));
    CHECK(ak_malloc(ctx, n, &coef));
    CHECK(ak_upload(ctx, u, u0, n));
    CHECK(ak_fill(ctx, n, res, 0.0));

    ak_problem p;
    memset(&p, 0, sizeof(p));
    p.kind = AK_BRATU2D;
    p.scheme = AK_STEADY;
    p.nx = N; p.ny = N; p.gny = N; p.gy0 = 0;
    p.dx = dx; p.dy = dx; p.lambda = lambda;
    p.coef = coef; /* lambda*exp(u) cached by the residual for the JVPs of the same Newton step */

    ak_newton_opts o;
    ak_newton_default_opts(&o); /* tol_rel 1e-6, tol_abs 1e-12, max_niter 50, Eisenstat-Walker, GMRES, memory 20 */
    ak_newton_stats st;
    int step = 0;
    printf("2-D Bratu %lld x %lld, lambda = %g\n", (long long)N, (long long)N, lambda);
    CHECK(ak_newton_solve(ctx, &p, u, res, &o, &st, NULL, NULL, NULL, 0, print_step, &step));
    printf("solved = %d  outer = %d  inner = %lld  ||F|| = %.6e  t = %.3f s  kernels = %lld\n", st.solved,
           st.outer_iterations, (long long)st.inner_iterations, st.n_res, st.t_seconds,
           (long long)ak_ctx_launch_count(ctx, 0));

    CHECK(ak_download(ctx, u0, u, n));
    double umax = 0.0;
    for (int64_t i = 0; i < n; ++i) umax = u0[i] > umax ? u0[i] : umax;
    printf("max u = %.12f\n", umax);

    ak_free(ctx, u); ak_free(ctx, res); ak_free(ctx, coef);
    free(u0);
    ak_ctx_destroy(ctx);
    return st.solved ? 0 : 2;
}
