// Newton outer loop, forcing policy and the implicit time stepper (host-side scalar logic
// that enqueues device work; only ||F|| and the Krylov verdict come back per Newton step).
//
//   ak_newton_solve     newton_krylov!          src/Ariadne.jl:288-372
//   ak_forcing_ew       EisenstatWalker         src/Ariadne.jl:197-217
//   ak_implicit_solve   solve(G!, f!, ...)      examples/implicit.jl:54-78
#include <math.h>
#include <string.h>
#include <time.h>

#include "ak_internal.h"

namespace ak {
int krylov_solve_internal(ak_krylov* ws, const ak_problem* p, const double* u, const double* b,
                          const ak_krylov_opts* opts, ak_krylov_stats* st, double* hist_host, int64_t hist_cap);

static double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int residual_norm(Ctx* c, const ak_problem* p, double* u, double* res, double* n_res) {
    AK_TRY(launch_residual(c, p, u, res, c->dscal));
    AK_CUDA(cudaMemcpyAsync(c->hscal, c->dscal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    AK_CUDA(cudaStreamSynchronize(c->stream));
    *n_res = sqrt(c->hscal[0]);
    return AK_OK;
}

// Newton loop on an existing workspace (`rhs` = device scratch of n doubles for copy(res)).
static int newton_ws(Ctx* c, const ak_problem* p, double* u, double* res, const ak_newton_opts* o, ak_krylov* ws,
                     double* rhs, ak_newton_stats* st, double* hist_nres, int64_t* hist_inner, double* hist_eta,
                     int32_t hist_cap, ak_newton_callback cb, void* cb_user) {
    const int64_t n = ak_problem_size(p);
    const double t0 = now_s();
    double n_res = 0.0;
    AK_TRY(residual_norm(c, p, u, res, &n_res));  // :302-303
    if (cb) cb(cb_user, u, res, n_res);           // :304
    int32_t nh = 0;
    auto record = [&](double nr, int64_t inner, double eta_used) {
        if (nh < hist_cap) {
            if (hist_nres) hist_nres[nh] = nr;
            if (hist_inner) hist_inner[nh] = inner;
            if (hist_eta) hist_eta[nh] = eta_used;
        }
        nh++;
    };
    record(n_res, 0, 0.0);
    const double tol = o->tol_rel * n_res + o->tol_abs;  // :306, computed once
    double eta = 0.0;
    if (o->forcing == AK_FORCING_FIXED) eta = o->eta;         // inital(Fixed)          :192
    else if (o->forcing == AK_FORCING_EW) eta = o->eta_max;   // inital(EisenstatWalker) :217
    int32_t outer = 0, flags = 0;
    int64_t inner = 0;
    if (o->verbose > 0)
        fprintf(stderr, "[ariadne_b200] Jacobian-Free Newton-Krylov res0=%.16e tol=%.6e eta=%g\n", n_res, tol, eta);
    while (n_res > tol && outer <= o->max_niter) {  // :321 (admits max_niter+1 steps)
        ak_krylov_opts ko = o->krylov;
        if (o->forcing != AK_FORCING_NONE && !o->krylov_rtol_override) ko.rtol = eta;  // :330-333
        AK_TRY(launch_copy(c, n, rhs, res));  // copy(res) :338
        ak_krylov_stats ks;
        int krc = krylov_solve_internal(ws, p, u, rhs, &ko, &ks, nullptr, 0);
        if (krc < 0) return krc;
        AK_TRY(launch_axpy(c, n, -1.0, ak_krylov_x(ws), u));  // u .-= s .* d, s = 1  :340-344
        const double n_res_prior = n_res;
        AK_TRY(residual_norm(c, p, u, res, &n_res));  // :349-350
        if (cb) cb(cb_user, u, res, n_res);           // :351
        const double eta_used = eta;
        if (isinf(n_res) || isnan(n_res)) {  // :353-356
            flags |= AK_FLAG_NAN;
            if (o->verbose > 0) fprintf(stderr, "[ariadne_b200] Inner solver blew up\n");
            break;
        }
        if (o->forcing == AK_FORCING_EW)
            eta = ak_forcing_ew(o->eta_max, o->gamma, eta, tol, n_res, n_res_prior);  // :358-360
        outer += 1;  // update(stats, ...) :367
        inner += ks.niter;
        record(n_res, ks.niter, eta_used);
        if (o->verbose > 0)
            fprintf(stderr, "[ariadne_b200] Newton iter=%d n_res=%.16e eta=%g inner=%lld\n", outer, n_res, eta,
                    (long long)ks.niter);
    }
    st->solved = n_res <= tol;
    st->outer_iterations = outer;
    st->inner_iterations = inner;
    st->n_res = n_res;
    st->tol = tol;
    st->t_seconds = now_s() - t0;
    st->flags = flags | (st->solved ? 0 : AK_FLAG_NOT_SOLVED);
    return AK_OK;
}

}  // namespace ak

using namespace ak;

AK_API double ak_forcing_ew(double eta_max, double gamma, double eta, double tol, double n_res, double n_res_prior) {
    // src/Ariadne.jl:207-216
    const double eta_res = gamma * (n_res * n_res) / (n_res_prior * n_res_prior);
    double eta_safe;
    if (gamma * (eta * eta) <= 0.1) eta_safe = fmin(eta_max, eta_res);  // Eq 3.6
    else eta_safe = fmin(eta_max, fmax(eta_res, gamma * (eta * eta)));
    return fmin(eta_max, fmax(eta_safe, 0.5 * tol / n_res));  // Eq 3.5
}

AK_API void ak_newton_default_opts(ak_newton_opts* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->tol_rel = 1.0e-6;
    o->tol_abs = 1.0e-12;
    o->max_niter = 50;
    o->forcing = AK_FORCING_EW;
    o->eta = 0.1;
    o->eta_max = 0.999;
    o->gamma = 0.9;
    o->algo = AK_ALGO_GMRES;
    o->memory = 20;
    o->max_basis = 0;
    ak_krylov_default_opts(&o->krylov);
}

AK_API int ak_newton_solve(ak_ctx* ctx, const ak_problem* p, double* u, double* res, const ak_newton_opts* opts,
                           ak_newton_stats* stats_out, double* hist_nres_host, int64_t* hist_inner_host,
                           double* hist_eta_host, int32_t hist_cap, ak_newton_callback cb, void* cb_user) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && p && u && res && opts && stats_out, "ak_newton_solve: NULL argument");
    Ctx* c = &ctx->c;
    const int64_t n = ak_problem_size(p);
    ak_krylov* ws = nullptr;
    AK_TRY(ak_krylov_create(ctx, opts->algo, n, opts->memory, opts->max_basis, &ws));  // :317-318
    double* rhs = nullptr;
    int rc = ak_malloc(ctx, n, &rhs);
    if (rc == AK_OK)
        rc = newton_ws(c, p, u, res, opts, ws, rhs, stats_out, hist_nres_host, hist_inner_host, hist_eta_host,
                       hist_cap, cb, cb_user);
    cudaStreamSynchronize(c->stream);
    if (rhs) cudaFreeAsync(rhs, c->stream);
    ak_krylov_destroy(ws);
    return rc;
}

AK_API int ak_newton_solve_host(ak_ctx* ctx, const ak_problem* p_in, double* u_host, const double* un_host,
                                const ak_newton_opts* opts, ak_newton_stats* stats_out, double* hist_nres_host,
                                int64_t* hist_inner_host, int32_t hist_cap) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && p_in && u_host && opts && stats_out, "ak_newton_solve_host: NULL argument");
    Ctx* c = &ctx->c;
    ak_problem p = *p_in;
    const int64_t n = ak_problem_size(&p);
    double *u = nullptr, *res = nullptr, *un = nullptr, *coef = nullptr;
    int rc = AK_OK;
    do {
        if ((rc = ak_malloc(ctx, n, &u)) != AK_OK) break;
        if ((rc = ak_malloc(ctx, n, &res)) != AK_OK) break;
        if (un_host) {
            if ((rc = ak_malloc(ctx, n, &un)) != AK_OK) break;
            if (cudaMemcpyAsync(un, un_host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = AK_ERR_CUDA; break; }
            p.un = un;
        }
        if ((p.kind == AK_BRATU1D || p.kind == AK_BRATU2D) && p.coef == nullptr) {
            if ((rc = ak_malloc(ctx, n, &coef)) != AK_OK) break;
            p.coef = coef;
        }
        if (cudaMemcpyAsync(u, u_host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = AK_ERR_CUDA; break; }
        if (cudaMemsetAsync(res, 0, sizeof(double) * (size_t)n, c->stream) != cudaSuccess) { rc = AK_ERR_CUDA; break; }  // make_zero!(res) :261
        rc = ak_newton_solve(ctx, &p, u, res, opts, stats_out, hist_nres_host, hist_inner_host, nullptr, hist_cap,
                             nullptr, nullptr);
        if (rc < 0) break;
        if (cudaMemcpyAsync(u_host, u, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = AK_ERR_CUDA; break; }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = AK_ERR_CUDA; break; }
    } while (0);
    if (rc == AK_ERR_CUDA) set_error("ak_newton_solve_host: CUDA copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(c->stream);
    for (double* q : {u, res, un, coef})
        if (q) cudaFreeAsync(q, c->stream);
    return rc;
}

AK_API int ak_implicit_solve(ak_ctx* ctx, ak_problem* p, double* un_dev, int32_t nsteps, const ak_newton_opts* opts_in,
                             int32_t* per_step_newton_host, int64_t* per_step_inner_host,
                             int32_t* per_step_solved_host) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && p && un_dev && opts_in && nsteps >= 0, "ak_implicit_solve: bad argument");
    Ctx* c = &ctx->c;
    const int64_t n = ak_problem_size(p);
    ak_newton_opts o = *opts_in;
    double *u = nullptr, *res = nullptr, *rhs = nullptr;
    ak_krylov* ws = nullptr;
    const double* saved_un = p->un;
    int rc = AK_OK;
    do {
        if ((rc = ak_malloc(ctx, n, &u)) != AK_OK) break;
        if ((rc = ak_malloc(ctx, n, &res)) != AK_OK) break;
        if ((rc = ak_malloc(ctx, n, &rhs)) != AK_OK) break;
        // u = copy(u_n); res = zero(u_n)   implicit.jl:58-60
        if ((rc = launch_copy(c, n, u, un_dev)) != AK_OK) break;
        if ((rc = launch_fill(c, n, res, 0.0)) != AK_OK) break;
        // the reference re-allocates the Krylov workspace every step (src/Ariadne.jl:317-318 inside
        // newton_krylov!); reusing one workspace gives the same numbers
        if ((rc = ak_krylov_create(ctx, o.algo, n, o.memory, o.max_basis, &ws)) != AK_OK) break;
        p->un = un_dev;
        for (int32_t s = 0; s < nsteps; ++s) {
            ak_newton_stats st;
            rc = newton_ws(c, p, u, res, &o, ws, rhs, &st, nullptr, nullptr, nullptr, 0, nullptr, nullptr);
            if (rc < 0) break;
            if (per_step_newton_host) per_step_newton_host[s] = st.outer_iterations;
            if (per_step_inner_host) per_step_inner_host[s] = st.inner_iterations;
            if (per_step_solved_host) per_step_solved_host[s] = st.solved;
            if ((rc = launch_copy(c, n, un_dev, u)) != AK_OK) break;  // u_n .= u  implicit.jl:75
        }
    } while (0);
    p->un = saved_un;
    cudaStreamSynchronize(c->stream);
    for (double* q : {u, res, rhs})
        if (q) cudaFreeAsync(q, c->stream);
    ak_krylov_destroy(ws);
    return rc < 0 ? rc : AK_OK;
}
