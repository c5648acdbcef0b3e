// Generic residual seam: caller-supplied F!(res, u, p) / tangent callbacks (AK_USER) and the generic
// finite-difference JVP (AK_JVP_FD).
//
//   F!(res, u, p) -> nothing                          src/Ariadne.jl:250-256 (contract), :302, :349 (call sites)
//   mul!(out, J::JacobianOperator, v)                 src/Ariadne.jl:48-57
//
// The reference differentiates an arbitrary user F! with Enzyme.  A native library cannot do that, so the
// seam takes the tangent as a second callback (what a Julia caller gets from `Enzyme.autodiff(Forward, ...)`
// on their CUDA.jl kernel, a Python caller from torch.func.jvp); without one, J v is the finite difference
// BASELINE.json's north_star words: (F(u + eps v) - F(u)) / eps, two residual evaluations (one when
// ak_residual cached F(u) in p->coef), O(sqrt(eps_mach)) accurate.
#include <math.h>

#include "ak_internal.h"
#include "common.cuh"

namespace ak {

namespace {

// eps[0] = step, eps[1] = 1 / step (0 when v == 0: the quotient kernel then writes zeros)
__global__ void k_fd_step(const double* __restrict__ uu, const double* __restrict__ vv, double fd_eps,
                          double* __restrict__ eps) {
    if (threadIdx.x != 0) return;
    const double vn = sqrt(*vv);
    double e = 0.0;
    if (vn > 0.0) e = fd_eps > 0.0 ? fd_eps : 1.4901161193847656e-08 * (1.0 + sqrt(*uu)) / vn;
    eps[0] = e;
    eps[1] = e > 0.0 ? 1.0 / e : 0.0;
}

// t = u + eps v
__global__ void __launch_bounds__(256) k_fd_perturb(double* __restrict__ t, const double* __restrict__ u,
                                                    const double* __restrict__ v, const double* __restrict__ eps,
                                                    int64_t n) {
    const double e = eps[0];
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += nth) t[j] = fma(e, v[j], u[j]);
}

// out = (out - f0) / eps
__global__ void __launch_bounds__(256) k_fd_quotient(double* __restrict__ out, const double* __restrict__ f0,
                                                     const double* __restrict__ eps, int64_t n) {
    const double inv = eps[1];
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += nth)
        out[j] = inv == 0.0 ? 0.0 : (out[j] - f0[j]) * inv;
}

int blocks_for(const Ctx* ctx, int64_t n) {
    int64_t b = (n + 255) / 256, cap = (int64_t)ctx->num_sms * 8;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

struct Scratch {  // stream-ordered scratch vector
    Ctx* c;
    double* p = nullptr;
    Scratch(Ctx* ctx, int64_t n) : c(ctx) {
        if (pool_alloc(c, (void**)&p, sizeof(double) * (size_t)(n > 0 ? n : 1)) != cudaSuccess) {
            p = nullptr;
            (void)cudaGetLastError();
        }
    }
    ~Scratch() { if (p) cudaFreeAsync(p, c->stream); }
};

}  // namespace

int user_residual(Ctx* ctx, const ak_problem* p, double* u, double* res) {
    AK_REQUIRE(p->user_residual != nullptr, "AK_USER problem without a residual callback");
    const int rc = p->user_residual(p->user_data, (uint64_t)(uintptr_t)ctx->stream, u, res);
    if (rc != 0) {
        set_error("user residual callback returned %d", rc);
        return AK_ERR_USER;
    }
    return AK_OK;
}

int user_jvp(Ctx* ctx, const ak_problem* p, const double* u, double* v, double* out) {
    AK_REQUIRE(p->user_jvp != nullptr, "AK_USER problem without a tangent callback");
    const int rc = p->user_jvp(p->user_data, (uint64_t)(uintptr_t)ctx->stream, u, v, out);
    if (rc != 0) {
        set_error("user tangent callback returned %d", rc);
        return AK_ERR_USER;
    }
    return AK_OK;
}

// out <- (F(u + eps v) - F(u)) / eps with F = the problem's own residual (any kind)
int launch_jvp_fd(Ctx* ctx, const ak_problem* p, const double* u, double* v, double* out) {
    AK_REQUIRE(u != nullptr, "AK_JVP_FD needs u");
    const int64_t n = ak_problem_size(p);
    ak_problem q = *p;           // plain residual evaluations: no caches are touched
    q.coef = nullptr;
    q.jvp_mode = AK_JVP_ANALYTIC;
    Scratch t(ctx, n), f0(ctx, p->coef ? 1 : n);
    if (!t.p || !f0.p) { set_error("out of device memory for the finite-difference scratch"); return AK_ERR_NOMEM; }
    double* eps = ctx->dscal + 4;  // {step, 1/step}; dscal[2..3] hold the two sums of squares
    AK_TRY(launch_sumsq(ctx, n, u, ctx->dscal + 2));
    AK_TRY(launch_sumsq(ctx, n, v, ctx->dscal + 3));
    k_fd_step<<<1, 32, 0, ctx->stream>>>(ctx->dscal + 2, ctx->dscal + 3, p->fd_eps, eps);
    const int blocks = blocks_for(ctx, n);
    k_fd_perturb<<<blocks, 256, 0, ctx->stream>>>(t.p, u, v, eps, n);
    ctx->launches += 2;
    AK_CUDA(cudaGetLastError());
    AK_TRY(launch_residual(ctx, &q, t.p, out, nullptr));
    const double* base = p->coef;  // F(u) cached by ak_residual of the same Newton step
    if (!base) {
        // the residual may apply its boundary code to its input in place: evaluate on a copy of u
        AK_TRY(launch_copy(ctx, n, t.p, u));
        AK_TRY(launch_residual(ctx, &q, t.p, f0.p, nullptr));
        base = f0.p;
    }
    k_fd_quotient<<<blocks, 256, 0, ctx->stream>>>(out, base, eps, n);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

}  // namespace ak
