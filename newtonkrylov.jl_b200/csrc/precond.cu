// Native preconditioners for the `M` / `N` hooks of newton_krylov! (src/Ariadne.jl:296-297,324-329).
//
//   AK_PRECOND_JACOBI      y = x ./ diag(J(u))
//   AK_PRECOND_TRIDIAG_LU  y = J(u) \ x for the tridiagonal 1-D Bratu Jacobian — the native stand-in for
//                          `N = (J) -> ilu(collect(J))` with `ldiv = true` (examples/bratu.jl:121-139): LU of a
//                          tridiagonal matrix has no fill-in, so its incomplete factors are the complete ones.
//   AK_PRECOND_USER        caller-supplied apply callback
//
// The tridiagonal solve is a partitioned (substructured) Thomas algorithm, the parallel form of what the
// sequential triangular solves of an ILU object do on the CPU:
//   1. rows are cut into blocks of kBlk rows; the last row of each block is a separator.  One thread per block
//      eliminates its interior rows with two O(1)-state sweeps (forward: last interior unknown, backward: first
//      interior unknown) for three right-hand sides (r, the coupling to the previous separator, the coupling to
//      its own separator),
//   2. the Schur complement on the separators is again tridiagonal and 1/kBlk the size: recurse, down to one
//      thread running the plain Thomas algorithm,
//   3. with the separator values known, every block solves its interior by one Thomas sweep.
// Blocks are walked by one thread each (stride kBlk between neighbouring threads); the walks stay in L1
// (kBlk * 8 B = 2 cache lines per thread and array).  This is a preconditioner apply, not the hot path: at
// n = 2^24 it moves ~80n bytes per solve.
#include <math.h>

#include <vector>

#include "ak_internal.h"
#include "common.cuh"

namespace ak {

namespace {

constexpr int kBlk = 32;        // rows per block (interior rows + 1 separator)
constexpr int kSerial = 2048;   // systems up to this size are solved by one thread

// ---- coefficient providers: row i of the tridiagonal matrix is (lo(i), di(i), up(i)), right-hand side rhs(i) ----
struct BratuRows {  // J = tridiag(o, -2 o + lambda e^{u_i}, o)   (examples/bratu.jl:14-24 linearised)
    const double* coef;  // lambda * exp(u) cached by the residual kernel, or nullptr
    const double* u;
    const double* r;
    double o, lambda;
    int64_t n;
    __device__ double lo(int64_t i) const { return i > 0 ? o : 0.0; }
    __device__ double up(int64_t i) const { return i < n - 1 ? o : 0.0; }
    __device__ double di(int64_t i) const { return -2.0 * o + (coef ? coef[i] : lambda * exp(u[i])); }
    __device__ double rhs(int64_t i) const { return r[i]; }
};
struct ArrayRows {  // explicit arrays (the reduced systems)
    const double *a, *b, *c, *r;
    int64_t n;
    __device__ double lo(int64_t i) const { return a[i]; }
    __device__ double up(int64_t i) const { return c[i]; }
    __device__ double di(int64_t i) const { return b[i]; }
    __device__ double rhs(int64_t i) const { return r[i]; }
};

__host__ __device__ inline int64_t num_blocks(int64_t n) { return (n + kBlk - 1) / kBlk; }

// 1. eliminate the interior rows of every block -> reduced tridiagonal system on the separators (ra, rb, rc, rr).
//    Thread k owns interior rows [k kBlk, sep_k) and separator sep_k = min((k+1) kBlk, n) - 1; it writes its own
//    contribution to reduced row k ("left" part) and to reduced row k-1 ("right" part, through part_*).
template <class Rows>
__global__ void __launch_bounds__(128) k_tri_reduce(const Rows M, int64_t P, double* __restrict__ ra,
                                                    double* __restrict__ rb, double* __restrict__ rc,
                                                    double* __restrict__ rr, double* __restrict__ part_up,
                                                    double* __restrict__ part_d, double* __restrict__ part_r) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const int64_t s = k * kBlk;
    const int64_t sep = (s + kBlk < M.n ? s + kBlk : M.n) - 1;
    const int64_t m = sep - s;  // interior rows
    // left part of separator row `sep` (through the block's LAST interior unknown), right part of the previous
    // separator row (through the block's FIRST interior unknown)
    double L = M.lo(sep), DL = 0.0, RL = 0.0;     // no interior rows: sep couples to the previous separator directly
    double U = 0.0, DR = 0.0, RR = 0.0;
    if (m > 0) {
        // forward sweep: last interior unknown  z_l = y_l - x_prev p_l - x_sep q_l
        // (B y = r, B p = lo(s) e_1, B q = up(sep-1) e_m)
        double piv = M.di(s);
        double cp = M.up(s) / piv;
        double y = M.rhs(s) / piv, pp = M.lo(s) / piv, q = (m == 1 ? M.up(s) : 0.0) / piv;
        if (m == 1) cp = 0.0;
        for (int64_t i = s + 1; i < sep; ++i) {
            const double a = M.lo(i);
            piv = M.di(i) - a * cp;
            const bool last = (i == sep - 1);
            cp = last ? 0.0 : M.up(i) / piv;
            y = (M.rhs(i) - a * y) / piv;
            pp = (0.0 - a * pp) / piv;
            q = ((last ? M.up(i) : 0.0) - a * q) / piv;
        }
        const double as = M.lo(sep);
        L = -as * pp;
        DL = -as * q;
        RL = -as * y;
        // backward sweep (UL elimination): first interior unknown  z_f = y_f - x_prev p_f - x_sep q_f
        piv = M.di(sep - 1);
        double ap = (m == 1 ? 0.0 : M.lo(sep - 1)) / piv;
        y = M.rhs(sep - 1) / piv;
        q = M.up(sep - 1) / piv;
        pp = (m == 1 ? M.lo(s) : 0.0) / piv;
        for (int64_t i = sep - 2; i >= s; --i) {
            const double c = M.up(i);
            piv = M.di(i) - c * ap;
            const bool first = (i == s);
            ap = first ? 0.0 : M.lo(i) / piv;
            y = (M.rhs(i) - c * y) / piv;
            q = (0.0 - c * q) / piv;
            pp = ((first ? M.lo(i) : 0.0) - c * pp) / piv;
        }
        if (k > 0) {  // row sep_{k-1}: up(sep_{k-1}) * x_first
            const double cs = M.up(s - 1);
            U = -cs * q;
            DR = -cs * pp;
            RR = -cs * y;
        }
    } else if (k > 0) {
        U = M.up(s - 1);  // the previous separator couples to this one directly
    }
    ra[k] = (k > 0) ? L : 0.0;
    rb[k] = M.di(sep) + DL;   // + DR of block k+1, added by k_tri_assemble
    rr[k] = M.rhs(sep) + RL;  // + RR of block k+1
    if (k > 0) {
        part_up[k - 1] = U;
        part_d[k - 1] = DR;
        part_r[k - 1] = RR;
    }
    if (k == P - 1) {
        part_up[k] = 0.0;
        part_d[k] = 0.0;
        part_r[k] = 0.0;
    }
}
__global__ void __launch_bounds__(256) k_tri_assemble(int64_t P, double* __restrict__ rb, double* __restrict__ rc,
                                                      double* __restrict__ rr, const double* __restrict__ part_up,
                                                      const double* __restrict__ part_d,
                                                      const double* __restrict__ part_r) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    rc[k] = part_up[k];
    rb[k] += part_d[k];
    rr[k] += part_r[k];
}

// plain Thomas algorithm, one thread (top of the recursion).  `cp` scratch of n doubles.
template <class Rows>
__global__ void k_tri_serial(const Rows M, double* __restrict__ x, double* __restrict__ cp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int64_t n = M.n;
    double piv = M.di(0);
    double c = M.up(0) / piv, g = M.rhs(0) / piv;
    cp[0] = c;
    x[0] = g;
    for (int64_t i = 1; i < n; ++i) {
        const double a = M.lo(i);
        piv = M.di(i) - a * c;
        c = M.up(i) / piv;
        g = (M.rhs(i) - a * g) / piv;
        cp[i] = c;
        x[i] = g;
    }
    for (int64_t i = n - 2; i >= 0; --i) {
        g = x[i] - cp[i] * g;
        x[i] = g;
    }
}

// 3. separators known (xs[k]): every block solves its interior rows by one Thomas sweep; x also receives the
//    separator values.  `cp` scratch of n doubles.
template <class Rows>
__global__ void __launch_bounds__(128) k_tri_interior(const Rows M, int64_t P, const double* __restrict__ xs,
                                                      double* __restrict__ x, double* __restrict__ cp) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const int64_t s = k * kBlk;
    const int64_t sep = (s + kBlk < M.n ? s + kBlk : M.n) - 1;
    const double xsep = xs[k];
    const double xprev = k > 0 ? xs[k - 1] : 0.0;
    x[sep] = xsep;
    if (sep == s) return;
    double piv = M.di(s);
    double c = M.up(s) / piv;
    double g = ((M.rhs(s) - M.lo(s) * xprev) - (sep - 1 == s ? M.up(s) * xsep : 0.0)) / piv;
    cp[s] = c;
    x[s] = g;
    for (int64_t i = s + 1; i < sep; ++i) {
        const double a = M.lo(i);
        piv = M.di(i) - a * c;
        c = M.up(i) / piv;
        double r = M.rhs(i);
        if (i == sep - 1) r -= M.up(i) * xsep;
        g = (r - a * g) / piv;
        cp[i] = c;
        x[i] = g;
    }
    for (int64_t i = sep - 2; i >= s; --i) {
        g = x[i] - cp[i] * g;
        x[i] = g;
    }
}

struct Pool {  // stream-ordered scratch, freed when the solve has been enqueued
    Ctx* c;
    std::vector<void*> ptrs;
    explicit Pool(Ctx* ctx) : c(ctx) {}
    double* get(int64_t n) {
        void* p = nullptr;
        if (pool_alloc(c, (void**)&p, sizeof(double) * (size_t)(n > 0 ? n : 1)) != cudaSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        ptrs.push_back(p);
        return (double*)p;
    }
    ~Pool() { for (void* p : ptrs) cudaFreeAsync(p, c->stream); }
};

template <class Rows>
int tri_solve(Ctx* ctx, Pool& pool, const Rows& M, double* x, double* cp) {
    const int64_t n = M.n;
    if (n <= kSerial) {
        k_tri_serial<Rows><<<1, 32, 0, ctx->stream>>>(M, x, cp);
        ctx->launches++;
        AK_CUDA(cudaGetLastError());
        return AK_OK;
    }
    const int64_t P = num_blocks(n);
    double* buf = pool.get(8 * P);  // ra rb rc rr | part_up part_d part_r | xs
    double* cp2 = pool.get(P);
    if (!buf || !cp2) { set_error("tridiagonal solve: out of device memory"); return AK_ERR_NOMEM; }
    double *ra = buf, *rb = buf + P, *rc = buf + 2 * P, *rr = buf + 3 * P;
    double *pu = buf + 4 * P, *pd = buf + 5 * P, *pr = buf + 6 * P, *xs = buf + 7 * P;
    const int g1 = (int)((P + 127) / 128), g2 = (int)((P + 255) / 256);
    k_tri_reduce<Rows><<<g1, 128, 0, ctx->stream>>>(M, P, ra, rb, rc, rr, pu, pd, pr);
    k_tri_assemble<<<g2, 256, 0, ctx->stream>>>(P, rb, rc, rr, pu, pd, pr);
    ctx->launches += 2;
    AK_CUDA(cudaGetLastError());
    ArrayRows R{ra, rb, rc, rr, P};
    AK_TRY(tri_solve<ArrayRows>(ctx, pool, R, xs, cp2));
    k_tri_interior<Rows><<<g1, 128, 0, ctx->stream>>>(M, P, xs, x, cp);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

// y = x ./ d(u)
__global__ void __launch_bounds__(256) k_jacobi(double* __restrict__ y, const double* __restrict__ x,
                                                const double* __restrict__ coef, const double* __restrict__ u,
                                                double d0, double lambda, double d_edge, int64_t n) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += nth) {
        double d = d0;
        if (coef) d = __dadd_rn(d0, coef[j]);
        else if (u) d = __dadd_rn(d0, __dmul_rn(lambda, exp(u[j])));
        if (d_edge != 0.0 && (j == 0 || j == n - 1)) d = d_edge;
        y[j] = __ddiv_rn(x[j], d);
    }
}

}  // namespace

int precond_apply(Ctx* ctx, const ak_problem* p, const double* u, int32_t kind, ak_precond_apply_fn fn, void* user,
                  const double* x, double* y) {
    const int64_t n = ak_problem_size(p);
    if (kind == AK_PRECOND_USER) {
        AK_REQUIRE(fn != nullptr, "AK_PRECOND_USER without an apply callback");
        const int rc = fn(user, (uint64_t)(uintptr_t)ctx->stream, x, y);
        if (rc != 0) { set_error("user preconditioner callback returned %d", rc); return AK_ERR_USER; }
        return AK_OK;
    }
    if (kind == AK_PRECOND_JACOBI) {
        const double c1 = (p->scheme == AK_EULER) ? p->dt : p->dt / 2.0;
        double d0 = 0.0, d_edge = 0.0;
        bool bratu = false;
        switch (p->kind) {
            case AK_BRATU1D: d0 = -2.0 / (p->dx * p->dx); bratu = true; break;
            case AK_BRATU2D: d0 = -2.0 / (p->dx * p->dx) - 2.0 / (p->dy * p->dy); bratu = true; break;
            case AK_HEAT1D:
                d0 = c1 * (p->a * (-2.0 / (p->dx * p->dx))) - 1.0;
                // zero rows of J at the two boundary points (bc!): any non-zero pivot; multi-GPU: global ends only
                if (p->bc == AK_BC_ZERO) d_edge = -1.0;
                break;
            case AK_HEAT2D: d0 = c1 * (p->a * (-2.0 / (p->dx * p->dx) - 2.0 / (p->dy * p->dy))) - 1.0; break;
            default:
                set_error("AK_PRECOND_JACOBI is not implemented for problem kind %d", p->kind);
                return AK_ERR_UNSUPPORTED;
        }
        if (d_edge != 0.0 && ctx->nranks > 1) {
            set_error("AK_PRECOND_JACOBI for the 1-D heat problem with bc! is single-GPU in this version");
            return AK_ERR_UNSUPPORTED;
        }
        if (bratu) AK_REQUIRE(p->coef != nullptr || u != nullptr, "Jacobi preconditioner needs u");
        int64_t b = (n + 255) / 256, cap = (int64_t)ctx->num_sms * 8;
        k_jacobi<<<(int)(b < cap ? b : cap), 256, 0, ctx->stream>>>(y, x, bratu ? p->coef : nullptr,
                                                                   bratu && !p->coef ? u : nullptr, d0, p->lambda,
                                                                   d_edge, n);
        ctx->launches++;
        AK_CUDA(cudaGetLastError());
        return AK_OK;
    }
    if (kind == AK_PRECOND_TRIDIAG_LU) {
        if (p->kind != AK_BRATU1D) {
            set_error("AK_PRECOND_TRIDIAG_LU is implemented for AK_BRATU1D (the ilu(collect(J)) call sites of examples/bratu.jl)");
            return AK_ERR_UNSUPPORTED;
        }
        if (ctx->nranks > 1) {
            set_error("AK_PRECOND_TRIDIAG_LU is single-GPU in this version");
            return AK_ERR_UNSUPPORTED;
        }
        AK_REQUIRE(p->coef != nullptr || u != nullptr, "tridiagonal LU preconditioner needs u");
        AK_REQUIRE(x != y, "tridiagonal LU preconditioner: in-place apply is not supported");
        Pool pool(ctx);
        double* cp = pool.get(n);
        if (!cp) { set_error("tridiagonal solve: out of device memory"); return AK_ERR_NOMEM; }
        BratuRows M{p->coef, u, x, 1.0 / (p->dx * p->dx), p->lambda, n};
        return tri_solve<BratuRows>(ctx, pool, M, y, cp);
    }
    set_error("unknown preconditioner kind %d", kind);
    return AK_ERR_UNSUPPORTED;
}

}  // namespace ak
