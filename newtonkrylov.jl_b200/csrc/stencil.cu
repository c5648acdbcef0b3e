// Residual F!(res,u,p) and exact tangent-linear JVP kernels (fp64, HBM-bound).
//
// Reference computations replaced (paths relative to the reference tree):
//   bratu!               examples/bratu.jl:14-24            -> OP_RES_BRATU (1-D), and its 2-D extension
//   heat_1D! + bc!       examples/heat_1D.jl:12-37,39-42     -> OP_RES_HEAT (1-D)
//   diffusion! + bc_*!   examples/heat_2D.jl:15-62           -> OP_RES_HEAT (2-D)
//   heat_1D! (DG)        examples/heat_1D_DG.jl:17-36        -> k_dg
//   G_Euler!/G_Midpoint!/G_Trapezoid!  examples/implicit.jl:8-37
//   mul!(out, J, v)      src/Ariadne.jl:48-57 (Enzyme forward mode) -> OP_JVP_* (hand-derived tangents)
//
// Layout: compact nx*ny slab, x fastest, no ghost cells.  Dirichlet ghosts are implicit zeros,
// periodic ghosts are index wraps, inter-GPU ghosts arrive as two row pointers (lo, hi).
//
// Arithmetic follows the Julia source's operation order; __dadd_rn/__dmul_rn/__ddiv_rn are used
// so that nvcc does not contract across the reference's rounding points.
//
// 2-D kernels march down RY rows keeping a 3-row window in registers (each row is read once per
// tile, +2 ghost rows per tile that hit L2), x-neighbours come from warp shuffles.
#include <stdlib.h>

#include "ak_internal.h"
#include "common.cuh"

namespace ak {

enum { OP_RES_BRATU = 0, OP_JVP_BRATU = 1, OP_RES_HEAT = 2, OP_JVP_HEAT = 3, OP_RHS_HEAT = 4, OP_JVP_BRATU_FD = 5 };  // RHS: du = f(u) only; FD: (F(u+eps v)-F(u))/eps
enum { RED_NONE = 0, RED_SUMSQ = 1, RED_DOT = 2, RED_PROJ = 3 };  // PROJ: <proj[b], out> for up to kBlkMax vectors (2-D tangents)

struct StencilArgs {
    int64_t nx, ny;
    int32_t ry;          // rows per tile (2-D)
    int32_t wrap_x;      // periodic in x
    int32_t bc;          // AK_BC_* (1-D boundary-point semantics)
    int32_t scheme;      // AK_EULER / AK_MIDPOINT / AK_TRAPEZOID (heat)
    double dx2, dy2, lambda, a, dt;
    double c0, c1;       // JVP_HEAT: out = c1 * L(c0 * v) - v
    double fd_eps;       // JVP_BRATU_FD
    const double* in;    // u (residual) or v (JVP) or scale_src (fused divcopy)
    const double* lo;    // ghost row y = -1  (nullptr -> 0)
    const double* hi;    // ghost row y = ny  (nullptr -> 0)
    const double* aux;   // RES_HEAT: u_n ; JVP_BRATU: coef (or u when coef_from_u)
    const double* aux_lo;  // ghost rows of u_n (midpoint / trapezoid only)
    const double* aux_hi;
    double* aux_out;     // RES_BRATU: coef out (may be null)
    double* out;
    double* in_write;    // fused divcopy: scaled `in` is stored here (own rows) ; 1-D heat: BC write-back target
    const double* out_scale;  // tangent kernels, un-normalised Krylov basis: out = J(in) / *out_scale (device scalar)
    const double* out_scale_inv;  // optional: 1 / *out_scale already formed (same bits as __ddiv_rn(1, *out_scale))
    Divisor dx2d, dy2d;       // dx2, dy2 with their reciprocals, formed on the host
    const double* bminus;     // tangent kernels: out = bminus - J(in)  (restart residual b - A x of gmres!)
    // RED_PROJ: the first projection pass of the blocked Gram-Schmidt sweep folded into the tangent kernel: raw sums
    // <proj[b], out>, b < nproj, of the vector the kernel has just formed (saves re-reading it: 8n bytes per iteration)
    const double* proj[kBlkMax];
    int32_t nproj;
    double* proj_out;
    P2PDev pd;                    // multi-GPU with peer memory: the sums go to the ranks' mailboxes (seq_out != 0)
    unsigned long long seq_out;
    const double* denom; // fused divcopy: device scalar
    const double* dot_with;
    double* red_out;
    double* partials;
    unsigned int* ticket;
    const int* stop;
    int32_t coef_from_u;
    int32_t seg_first, seg_last;  // 1-D slabs: this rank holds the global first / last point
};

// ---- vector row accessors --------------------------------------------------------------
template <int VEC>
AK_DEV void ldv(const double* p, double (&r)[VEC]);
template <>
AK_DEV void ldv<1>(const double* p, double (&r)[1]) { r[0] = *p; }
template <>
AK_DEV void ldv<2>(const double* p, double (&r)[2]) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    r[0] = t.x; r[1] = t.y;
}
template <>
AK_DEV void ldv<4>(const double* p, double (&r)[4]) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3]) : "l"(p));
}
// streaming variant for operands read exactly once (coef, u_n, dot_with)
template <int VEC>
AK_DEV void ldv_s(const double* p, double (&r)[VEC]);
template <>
AK_DEV void ldv_s<1>(const double* p, double (&r)[1]) { r[0] = ld1_stream(p); }
template <>
AK_DEV void ldv_s<2>(const double* p, double (&r)[2]) {
    const double2 t = ld2_stream(p);
    r[0] = t.x; r[1] = t.y;
}
template <>
AK_DEV void ldv_s<4>(const double* p, double (&r)[4]) {
    const d4 t = ld4_stream(p);
    r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
}
template <int VEC>
AK_DEV void stv(double* p, const double (&r)[VEC]);
template <>
AK_DEV void stv<1>(double* p, const double (&r)[1]) { *p = r[0]; }
template <>
AK_DEV void stv<2>(double* p, const double (&r)[2]) { *reinterpret_cast<double2*>(p) = make_double2(r[0], r[1]); }
template <>
AK_DEV void stv<4>(double* p, const double (&r)[4]) {
    d4 t = {r[0], r[1], r[2], r[3]};
    st4(p, t);
}

// (Divisor / make_divisor / div_by / second_diff: common.cuh)

// ========================================================================================
// 2-D five-point kernels
// ========================================================================================
constexpr int kTX = 128;  // threads per block, all along x

template <int OP, int VEC, bool SCALE, int RED>
__global__ void __launch_bounds__(kTX, OP == OP_JVP_BRATU_FD ? 4 : (RED == RED_PROJ ? 3 : 8)) k_stencil2d(const StencilArgs p) {
    __shared__ double sh[32];
    constexpr bool PROJ = (RED == RED_PROJ);
    double accp[PROJ ? kBlkMax : 1];
#pragma unroll
    for (int b = 0; b < (PROJ ? kBlkMax : 1); ++b) accp[b] = 0.0;
    if (p.stop != nullptr && *p.stop != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t nx = p.nx, ny = p.ny;
    const int64_t x0 = ((int64_t)blockIdx.x * kTX + threadIdx.x) * VEC;
    const bool active = x0 < nx;
    const int64_t y0 = (int64_t)blockIdx.y * p.ry;
    const int64_t y1 = (y0 + p.ry < ny) ? y0 + p.ry : ny;
    Divisor denom;  // only the fused normalisation (fuse = full) divides by a device scalar
    if (SCALE) denom = make_divisor(*p.denom);
    else denom = Divisor{1.0, 1.0, 0};
    const Divisor dx2 = p.dx2d, dy2 = p.dy2d;
    // un-normalised Krylov basis: J is linear, so J(v / rho) is formed as J(v) * (1 / rho) at the store
    const bool oscale_on = (OP == OP_JVP_BRATU || OP == OP_JVP_HEAT || OP == OP_JVP_BRATU_FD) && p.out_scale != nullptr;
    const double oscale = oscale_on ? (p.out_scale_inv != nullptr ? *p.out_scale_inv : __ddiv_rn(1.0, *p.out_scale)) : 1.0;

    auto row_ptr = [&](int64_t y) -> const double* {
        if (y < 0) return p.lo;
        if (y >= ny) return p.hi;
        return p.in + y * nx;
    };
    auto load_row = [&](const double* src, double (&r)[VEC]) {
        if (!active || src == nullptr) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] = 0.0;
            return;
        }
        ldv<VEC>(src + x0, r);
        if (SCALE) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] = div_by(r[i], denom);
        }
    };
    // x-neighbour that is not held by a lane of this warp
    auto edge = [&](const double* src, int64_t x) -> double {
        if (src == nullptr) return 0.0;
        if (x < 0) { if (!p.wrap_x) return 0.0; x += nx; }
        else if (x >= nx) { if (!p.wrap_x) return 0.0; x -= nx; }
        double v = src[x];
        if (SCALE) v = div_by(v, denom);
        return v;
    };

    // second window over u for the fused finite-difference JVP (u + eps v is never materialised)
    constexpr bool FD = (OP == OP_JVP_BRATU_FD);
    auto urow_ptr = [&](int64_t y) -> const double* {
        if (y < 0) return p.aux_lo;
        if (y >= ny) return p.aux_hi;
        return p.aux + y * nx;
    };
    auto uload = [&](const double* src, double (&r)[VEC]) {
        if (!active || src == nullptr) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] = 0.0;
            return;
        }
        ldv<VEC>(src + x0, r);
    };
    auto uedge = [&](const double* src, int64_t x) -> double {
        if (src == nullptr) return 0.0;
        if (x < 0) { if (!p.wrap_x) return 0.0; x += nx; }
        else if (x >= nx) { if (!p.wrap_x) return 0.0; x -= nx; }
        return src[x];
    };
    double uprev[FD ? VEC : 1], ucur[FD ? VEC : 1], unext[FD ? VEC : 1];
    const double* ucur_src = nullptr;
    if (FD) {
        ucur_src = urow_ptr(y0);
        uload(urow_ptr(y0 - 1), reinterpret_cast<double(&)[VEC]>(uprev));
        uload(ucur_src, reinterpret_cast<double(&)[VEC]>(ucur));
    }

    double prev[VEC], cur[VEC], next[VEC];
    const double* cur_src = row_ptr(y0);
    load_row(row_ptr(y0 - 1), prev);
    load_row(cur_src, cur);
    double acc = 0.0;
    // x-neighbours outside the warp are fetched one row ahead so that their latency hides behind a row of work
    const bool need_l = active && lane == 0;
    const bool need_r = active && (lane == 31 || x0 + VEC >= nx);
    double edge_l = need_l ? edge(cur_src, x0 - 1) : 0.0;
    double edge_r = need_r ? edge(cur_src, x0 + VEC) : 0.0;

    for (int64_t y = y0; y < y1; ++y) {
        const double* next_src = row_ptr(y + 1);
        load_row(next_src, next);
        double edge_l_next = 0.0, edge_r_next = 0.0;
        if (y + 1 < y1) {
            if (need_l) edge_l_next = edge(next_src, x0 - 1);
            if (need_r) edge_r_next = edge(next_src, x0 + VEC);
        }
        double left = __shfl_up_sync(0xffffffffu, cur[VEC - 1], 1);
        double right = __shfl_down_sync(0xffffffffu, cur[0], 1);
        const double* unext_src = nullptr;
        double uleft = 0.0, uright = 0.0;
        if (FD) {
            unext_src = urow_ptr(y + 1);
            uload(unext_src, reinterpret_cast<double(&)[VEC]>(unext));
            uleft = __shfl_up_sync(0xffffffffu, ucur[FD ? VEC - 1 : 0], 1);
            uright = __shfl_down_sync(0xffffffffu, ucur[0], 1);
        }
        if (active) {
            if (need_l) left = edge_l;
            if (need_r) right = edge_r;
            if (FD) {
                if (lane == 0) uleft = uedge(ucur_src, x0 - 1);
                if (lane == 31 || x0 + VEC >= nx) uright = uedge(ucur_src, x0 + VEC);
            }
            const int64_t off = y * nx + x0;
            double pv[PROJ ? kBlkMax : 1][VEC];  // the block to project on: loads issued before the row's arithmetic
            if (PROJ) {
#pragma unroll
                for (int b = 0; b < (PROJ ? kBlkMax : 1); ++b) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) pv[b][i] = 0.0;
                    if (b < p.nproj) ldv_s<VEC>(p.proj[b] + off, pv[b]);
                }
            }
            double aux[VEC], o[VEC];
            if (OP == OP_JVP_BRATU || OP == OP_RES_HEAT) ldv_s<VEC>(p.aux + off, aux);
            double cf[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const double w = (i == 0) ? left : cur[i - 1];
                const double e = (i == VEC - 1) ? right : cur[(i + 1) % VEC];
                const double c = cur[i];
                const double xx = second_diff(e, c, w, dx2);
                const double yy = second_diff(next[i], c, prev[i], dy2);
                const double lap = __dadd_rn(xx, yy);
                if (FD) {
                    // J v ~ (F(u + eps v) - F(u)) / eps, both residuals evaluated at this point only
                    const int iu = FD ? i : 0;
                    const double eps = p.fd_eps;
                    const double uc = ucur[iu];
                    const double uw = (i == 0) ? uleft : ucur[FD ? i - 1 : 0];
                    const double ue = (i == VEC - 1) ? uright : ucur[FD ? (i + 1) % VEC : 0];
                    const double un_ = unext[iu], us = uprev[iu];
                    const double f0 = __dadd_rn(__dadd_rn(second_diff(ue, uc, uw, dx2), second_diff(un_, uc, us, dy2)),
                                                __dmul_rn(p.lambda, exp(uc)));
                    const double pc = fma(eps, c, uc), pw = fma(eps, w, uw), pe = fma(eps, e, ue);
                    const double pn = fma(eps, next[i], un_), ps = fma(eps, prev[i], us);
                    const double f1 = __dadd_rn(__dadd_rn(second_diff(pe, pc, pw, dx2), second_diff(pn, pc, ps, dy2)),
                                                __dmul_rn(p.lambda, exp(pc)));
                    o[i] = __ddiv_rn(__dsub_rn(f1, f0), eps);
                } else if (OP == OP_RES_BRATU) {
                    cf[i] = __dmul_rn(p.lambda, exp(c));
                    o[i] = __dadd_rn(lap, cf[i]);
                } else if (OP == OP_JVP_BRATU) {
                    const double k = p.coef_from_u ? __dmul_rn(p.lambda, exp(aux[i])) : aux[i];
                    o[i] = __dadd_rn(lap, __dmul_rn(k, c));
                } else if (OP == OP_RES_HEAT) {  // G_Euler!: (un + dt*du) - u
                    const double du = __dmul_rn(p.a, lap);
                    o[i] = __dsub_rn(__dadd_rn(aux[i], __dmul_rn(p.dt, du)), c);
                } else if (OP == OP_RHS_HEAT) {  // du = a * lap (diffusion! alone: heat_2D.jl:53-60)
                    o[i] = __dmul_rn(p.a, lap);
                } else {  // OP_JVP_HEAT: c1 * dv - v
                    const double dv = __dmul_rn(p.a, lap);
                    o[i] = __dsub_rn(__dmul_rn(p.c1, dv), c);
                }
            }
            if (oscale_on) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = __dmul_rn(o[i], oscale);
            }
            if (p.bminus != nullptr) {
                double bb[VEC];
                ldv_s<VEC>(p.bminus + off, bb);
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = __dsub_rn(bb[i], o[i]);
            }
            stv<VEC>(p.out + off, o);
            if (OP == OP_RES_BRATU && p.aux_out != nullptr) stv<VEC>(p.aux_out + off, cf);
            if (SCALE) stv<VEC>(p.in_write + off, cur);
            if (RED == RED_SUMSQ) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc = fma(o[i], o[i], acc);
            } else if (RED == RED_DOT) {
                double dw[VEC];
                ldv_s<VEC>(p.dot_with + off, dw);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc = fma(dw[i], o[i], acc);
            } else if (PROJ) {
#pragma unroll
                for (int b = 0; b < (PROJ ? kBlkMax : 1); ++b) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) accp[b] = fma(pv[b][i], o[i], accp[b]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) { prev[i] = cur[i]; cur[i] = next[i]; }
        cur_src = next_src;
        edge_l = edge_l_next;
        edge_r = edge_r_next;
        if (FD) {
#pragma unroll
            for (int i = 0; i < (FD ? VEC : 1); ++i) { uprev[i] = ucur[i]; ucur[i] = unext[i]; }
            ucur_src = unext_src;
        }
    }
    if (PROJ) {
        double tot[PROJ ? kBlkMax : 1];
        const int bid = blockIdx.y * gridDim.x + blockIdx.x;
        if (grid_reduce_n<(PROJ ? kBlkMax : 1)>(accp, p.partials, p.ticket, bid, gridDim.x * gridDim.y, sh, tot, false)) {
            if (p.seq_out != 0) {
                mail_post(p.pd, p.seq_out, tot, p.nproj);
            } else {
#pragma unroll
                for (int b = 0; b < (PROJ ? kBlkMax : 1); ++b)
                    if (b < p.nproj) p.proj_out[b] = tot[b];
            }
        }
    } else if (RED != RED_NONE) {
        const double s = block_sum(acc, sh);
        const int bid = blockIdx.y * gridDim.x + blockIdx.x;
        grid_sum_finish(s, p.partials, p.ticket, bid, gridDim.x * gridDim.y, p.red_out, sh);
    }
}

// ========================================================================================
// 1-D three-point kernels (Bratu 1-D, heat 1-D with its two boundary points)
// ========================================================================================
constexpr int kT1 = 256;

// value of the 1-D heat state at index i after bc!/periodic_bc! has been applied
AK_DEV double heat1d_bc_value(const double* src, int64_t i, int64_t n, int bc) {
    if (i == 0) return bc == AK_BC_ZERO ? 0.0 : src[n - 2];
    if (i == n - 1) return bc == AK_BC_ZERO ? 0.0 : src[1];
    return src[i];
}

// One chunk of VEC points per thread, one launch-wide wave of small blocks (44-64 registers, 32 warps per SM).
// Two persistent variants were tried in round 2 and dropped (profiles/r02_ncu_1d_two_chunks.txt, r02_ncu_1d_v2_summary.txt):
// with 2-4 chunks in flight per thread they need ~120 registers (16 warps per SM) and, although they execute fewer
// instructions per point, the kernels are instruction-issue bound, not DRAM bound (~50 SASS instructions per point of
// index arithmetic, boundary predicates and the exact-division fix-up around ~12 fp64 operations): half the warps
// issued at 45-60 % and ran 55-78 % of the copy bandwidth, this version 74-91 %.
template <int OP, int VEC, bool SCALE, int RED>
__global__ void __launch_bounds__(kT1) k_stencil1d(const StencilArgs p) {
    __shared__ double sh[32];
    if (p.stop != nullptr && *p.stop != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t n = p.nx;
    const int64_t x0 = ((int64_t)blockIdx.x * kT1 + threadIdx.x) * VEC;
    const bool active = x0 < n;
    Divisor denom;  // only the fused normalisation divides by a device scalar
    if (SCALE) denom = make_divisor(*p.denom);
    else denom = Divisor{1.0, 1.0, 0};
    const Divisor dx2 = p.dx2d;
    constexpr bool HEAT = (OP == OP_RES_HEAT || OP == OP_JVP_HEAT || OP == OP_RHS_HEAT);
    constexpr bool FD = (OP == OP_JVP_BRATU_FD);
    const bool oscale_on = (OP == OP_JVP_BRATU || OP == OP_JVP_HEAT || FD) && p.out_scale != nullptr;
    const double oscale = oscale_on ? (p.out_scale_inv != nullptr ? *p.out_scale_inv : __ddiv_rn(1.0, *p.out_scale)) : 1.0;
    // the chunks that hold a global end point of the heat problem (bc! / periodic_bc! act there, du = 0 there)
    const bool first_chunk = HEAT && x0 == 0 && p.seg_first;
    const bool last_chunk = HEAT && x0 + VEC == n && p.seg_last;
    // fused finite-difference JVP: second window over u (p.aux, ghosts p.aux_lo / p.aux_hi); u + eps v is never stored
    auto uvalue = [&](int64_t i) -> double {
        if (i < 0) return p.aux_lo ? p.aux_lo[0] : 0.0;
        if (i >= n) return p.aux_hi ? p.aux_hi[0] : 0.0;
        return p.aux[i];
    };

    auto value = [&](int64_t i) -> double {  // scalar access incl. boundary semantics
        double v;
        if (i < 0) {  // left of this segment: neighbour rank's last point, or y_0 = 0 (Bratu: bratu.jl:17)
            if (p.lo == nullptr) return 0.0;
            v = p.lo[0];
        } else if (i >= n) {
            if (p.hi == nullptr) return 0.0;
            v = p.hi[0];
        } else if (HEAT && ((i == 0 && p.seg_first) || (i == n - 1 && p.seg_last))) {
            v = heat1d_bc_value(p.in, i, n, p.bc);  // bc! / periodic_bc! act on the global end points only
        } else {
            v = p.in[i];
        }
        if (SCALE) v = div_by(v, denom);
        return v;
    };

    double cur[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) cur[i] = 0.0;
    if (active) {
        ldv<VEC>(p.in + x0, cur);
        if (SCALE) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) cur[i] = div_by(cur[i], denom);
        }
        if (HEAT) {  // the global end points take their BC value
            if (first_chunk) cur[0] = value(0);
            if (last_chunk) cur[VEC - 1] = value(n - 1);
        }
    }
    double left = __shfl_up_sync(0xffffffffu, cur[VEC - 1], 1);
    double right = __shfl_down_sync(0xffffffffu, cur[0], 1);
    double ucur[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) ucur[i] = 0.0;
    if (FD && active) ldv<VEC>(p.aux + x0, ucur);
    double uleft = 0.0, uright = 0.0;
    if (FD) {
        uleft = __shfl_up_sync(0xffffffffu, ucur[VEC - 1], 1);
        uright = __shfl_down_sync(0xffffffffu, ucur[0], 1);
    }
    double acc = 0.0;
    if (active) {
        if (lane == 0) left = value(x0 - 1);
        if (lane == 31 || x0 + VEC >= n) right = value(x0 + VEC);
        if (FD) {
            if (lane == 0) uleft = uvalue(x0 - 1);
            if (lane == 31 || x0 + VEC >= n) uright = uvalue(x0 + VEC);
        }
        double aux[VEC], o[VEC], cf[VEC];
        if (OP == OP_JVP_BRATU || OP == OP_RES_HEAT) ldv_s<VEC>(p.aux + x0, aux);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const double w = (i == 0) ? left : cur[i - 1];
            const double e = (i == VEC - 1) ? right : cur[(i + 1) % VEC];
            const double c = cur[i];
            if (OP == OP_RES_BRATU) {
                cf[i] = __dmul_rn(p.lambda, exp(c));
                o[i] = __dadd_rn(second_diff(e, c, w, dx2), cf[i]);
            } else if (OP == OP_JVP_BRATU) {
                const double k = p.coef_from_u ? __dmul_rn(p.lambda, exp(aux[i])) : aux[i];
                o[i] = __dadd_rn(second_diff(e, c, w, dx2), __dmul_rn(k, c));
            } else if (FD) {
                // J v ~ (F(u + eps v) - F(u)) / eps, both residuals of bratu! (bratu.jl:14-24) evaluated at this point
                const double eps = p.fd_eps;
                const double uc = ucur[i];
                const double uw = (i == 0) ? uleft : ucur[i > 0 ? i - 1 : 0];
                const double ue = (i == VEC - 1) ? uright : ucur[(i + 1) % VEC];
                const double f0 = __dadd_rn(second_diff(ue, uc, uw, dx2), __dmul_rn(p.lambda, exp(uc)));
                const double pc = fma(eps, c, uc), pw = fma(eps, w, uw), pe = fma(eps, e, ue);
                const double f1 = __dadd_rn(second_diff(pe, pc, pw, dx2), __dmul_rn(p.lambda, exp(pc)));
                o[i] = __ddiv_rn(__dsub_rn(f1, f0), eps);
            } else {
                // heat_1D.jl:22: du[i] = a * (u[i+1] - 2u[i] + u[i-1]) / dx^2 ; du[1] = du[end] = 0
                const bool bnd = (i == 0 && first_chunk) || (i == VEC - 1 && last_chunk);
                const double du =
                    bnd ? 0.0 : div_by(__dmul_rn(p.a, __dadd_rn(__dsub_rn(e, __dmul_rn(2.0, c)), w)), dx2);
                if (OP == OP_RES_HEAT) o[i] = __dsub_rn(__dadd_rn(aux[i], __dmul_rn(p.dt, du)), c);
                else if (OP == OP_RHS_HEAT) o[i] = du;
                else o[i] = __dsub_rn(__dmul_rn(p.c1, du), c);
            }
        }
        if (oscale_on) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = __dmul_rn(o[i], oscale);
        }
        stv<VEC>(p.out + x0, o);
        if (OP == OP_RES_BRATU && p.aux_out != nullptr) stv<VEC>(p.aux_out + x0, cf);
        if (SCALE) {
            stv<VEC>(p.in_write + x0, cur);
        } else if (HEAT && p.in_write != nullptr) {
            // the reference's bc!(u) mutates the state / tangent seed in place (heat_1D.jl:16,34-42)
            if (x0 == 0 && p.seg_first) p.in_write[0] = cur[0];
            if (x0 + VEC == n && p.seg_last) p.in_write[n - 1] = cur[VEC - 1];
        }
        if (RED == RED_SUMSQ) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc = fma(o[i], o[i], acc);
        } else if (RED == RED_DOT) {
            double dw[VEC];
            ldv_s<VEC>(p.dot_with + x0, dw);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc = fma(dw[i], o[i], acc);
        }
    }
    if (RED != RED_NONE) {
        const double s = block_sum(acc, sh);
        grid_sum_finish(s, p.partials, p.ticket, blockIdx.x, gridDim.x, p.red_out, sh);
    }
}

// ========================================================================================
// DG (SBP, 4 LGL nodes per element, periodic): du = D1m * (D1p * u)   heat_1D_DG.jl:32-36
// one thread per element; neighbour coupling through warp shuffles
// ========================================================================================
struct DgArgs {
    int64_t ne;        // elements
    double D[4][4];    // LGL derivative matrix
    double jac;        // 2/h
    double mw;         // (h/2) * w_edge,  w_edge = 1/6
    double dt, c0, c1;
    int rhs_only;      // out = D1m*(D1p*in) without the time-discretisation wrapper
    const double* lo;  // slabs of elements: the 4 nodes of the left neighbour rank's last element (nullptr: wrap locally)
    const double* hi;  // first node of the right neighbour rank's first element (nullptr: wrap locally)
    const double* in;
    const double* un;  // residual only
    double* out;
    double* in_write;
    const double* denom;
    const double* out_scale;  // tangent with an un-normalised Krylov basis: out = J(in) / *out_scale
    const double* out_scale_inv;  // optional: 1 / *out_scale already formed
    Divisor mwd;              // mw with its reciprocal, formed on the host
    const double* dot_with;
    double* red_out;
    double* partials;
    unsigned int* ticket;
    const int* stop;
};

AK_DEV void dg_local(const double (&D)[4][4], double jac, const double (&u)[4], double (&o)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double s = __dmul_rn(D[j][0], u[0]);
        s = __dadd_rn(s, __dmul_rn(D[j][1], u[1]));
        s = __dadd_rn(s, __dmul_rn(D[j][2], u[2]));
        s = __dadd_rn(s, __dmul_rn(D[j][3], u[3]));
        o[j] = __dmul_rn(jac, s);
    }
}

// MB: resident blocks per SM the register allocation is sized for.  One element (32 bytes) per thread: the bytes in flight
// are set by the resident warps, so the register count decides the bandwidth (see launch_dg).
// (Round 2 also tried: the two neighbour loads issued up front - 54 registers, 4 blocks, 71 %; the neighbour values
// through shared memory - two block barriers, 69 %.)
template <bool RESIDUAL, bool SCALE, int RED, int MB>
__global__ void __launch_bounds__(kT1, MB) k_dg(const DgArgs p) {
    __shared__ double sh[32];
    if (p.stop != nullptr && *p.stop != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t ne = p.ne;
    const int64_t e = (int64_t)blockIdx.x * kT1 + threadIdx.x;
    const bool active = e < ne;
    Divisor denom;
    if (SCALE) denom = make_divisor(*p.denom);
    else denom = Divisor{1.0, 1.0, 0};
    const Divisor mw = p.mwd;
    // scale of the un-normalised Krylov basis: read up front, not behind the arithmetic
    double oscale = 1.0;
    if (!RESIDUAL && p.out_scale != nullptr)
        oscale = p.out_scale_inv != nullptr ? *p.out_scale_inv : __ddiv_rn(1.0, *p.out_scale);

    auto load_elem = [&](int64_t el, double (&r)[4]) {
        ldv<4>((el < 0) ? p.lo : p.in + 4 * el, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (SCALE) r[i] = div_by(r[i], denom);
            if (!RESIDUAL) r[i] = __dmul_rn(p.c0, r[i]);
        }
    };
    double u[4] = {0, 0, 0, 0}, raw[4] = {0, 0, 0, 0};
    if (active) {
        ldv<4>(p.in + 4 * e, raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (SCALE) raw[i] = div_by(raw[i], denom);
            u[i] = RESIDUAL ? raw[i] : __dmul_rn(p.c0, raw[i]);
        }
    }
    // D1p: needs first node of the element to the right
    double u_next0 = __shfl_down_sync(0xffffffffu, u[0], 1);
    if (active && (lane == 31 || e + 1 >= ne)) {
        const int64_t en = (e + 1 == ne) ? 0 : e + 1;
        double t = (e + 1 == ne && p.hi != nullptr) ? p.hi[0] : p.in[4 * en];
        if (SCALE) t = div_by(t, denom);
        u_next0 = RESIDUAL ? t : __dmul_rn(p.c0, t);
    }
    double t1[4] = {0, 0, 0, 0};
    if (active) {
        dg_local(p.D, p.jac, u, t1);
        t1[3] = __dadd_rn(t1[3], div_by(__dsub_rn(u_next0, u[3]), mw));
    }
    // D1m: needs last node of (D1p u) of the element to the left
    double t_prev3 = __shfl_up_sync(0xffffffffu, t1[3], 1);
    if (active && lane == 0) {
        const int64_t ep = (e == 0) ? ((p.lo != nullptr) ? -1 : ne - 1) : e - 1;
        double up[4], tp[4];
        load_elem(ep, up);
        dg_local(p.D, p.jac, up, tp);
        t_prev3 = __dadd_rn(tp[3], div_by(__dsub_rn(u[0], up[3]), mw));
    }
    double acc = 0.0;
    if (active) {
        double du[4], o[4];
        dg_local(p.D, p.jac, t1, du);
        du[0] = __dadd_rn(du[0], div_by(__dsub_rn(t1[0], t_prev3), mw));
        if (p.rhs_only) {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = du[i];
        } else if (RESIDUAL) {
            double un[4];
            ldv_s<4>(p.un + 4 * e, un);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = __dsub_rn(__dadd_rn(un[i], __dmul_rn(p.dt, du[i])), raw[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = __dsub_rn(__dmul_rn(p.c1, du[i]), raw[i]);
            if (p.out_scale != nullptr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = __dmul_rn(o[i], oscale);
            }
        }
        stv<4>(p.out + 4 * e, o);
        if (SCALE) stv<4>(p.in_write + 4 * e, raw);
        if (RED == RED_SUMSQ) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc = fma(o[i], o[i], acc);
        } else if (RED == RED_DOT) {
            double dw[4];
            ldv_s<4>(p.dot_with + 4 * e, dw);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc = fma(dw[i], o[i], acc);
        }
    }
    if (RED != RED_NONE) {
        const double s = block_sum(acc, sh);
        grid_sum_finish(s, p.partials, p.ticket, blockIdx.x, gridDim.x, p.red_out, sh);
    }
}

// ========================================================================================
// 2x2 system of test/runtests.jl:4-7 (known-answer path for the host logic)
// ========================================================================================
__global__ void k_simple2(const double* u, const double* v, double* out, double* red_out, int mode,
                          const double* dot_with, const int* stop) {
    if (stop != nullptr && *stop != 0) return;
    if (threadIdx.x != 0) return;
    const double x = u[0], y = u[1];
    double o0, o1;
    if (mode == 0) {  // residual
        o0 = __dsub_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), 2.0);
        o1 = __dsub_rn(__dadd_rn(exp(__dsub_rn(x, 1.0)), __dmul_rn(y, y)), 2.0);
    } else if (mode == 1) {  // J v
        o0 = __dadd_rn(__dmul_rn(__dmul_rn(2.0, x), v[0]), __dmul_rn(__dmul_rn(2.0, y), v[1]));
        o1 = __dadd_rn(__dmul_rn(exp(__dsub_rn(x, 1.0)), v[0]), __dmul_rn(__dmul_rn(2.0, y), v[1]));
    } else {  // J^T v
        o0 = __dadd_rn(__dmul_rn(__dmul_rn(2.0, x), v[0]), __dmul_rn(exp(__dsub_rn(x, 1.0)), v[1]));
        o1 = __dadd_rn(__dmul_rn(__dmul_rn(2.0, y), v[0]), __dmul_rn(__dmul_rn(2.0, y), v[1]));
    }
    out[0] = o0;
    out[1] = o1;
    if (red_out != nullptr) {
        if (dot_with != nullptr) *red_out = fma(dot_with[1], o1, dot_with[0] * o0);
        else *red_out = fma(o1, o1, o0 * o0);
    }
}

// ========================================================================================
// Multi-RHS tangent: Out[:, c] = J(u) V[:, c], c < ncols  —  `mul!(Out, J, V)` of src/Ariadne.jl:69-83 and the probe
// products of `collect(J)` (:140-162).  The operand every column shares, lambda e^u, is read (or computed: one `exp`
// per point) once per group of kCB = 4 columns: (16 + 8/4) n = 18n bytes per column instead of 24n, one `exp` per
// point and group instead of one per column, and one launch instead of ncols.  Arithmetic per column is the single-column kernel's, bit for bit.
// (The heat and DG tangents do not depend on u: their columns share nothing and stay a loop of single launches.)
// ========================================================================================
struct BatchArgs {
    int64_t nx, ny;
    double dx2, dy2, lambda;
    const double* aux;   // cached lambda e^u, or u when coef_from_u
    int32_t coef_from_u;
    const double* V;
    int64_t ldv;
    double* Out;
    int64_t ldo;
    int32_t ncols;
};
constexpr int kBatchRY = 16;  // rows per tile of the 2-D multi-RHS kernel
constexpr int kCB = 4;        // columns that share one read of lambda e^u (and are in flight together)

// grid = (x tiles, row tiles, column groups of kCB).  Per row step every thread issues the loads of the next row of all
// kCB columns plus the lambda e^u row before it consumes anything (kCB + 1 requests in flight, 3-row register windows of
// kCB columns): (16 + 8/kCB) n bytes per column.
template <int VEC>
__global__ void __launch_bounds__(kTX, 3) k_bratu2d_jvp_batched(const BatchArgs p) {
    const int lane = threadIdx.x & 31;
    const int64_t nx = p.nx, ny = p.ny;
    const int64_t x0 = ((int64_t)blockIdx.x * kTX + threadIdx.x) * VEC;
    const bool active = x0 < nx;
    const int64_t y0 = (int64_t)blockIdx.y * kBatchRY;
    const int64_t y1 = (y0 + kBatchRY < ny) ? y0 + kBatchRY : ny;
    const int c0 = blockIdx.z * kCB;
    const Divisor dx2 = make_divisor(p.dx2), dy2 = make_divisor(p.dy2);
    const bool need_l = active && lane == 0;
    const bool need_r = active && (lane == 31 || x0 + VEC >= nx);
    auto col_in = [&](int c) -> const double* { return p.V + (int64_t)(c0 + c) * p.ldv; };
    auto load_row = [&](int c, int64_t y, double (&r)[VEC]) {
        if (!active || y < 0 || y >= ny || c0 + c >= p.ncols) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] = 0.0;
            return;
        }
        ldv<VEC>(col_in(c) + y * nx + x0, r);
    };
    auto edge = [&](int c, int64_t y, int64_t x) -> double {
        return (x < 0 || x >= nx || c0 + c >= p.ncols) ? 0.0 : col_in(c)[y * nx + x];
    };
    double prev[kCB][VEC], cur[kCB][VEC], next[kCB][VEC], el[kCB], er[kCB];
#pragma unroll
    for (int c = 0; c < kCB; ++c) {
        load_row(c, y0 - 1, prev[c]);
        load_row(c, y0, cur[c]);
        el[c] = need_l ? edge(c, y0, x0 - 1) : 0.0;
        er[c] = need_r ? edge(c, y0, x0 + VEC) : 0.0;
    }
    for (int64_t y = y0; y < y1; ++y) {
        double kc[VEC], eln[kCB], ern[kCB];
#pragma unroll
        for (int i = 0; i < VEC; ++i) kc[i] = 0.0;
        // every load of this row step first
#pragma unroll
        for (int c = 0; c < kCB; ++c) {
            load_row(c, y + 1, next[c]);
            eln[c] = (need_l && y + 1 < y1) ? edge(c, y + 1, x0 - 1) : 0.0;
            ern[c] = (need_r && y + 1 < y1) ? edge(c, y + 1, x0 + VEC) : 0.0;
        }
        if (active) ldv_s<VEC>(p.aux + y * nx + x0, kc);
        if (p.coef_from_u) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) kc[i] = __dmul_rn(p.lambda, exp(kc[i]));
        }
#pragma unroll
        for (int c = 0; c < kCB; ++c) {
            double left = __shfl_up_sync(0xffffffffu, cur[c][VEC - 1], 1);
            double right = __shfl_down_sync(0xffffffffu, cur[c][0], 1);
            if (active && c0 + c < p.ncols) {
                if (need_l) left = el[c];
                if (need_r) right = er[c];
                double o[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const double w = (i == 0) ? left : cur[c][i > 0 ? i - 1 : 0];
                    const double e = (i == VEC - 1) ? right : cur[c][(i + 1) % VEC];
                    const double cc = cur[c][i];
                    const double lap = __dadd_rn(second_diff(e, cc, w, dx2), second_diff(next[c][i], cc, prev[c][i], dy2));
                    o[i] = __dadd_rn(lap, __dmul_rn(kc[i], cc));
                }
                stv<VEC>(p.Out + (int64_t)(c0 + c) * p.ldo + y * nx + x0, o);
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) { prev[c][i] = cur[c][i]; cur[c][i] = next[c][i]; }
            el[c] = eln[c];
            er[c] = ern[c];
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(kT1, 2) k_bratu1d_jvp_batched(const BatchArgs p) {
    const int lane = threadIdx.x & 31;
    const int64_t n = p.nx;
    const Divisor dx2 = make_divisor(p.dx2);
    const int64_t nchunks = (n + VEC - 1) / VEC;
    const int64_t nth = (int64_t)gridDim.x * kT1;
    for (int64_t ch = (int64_t)blockIdx.x * kT1 + threadIdx.x; ch - lane < nchunks; ch += nth) {
        const int64_t x0 = ch * VEC;
        const bool active = x0 < n;
        const bool last = lane == 31 || x0 + VEC >= n;
        double kc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) kc[i] = 0.0;
        if (active) ldv_s<VEC>(p.aux + x0, kc);
        for (int cg = 0; cg < p.ncols; cg += kCB) {  // kCB columns in flight together
            double cur[kCB][VEC], el[kCB], er[kCB];
#pragma unroll
            for (int c = 0; c < kCB; ++c) {
                const double* in = p.V + (int64_t)(cg + c) * p.ldv;
                const bool on = active && cg + c < p.ncols;
#pragma unroll
                for (int i = 0; i < VEC; ++i) cur[c][i] = 0.0;
                el[c] = er[c] = 0.0;
                if (on) {
                    ldv<VEC>(in + x0, cur[c]);
                    if (lane == 0 && x0 > 0) el[c] = in[x0 - 1];
                    if (last && x0 + VEC < n) er[c] = in[x0 + VEC];
                }
            }
            if (cg == 0 && p.coef_from_u) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) kc[i] = __dmul_rn(p.lambda, exp(kc[i]));
            }
#pragma unroll
            for (int c = 0; c < kCB; ++c) {
                double left = __shfl_up_sync(0xffffffffu, cur[c][VEC - 1], 1);
                double right = __shfl_down_sync(0xffffffffu, cur[c][0], 1);
                if (active && cg + c < p.ncols) {
                    if (lane == 0) left = el[c];
                    if (last) right = er[c];
                    double o[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const double w = (i == 0) ? left : cur[c][i > 0 ? i - 1 : 0];
                        const double e = (i == VEC - 1) ? right : cur[c][(i + 1) % VEC];
                        o[i] = __dadd_rn(second_diff(e, cur[c][i], w, dx2), __dmul_rn(kc[i], cur[c][i]));
                    }
                    stv<VEC>(p.Out + (int64_t)(cg + c) * p.ldo + x0, o);
                }
            }
        }
    }
}

// out <- J^T v for the 1-D heat operator with periodic_bc! (heat_1D.jl:39-42).  The forward tangent is
// J = (c1 a L - I) B: B copies v[n-2] -> v[0], v[1] -> v[n-1] (the BC code), L is the three-point row
// a (e - 2c + w) / dx^2 on the interior rows and zero on rows 0 and n-1.  Hence J^T = B^T (c1 a L^T - I):
//   y[j]   = c1 * (a * ((wt[j+1] - 2 wt[j]) + wt[j-1]) / dx^2) - v[j],  wt = v with wt[0] = wt[n-1] = 0
//   out[j] = y[j] (interior), out[1] += y[n-1], out[n-2] += y[0], out[0] = out[n-1] = 0.
// Applied to a unit vector this reproduces the entries of the forward kernel bit for bit (same formula per entry).
__global__ void __launch_bounds__(256) k_heat1d_periodic_transpose(const double* __restrict__ v, double* __restrict__ out,
                                                                   int64_t n, double a, double dx2v, double c1) {
    const Divisor dx2 = make_divisor(dx2v);
    auto wt = [&](int64_t i) -> double { return (i <= 0 || i >= n - 1) ? 0.0 : v[i]; };
    auto y = [&](int64_t j) -> double {
        const double du = div_by(__dmul_rn(a, __dadd_rn(__dsub_rn(wt(j + 1), __dmul_rn(2.0, wt(j))), wt(j - 1))), dx2);
        return __dsub_rn(__dmul_rn(c1, du), v[j]);
    };
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += nth) {
        double o;
        if (j == 0 || j == n - 1) o = 0.0;
        else {
            o = y(j);
            if (j == 1) o = __dadd_rn(o, y(n - 1));
            if (j == n - 2) o = __dadd_rn(o, y(0));
        }
        out[j] = o;
    }
}

// ========================================================================================
// host-side dispatch
// ========================================================================================
static inline bool al(const void* p, int bytes) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % bytes) == 0; }

template <int OP, int VEC>
static void launch2d_v(Ctx* ctx, const StencilArgs& a, bool scale, int red, dim3 grid) {
#define AK_L2D(S, R) k_stencil2d<OP, VEC, S, R><<<grid, kTX, 0, ctx->stream>>>(a)
    if (scale) {
        if (red == RED_DOT) AK_L2D(true, RED_DOT);
        else AK_L2D(true, RED_NONE);
    } else {
        if (red == RED_DOT) AK_L2D(false, RED_DOT);
        else if (red == RED_SUMSQ) AK_L2D(false, RED_SUMSQ);
        else if (red == RED_PROJ) {
            if constexpr (OP == OP_JVP_BRATU || OP == OP_JVP_HEAT) AK_L2D(false, RED_PROJ);
        } else AK_L2D(false, RED_NONE);
    }
#undef AK_L2D
}
template <int OP>
static int launch2d(Ctx* ctx, StencilArgs& a, bool scale, int red) {
    // widest vector the row pitch and every operand allow
    int vec = 1;
    const void* ptrs[] = {a.in, a.lo, a.hi, a.aux, a.aux_out, a.out, a.in_write, a.dot_with, a.bminus};
    auto ok = [&](int v) {
        if (a.nx % v) return false;
        for (const void* q : ptrs)
            if (!al(q, 8 * v)) return false;
        for (int b = 0; b < a.nproj; ++b)
            if (!al(a.proj[b], 8 * v)) return false;
        return true;
    };
    if (ok(4)) vec = 4;
    else if (ok(2)) vec = 2;
    // rows per tile: enough tiles to fill the machine several times over, few ghost re-reads
    int64_t gx = (a.nx + (int64_t)kTX * vec - 1) / ((int64_t)kTX * vec);
    int ry = 16;
    while (ry > 2 && gx * ((a.ny + ry - 1) / ry) < (int64_t)ctx->num_sms * 16 * 4) ry >>= 1;
    if (const char* e = getenv("AK_RY")) {  // tuning knob (rows per tile)
        const int v = atoi(e);
        if (v >= 1) ry = v;
    }
    a.ry = ry;
    int64_t gy = (a.ny + ry - 1) / ry;
    const int64_t max_blocks = red == RED_PROJ ? kMaxPartials / kBlkMax : kMaxPartials;  // partials per block: 8 or 1
    if (gx * gy > max_blocks && red != RED_NONE) {  // keep the partials buffer in bounds
        ry = (int)((a.ny * gx + max_blocks - 1) / max_blocks);
        a.ry = ry;
        gy = (a.ny + ry - 1) / ry;
    }
    if (gy > 65535) { a.ry = (int32_t)((a.ny + 65534) / 65535); gy = (a.ny + a.ry - 1) / a.ry; }
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (vec == 4) launch2d_v<OP, 4>(ctx, a, scale, red, grid);
    else if (vec == 2) launch2d_v<OP, 2>(ctx, a, scale, red, grid);
    else launch2d_v<OP, 1>(ctx, a, scale, red, grid);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

template <int OP, int VEC>
static void launch1d_v(Ctx* ctx, const StencilArgs& a, bool scale, int red, int grid) {
#define AK_L1D(S, R) k_stencil1d<OP, VEC, S, R><<<grid, kT1, 0, ctx->stream>>>(a)
    if (scale) {
        if (red == RED_DOT) AK_L1D(true, RED_DOT);
        else AK_L1D(true, RED_NONE);
    } else {
        if (red == RED_DOT) AK_L1D(false, RED_DOT);
        else if (red == RED_SUMSQ) AK_L1D(false, RED_SUMSQ);
        else AK_L1D(false, RED_NONE);
    }
#undef AK_L1D
}
template <int OP>
static int launch1d(Ctx* ctx, StencilArgs& a, bool scale, int red) {
    int vec = 1;
    const void* ptrs[] = {a.in, a.aux, a.aux_out, a.out, a.in_write, a.dot_with};
    auto ok = [&](int v) {
        if (a.nx % v) return false;
        for (const void* q : ptrs)
            if (!al(q, 8 * v)) return false;
        return true;
    };
    if (ok(4)) vec = 4;
    else if (ok(2)) vec = 2;
    int64_t grid = (a.nx + (int64_t)kT1 * vec - 1) / ((int64_t)kT1 * vec);
    if (grid > kMaxPartials && red != RED_NONE) {
        set_error("1-D stencil with fused reduction: n too large for the partials buffer");
        return AK_ERR_UNSUPPORTED;
    }
    if (vec == 4) launch1d_v<OP, 4>(ctx, a, scale, red, (int)grid);
    else if (vec == 2) launch1d_v<OP, 2>(ctx, a, scale, red, (int)grid);
    else launch1d_v<OP, 1>(ctx, a, scale, red, (int)grid);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

static void dg_constants(DgArgs& d, double h) {
    const double s5 = sqrt(5.0);
    const double a = (5.0 + 5.0 * s5) / 4.0, b = (5.0 - 5.0 * s5) / 4.0;
    const double c = (1.0 + s5) / 4.0, dd = s5 / 2.0, e = (s5 - 1.0) / 4.0;
    const double M[4][4] = {{-3.0, a, b, 0.5}, {-c, 0.0, dd, -e}, {e, -dd, 0.0, c}, {-0.5, -b, -a, 3.0}};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) d.D[i][j] = M[i][j];
    d.jac = 2.0 / h;
    d.mw = (h / 2.0) * (1.0 / 6.0);
    d.mwd = make_divisor_host(d.mw);
}

static int launch_dg(Ctx* ctx, DgArgs& d, bool residual, bool scale, int red) {
    const int64_t grid = (d.ne + kT1 - 1) / kT1;
    if (grid > kMaxPartials && red != RED_NONE) {
        set_error("DG stencil with fused reduction: n too large for the partials buffer");
        return AK_ERR_UNSUPPORTED;
    }
    // resident blocks per SM (register budget) of the un-scaled kernels: 8 blocks = 32 registers = all 64 warps of an
    // SM in flight (measured at 2^22 elements: tangent 59.6 us at 46 registers, 49.3 us at 40, 45.4 us at 32 = 90 % of the
    // copy bandwidth; residual 69.8 / 64.1 / 61.4 us = 100 %).  AK_DG_MB = 1 / 6 / 8: tuning knob
    static const int mb = [] { const char* e = getenv("AK_DG_MB"); return e ? atoi(e) : 8; }();
#define AK_LDG1(RS, S, R, M) k_dg<RS, S, R, M><<<(int)grid, kT1, 0, ctx->stream>>>(d)
#define AK_LDG(RS, S, R)                                        \
    do {                                                        \
        if (!(S) && mb == 6) AK_LDG1(RS, false, R, 6);          \
        else if (!(S) && mb == 8) AK_LDG1(RS, false, R, 8);     \
        else AK_LDG1(RS, S, R, 1);                              \
    } while (0)
    if (residual) {
        if (red == RED_SUMSQ) AK_LDG(true, false, RED_SUMSQ);
        else AK_LDG(true, false, RED_NONE);
    } else if (scale) {
        if (red == RED_DOT) AK_LDG(false, true, RED_DOT);
        else AK_LDG(false, true, RED_NONE);
    } else {
        if (red == RED_DOT) AK_LDG(false, false, RED_DOT);
        else if (red == RED_SUMSQ) AK_LDG(false, false, RED_SUMSQ);
        else AK_LDG(false, false, RED_NONE);
    }
#undef AK_LDG1
#undef AK_LDG
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

static int check_problem(const ak_problem* p) {
    AK_REQUIRE(p != nullptr, "problem is NULL");
    AK_REQUIRE(p->kind >= AK_SIMPLE2 && p->kind <= AK_USER, "unknown problem kind");
    if (p->kind == AK_USER) {
        AK_REQUIRE(p->nx >= 1, "AK_USER: nx (number of unknowns) must be >= 1");
        AK_REQUIRE(p->user_residual != nullptr, "AK_USER: user_residual is NULL");
        AK_REQUIRE(p->jvp_mode == AK_JVP_ANALYTIC || p->jvp_mode == AK_JVP_FD, "AK_USER: jvp_mode must be AK_JVP_ANALYTIC or AK_JVP_FD");
        return AK_OK;
    }
    if (p->kind == AK_SIMPLE2) return AK_OK;
    AK_REQUIRE(p->nx >= 1, "nx must be >= 1");
    const bool is2d = (p->kind == AK_BRATU2D || p->kind == AK_HEAT2D);
    if (is2d) AK_REQUIRE(p->ny >= 1, "ny must be >= 1");
    if (p->kind == AK_HEAT1D) AK_REQUIRE(p->nx >= 3, "heat 1-D needs >= 3 points (2 boundary points)");
    if (p->kind == AK_HEAT1D_DG) AK_REQUIRE(p->nx % 4 == 0 && p->nx >= 8, "DG: nx must be 4 * elements, >= 2 elements per rank");
    const bool timedep = (p->kind == AK_HEAT1D || p->kind == AK_HEAT2D || p->kind == AK_HEAT1D_DG);
    if (timedep) {
        AK_REQUIRE(p->scheme == AK_EULER || p->scheme == AK_MIDPOINT || p->scheme == AK_TRAPEZOID,
                   "time-dependent problems need scheme AK_EULER, AK_MIDPOINT or AK_TRAPEZOID");
    } else {
        AK_REQUIRE(p->scheme == AK_STEADY, "Bratu problems are steady (scheme must be AK_STEADY)");
    }
    if (p->jvp_mode != AK_JVP_ANALYTIC && p->jvp_mode != AK_JVP_FD) {
        if (!(p->jvp_mode == AK_JVP_FD_FUSED && (p->kind == AK_BRATU2D || p->kind == AK_BRATU1D))) {
            set_error("jvp_mode %d: AK_JVP_FD_FUSED is implemented for the Bratu problems (heat/DG are linear: use AK_JVP_FD)", p->jvp_mode);
            return AK_ERR_UNSUPPORTED;
        }
    }
    return AK_OK;
}

static int check_multi_gpu(const Ctx* ctx, const ak_problem* p) {
    if (ctx->nranks > 1 && p->kind == AK_HEAT1D && p->bc == AK_BC_PERIODIC) {
        set_error("periodic_bc! of the 1-D heat example couples the two global end points: single-GPU only");
        return AK_ERR_UNSUPPORTED;
    }
    return AK_OK;
}

static void base_args(Ctx* ctx, const ak_problem* p, StencilArgs& a) {
    a = StencilArgs{};
    a.nx = p->nx;
    a.ny = p->ny;
    a.wrap_x = (p->bc == AK_BC_PERIODIC);
    a.bc = p->bc;
    a.scheme = p->scheme;
    a.dx2 = p->dx * p->dx;
    a.dy2 = p->dy * p->dy;
    a.dx2d = make_divisor_host(a.dx2);
    a.dy2d = make_divisor_host(a.dy2);
    a.lambda = p->lambda;
    a.a = p->a;
    a.dt = p->dt;
    a.c0 = 1.0;
    a.c1 = (p->scheme == AK_TRAPEZOID) ? p->dt / 2.0 : p->dt;  // tangent of G_Trapezoid!: (dt/2) f'(v) - v
    a.partials = ctx->partials;
    a.ticket = ctx->ticket;
    a.seg_first = (ctx->rank == 0);
    a.seg_last = (ctx->rank == ctx->nranks - 1);
}

// ghost values of a 1-D segment (multi-GPU only): `nlo` values from the left rank's end, `nhi` from the right rank's start
static int ghost_1d(Ctx* ctx, const double* v, int64_t n, int nlo, int nhi, bool periodic, const double** lo,
                    const double** hi) {
    *lo = nullptr;
    *hi = nullptr;
    if (ctx->nranks <= 1) return AK_OK;
    return exchange_halo_1d(ctx, v, n, nlo, nhi, periodic, lo, hi);
}

// ghost rows for a 2-D slab: neighbours' rows (multi-GPU), own rows (periodic on one GPU) or zeros
static int ghost_rows(Ctx* ctx, const ak_problem* p, const double* v, const double** lo, const double** hi) {
    if (ctx->nranks > 1) return exchange_halo_rows(ctx, v, p->nx, p->ny, p->bc, lo, hi);
    if (p->bc == AK_BC_PERIODIC) {
        *lo = v + (p->ny - 1) * p->nx;
        *hi = v;
    } else {
        *lo = nullptr;
        *hi = nullptr;
    }
    return AK_OK;
}

// ---- pieces of the Midpoint / Trapezoid wrappers (examples/implicit.jl:17-37), composed from
//      the RHS stencil and two point-wise kernels; every rounding point of the Julia broadcasts is kept
// y = a x + b z          (uu_n .= alpha .* u_n .+ (1 - alpha) .* u        implicit.jl:20)
__global__ void __launch_bounds__(256) k_lincomb(double* __restrict__ y, double a, const double* __restrict__ x, double b,
                                                 const double* __restrict__ z, int64_t n) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += nth)
        y[j] = (x != nullptr) ? __dadd_rn(__dmul_rn(a, x[j]), __dmul_rn(b, z[j])) : __dmul_rn(b, z[j]);
}
// res = (un + c (d1 [+ d2])) - u     (res .= u_n .+ dt .* du .- u ; u_n .+ (dt/2) .* (du_n .+ du) .- u)
// JVP form (un == nullptr): out = c d1 - v
__global__ void __launch_bounds__(256) k_time_combine(double* __restrict__ res, const double* __restrict__ un, double c,
                                                      const double* d1, const double* d2, const double* u,
                                                      int64_t n) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += nth) {
        const double d = (d2 != nullptr) ? __dadd_rn(d1[j], d2[j]) : d1[j];
        const double t = __dmul_rn(c, d);
        res[j] = (un != nullptr) ? __dsub_rn(__dadd_rn(un[j], t), u[j]) : __dsub_rn(t, u[j]);
    }
}
static int ew_blocks(const Ctx* ctx, int64_t n) {
    int64_t b = (n + 255) / 256, cap = (int64_t)ctx->num_sms * 8;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

static void dg_constants(DgArgs& d, double h);
static int launch_dg(Ctx* ctx, DgArgs& d, bool residual, bool scale, int red);
template <int OP> static int launch1d(Ctx* ctx, StencilArgs& a, bool scale, int red);
template <int OP> static int launch2d(Ctx* ctx, StencilArgs& a, bool scale, int red);
static void base_args(Ctx* ctx, const ak_problem* p, StencilArgs& a);
static int ghost_rows(Ctx* ctx, const ak_problem* p, const double* v, const double** lo, const double** hi);

// du <- f(in): the bare right-hand side of heat_1D!, diffusion! or the DG heat_1D!; `in` may have its
// boundary entries overwritten by the BC code (1-D heat), exactly like f!(du, u, p, t) in the reference.
static int launch_rhs(Ctx* ctx, const ak_problem* p, double* in, double* du) {
    if (p->kind == AK_HEAT1D_DG) {
        DgArgs d{};
        dg_constants(d, p->dx);
        d.ne = p->nx / 4;
        d.c0 = 1.0; d.c1 = 1.0; d.rhs_only = 1;
        d.in = in; d.out = du;
        AK_TRY(ghost_1d(ctx, in, p->nx, 4, 1, true, &d.lo, &d.hi));
        d.partials = ctx->partials; d.ticket = ctx->ticket;
        return launch_dg(ctx, d, false, false, RED_NONE);
    }
    StencilArgs a;
    base_args(ctx, p, a);
    a.in = in;
    a.out = du;
    if (p->kind == AK_HEAT1D) {
        AK_TRY(ghost_1d(ctx, in, p->nx, 1, 1, false, &a.lo, &a.hi));
        a.in_write = in;
        return launch1d<OP_RHS_HEAT>(ctx, a, false, RED_NONE);
    }
    AK_TRY(ghost_rows(ctx, p, in, &a.lo, &a.hi));
    return launch2d<OP_RHS_HEAT>(ctx, a, false, RED_NONE);
}

struct PoolVec {  // stream-ordered scratch vector
    Ctx* c;
    double* p = nullptr;
    PoolVec(Ctx* ctx, int64_t n) : c(ctx) {
        if (pool_alloc(c, (void**)&p, sizeof(double) * (size_t)(n > 0 ? n : 1)) != cudaSuccess) p = nullptr;
    }
    ~PoolVec() { if (p) cudaFreeAsync(p, c->stream); }
};

// G_Midpoint! / G_Trapezoid! residuals
static int launch_residual_composite(Ctx* ctx, const ak_problem* p, double* u, double* res, double* sumsq_dev) {
    const int64_t n = ak_problem_size(p);
    AK_REQUIRE(p->un != nullptr, "time-dependent residual needs p->un");
    PoolVec du(ctx, n), t2(ctx, n);
    if (!du.p || !t2.p) { set_error("out of device memory for the Midpoint/Trapezoid scratch"); return AK_ERR_NOMEM; }
    const int blocks = ew_blocks(ctx, n);
    double* un = const_cast<double*>(p->un);
    if (p->scheme == AK_MIDPOINT) {
        const double al = 0.5;  // alpha default of G_Midpoint! (implicit.jl:17)
        // the reference uses `res` as the temporary uu_n (implicit.jl:18-20); the BC code mutates it, not u
        k_lincomb<<<blocks, 256, 0, ctx->stream>>>(res, al, un, 1.0 - al, u, n);
        ctx->launches++;
        AK_TRY(launch_rhs(ctx, p, res, du.p));
        k_time_combine<<<blocks, 256, 0, ctx->stream>>>(res, un, p->dt, du.p, nullptr, u, n);
        ctx->launches++;
    } else {  // AK_TRAPEZOID (implicit.jl:29-37): f!(du_n, u_n) and f!(du, u) both run their BC code in place
        AK_TRY(launch_rhs(ctx, p, un, t2.p));
        AK_TRY(launch_rhs(ctx, p, u, du.p));
        k_time_combine<<<blocks, 256, 0, ctx->stream>>>(res, un, p->dt / 2.0, t2.p, du.p, u, n);
        ctx->launches++;
    }
    AK_CUDA(cudaGetLastError());
    if (sumsq_dev) AK_TRY(launch_sumsq(ctx, n, res, sumsq_dev));
    return AK_OK;
}

// tangent of G_Midpoint!: out = dt f'((1 - alpha) v) - v ; the BC code acts on the temporary, not on v
static int launch_jvp_midpoint(Ctx* ctx, const ak_problem* p, double* v, double* out) {
    const int64_t n = ak_problem_size(p);
    PoolVec dv(ctx, n);
    if (!dv.p) { set_error("out of device memory for the Midpoint scratch"); return AK_ERR_NOMEM; }
    const int blocks = ew_blocks(ctx, n);
    k_lincomb<<<blocks, 256, 0, ctx->stream>>>(out, 0.0, nullptr, 1.0 - 0.5, v, n);
    ctx->launches++;
    AK_TRY(launch_rhs(ctx, p, out, dv.p));
    k_time_combine<<<blocks, 256, 0, ctx->stream>>>(out, nullptr, p->dt, dv.p, nullptr, v, n);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

// true when J v is the generic finite difference of the residual (AK_JVP_FD, or AK_USER without a tangent)
static inline bool fd_generic(const ak_problem* p) {
    return p->jvp_mode == AK_JVP_FD || (p->kind == AK_USER && p->user_jvp == nullptr);
}

int launch_residual(Ctx* ctx, const ak_problem* p, double* u, double* res, double* sumsq_dev) {
    AK_TRY(check_problem(p));
    AK_TRY(check_multi_gpu(ctx, p));
    if (fd_generic(p) && p->coef != nullptr) {
        // p->coef caches F(u) for the finite-difference JVPs of this Newton step (not lambda e^u)
        ak_problem q = *p;
        q.coef = nullptr;
        AK_TRY(launch_residual(ctx, &q, u, res, sumsq_dev));
        return launch_copy(ctx, ak_problem_size(p), p->coef, res);
    }
    ProfScope prof(ctx, PK_RESIDUAL);
    if (p->kind == AK_USER) {
        AK_TRY(user_residual(ctx, p, u, res));
        if (sumsq_dev) AK_TRY(launch_sumsq(ctx, p->nx, res, sumsq_dev));
        return AK_OK;
    }
    if (p->scheme == AK_MIDPOINT || p->scheme == AK_TRAPEZOID) {
        return launch_residual_composite(ctx, p, u, res, sumsq_dev);
    }
    const int red = sumsq_dev ? RED_SUMSQ : RED_NONE;
    if (p->kind == AK_SIMPLE2) {
        k_simple2<<<1, 32, 0, ctx->stream>>>(u, nullptr, res, sumsq_dev, 0, nullptr, nullptr);
        ctx->launches++;
        AK_CUDA(cudaGetLastError());
        return AK_OK;
    }
    if (p->kind == AK_HEAT1D_DG) {
        AK_REQUIRE(p->un != nullptr, "time-dependent residual needs p->un");
        DgArgs d{};
        dg_constants(d, p->dx);
        d.ne = p->nx / 4;
        d.dt = p->dt; d.c0 = 1.0; d.c1 = p->dt;
        d.in = u; d.un = p->un; d.out = res;
        AK_TRY(ghost_1d(ctx, u, p->nx, 4, 1, true, &d.lo, &d.hi));
        d.red_out = sumsq_dev; d.partials = ctx->partials; d.ticket = ctx->ticket;
        AK_REQUIRE(al(u, 32) && al(res, 32) && al(p->un, 32), "DG vectors must be 32-byte aligned");
        AK_TRY(launch_dg(ctx, d, true, false, red));
        if (sumsq_dev) AK_TRY(allreduce_sum(ctx, sumsq_dev, 1));
        return AK_OK;
    }
    StencilArgs a;
    base_args(ctx, p, a);
    a.in = u;
    a.out = res;
    a.red_out = sumsq_dev;
    int rc = AK_OK;
    switch (p->kind) {
        case AK_BRATU1D:
            AK_TRY(ghost_1d(ctx, u, p->nx, 1, 1, false, &a.lo, &a.hi));
            a.aux_out = p->coef;
            rc = launch1d<OP_RES_BRATU>(ctx, a, false, red);
            break;
        case AK_HEAT1D:
            AK_REQUIRE(p->un != nullptr, "time-dependent residual needs p->un");
            AK_TRY(ghost_1d(ctx, u, p->nx, 1, 1, false, &a.lo, &a.hi));
            a.aux = p->un;
            a.in_write = u;  // bc!(u) side effect
            rc = launch1d<OP_RES_HEAT>(ctx, a, false, red);
            break;
        case AK_BRATU2D:
            AK_TRY(ghost_rows(ctx, p, u, &a.lo, &a.hi));
            a.aux_out = p->coef;
            rc = launch2d<OP_RES_BRATU>(ctx, a, false, red);
            break;
        case AK_HEAT2D:
            AK_REQUIRE(p->un != nullptr, "time-dependent residual needs p->un");
            AK_TRY(ghost_rows(ctx, p, u, &a.lo, &a.hi));
            a.aux = p->un;
            rc = launch2d<OP_RES_HEAT>(ctx, a, false, red);
            break;
        default: break;
    }
    AK_TRY(rc);
    if (sumsq_dev) AK_TRY(allreduce_sum(ctx, sumsq_dev, 1));
    return AK_OK;
}

int launch_jvp(Ctx* ctx, const ak_problem* p, const double* u, double* v, double* out, const JvpFusion* f) {
    AK_TRY(check_problem(p));
    JvpFusion nofuse;
    if (!f) f = &nofuse;
    AK_TRY(check_multi_gpu(ctx, p));
    if (f->rhs_minus != nullptr || f->sumsq_dev != nullptr) {
        // restart residual b - J v (+ its norm): fused into the store of the 2-D tangent stencils (the 512 MiB vectors of
        // the benchmark configs); everywhere else the reference's three operations (mul!, kaxpby!, knorm)
        const bool fused2d = (p->kind == AK_BRATU2D || p->kind == AK_HEAT2D) && p->scheme != AK_MIDPOINT &&
                             p->jvp_mode == AK_JVP_ANALYTIC;
        if (!fused2d) {
            const int64_t n = ak_problem_size(p);
            JvpFusion g = *f;
            g.rhs_minus = nullptr;
            g.sumsq_dev = nullptr;
            AK_TRY(launch_jvp(ctx, p, u, v, out, &g));
            if (f->rhs_minus) AK_TRY(launch_axpby(ctx, n, 1.0, f->rhs_minus, -1.0, out));
            if (f->sumsq_dev) AK_TRY(launch_sumsq(ctx, n, out, f->sumsq_dev));
            return AK_OK;
        }
    }
    ProfScope prof(ctx, PK_JVP);
    if (p->kind == AK_USER || fd_generic(p)) {
        // caller-supplied tangent, or (F(u + eps v) - F(u)) / eps through the residual; the fused normalisation
        // and first dot of the Arnoldi step run as separate launches (same arithmetic)
        const int64_t n = ak_problem_size(p);
        if (f->scale_src) AK_TRY(launch_divcopy_dev(ctx, n, v, f->scale_src, f->denom_dev, f->stop_flag));
        if (fd_generic(p)) AK_TRY(launch_jvp_fd(ctx, p, u, v, out));
        else AK_TRY(user_jvp(ctx, p, u, v, out));
        if (f->dot_with) AK_TRY(launch_mgs_step(ctx, n, out, nullptr, nullptr, f->dot_with, 0, f->dot_dev, f->stop_flag));
        return AK_OK;
    }
    if (p->scheme == AK_MIDPOINT) {
        // composed path: fused normalisation / dot are done as separate launches (same arithmetic)
        const int64_t n = ak_problem_size(p);
        if (f->stop_flag) {
            // kernels of the composite path do not read the stop flag; results are discarded by the caller
        }
        if (f->scale_src) AK_TRY(launch_divcopy_dev(ctx, n, v, f->scale_src, f->denom_dev, f->stop_flag));
        AK_TRY(launch_jvp_midpoint(ctx, p, v, out));
        if (f->dot_with) AK_TRY(launch_mgs_step(ctx, n, out, nullptr, nullptr, f->dot_with, 0, f->dot_dev, f->stop_flag));
        return AK_OK;
    }
    // un-normalised Krylov basis (f->raw): the stored vector scale_src is the seed, J(scale_src) / denom the result
    const bool raw = f->raw && f->scale_src != nullptr;
    const bool native = !(p->kind == AK_SIMPLE2);
    const bool scale = f->scale_src != nullptr && !(raw && native);
    const bool is2d = (p->kind == AK_BRATU2D || p->kind == AK_HEAT2D);
    AK_REQUIRE(f->nproj == 0 || (is2d && p->jvp_mode == AK_JVP_ANALYTIC && f->nproj <= kBlkMax && !f->dot_with && !f->sumsq_dev),
               "launch_jvp: the projection fusion belongs to the 2-D analytic tangents");
    const int red = f->nproj > 0 ? RED_PROJ : (f->dot_with ? RED_DOT : (f->sumsq_dev ? RED_SUMSQ : RED_NONE));
    double* red_out = f->dot_with ? f->dot_dev : f->sumsq_dev;
    if (raw && native) v = const_cast<double*>(f->scale_src);
    if (p->kind == AK_SIMPLE2) {
        if (scale) AK_TRY(launch_divcopy_dev(ctx, 2, v, f->scale_src, f->denom_dev, f->stop_flag));
        k_simple2<<<1, 32, 0, ctx->stream>>>(u, v, out, red_out, 1, f->dot_with, f->stop_flag);
        ctx->launches++;
        AK_CUDA(cudaGetLastError());
        return AK_OK;
    }
    if (p->kind == AK_HEAT1D_DG) {
        DgArgs d{};
        dg_constants(d, p->dx);
        d.ne = p->nx / 4;
        d.dt = p->dt; d.c0 = 1.0; d.c1 = (p->scheme == AK_TRAPEZOID) ? p->dt / 2.0 : p->dt;
        d.in = scale ? f->scale_src : v;
        AK_TRY(ghost_1d(ctx, d.in, p->nx, 4, 1, true, &d.lo, &d.hi));
        d.in_write = v; d.denom = f->denom_dev; d.out_scale = raw ? f->denom_dev : nullptr;
        d.out_scale_inv = raw ? f->inv_denom_dev : nullptr;
        d.out = out; d.dot_with = f->dot_with; d.red_out = red_out;
        d.partials = ctx->partials; d.ticket = ctx->ticket; d.stop = f->stop_flag;
        AK_REQUIRE(al(d.in, 32) && al(v, 32) && al(out, 32) && al(f->dot_with, 32), "DG vectors must be 32-byte aligned");
        AK_TRY(launch_dg(ctx, d, false, scale, red));
        if (red) AK_TRY(allreduce_sum(ctx, red_out, 1));
        return AK_OK;
    }
    StencilArgs a;
    base_args(ctx, p, a);
    a.in = scale ? f->scale_src : v;
    a.in_write = scale ? v : nullptr;
    a.denom = f->denom_dev;
    a.out_scale = raw ? f->denom_dev : nullptr;
    a.out_scale_inv = raw ? f->inv_denom_dev : nullptr;
    a.out = out;
    a.dot_with = f->dot_with;
    a.red_out = red_out;
    a.bminus = f->rhs_minus;
    a.stop = f->stop_flag;
    bool proj_p2p = false;
    if (f->nproj > 0) {
        a.nproj = f->nproj;
        for (int b = 0; b < f->nproj; ++b) a.proj[b] = f->proj[b];
        a.proj_out = f->proj_out;
        proj_p2p = f->proj_comm != nullptr && ctx->p2p_on && ctx->nranks > 1;
        if (proj_p2p) {
            a.pd = ctx->p2p_dev();
            a.seq_out = f->proj_comm->seq_out;
        }
    }
    int rc = AK_OK;
    switch (p->kind) {
        case AK_BRATU1D:
            if (p->jvp_mode == AK_JVP_FD_FUSED) {
                // north_star (2) in 1-D: one pass reads u and v and writes J v ~ (F(u + eps v) - F(u)) / eps
                if (ctx->nranks > 1) { set_error("AK_JVP_FD_FUSED is single-GPU in this version"); return AK_ERR_UNSUPPORTED; }
                AK_REQUIRE(u != nullptr, "AK_JVP_FD_FUSED needs u");
                a.aux = u;
                a.fd_eps = p->fd_eps > 0.0 ? p->fd_eps : 1.4901161193847656e-08;
                rc = launch1d<OP_JVP_BRATU_FD>(ctx, a, scale, red);
                break;
            }
            AK_TRY(ghost_1d(ctx, a.in, p->nx, 1, 1, false, &a.lo, &a.hi));
            a.aux = p->coef ? p->coef : u;
            a.coef_from_u = p->coef ? 0 : 1;
            rc = launch1d<OP_JVP_BRATU>(ctx, a, scale, red);
            break;
        case AK_HEAT1D:
            AK_TRY(ghost_1d(ctx, a.in, p->nx, 1, 1, false, &a.lo, &a.hi));
            if (!scale) a.in_write = v;  // tangent of bc!(u): boundary entries of v are overwritten
            rc = launch1d<OP_JVP_HEAT>(ctx, a, scale, red);
            break;
        case AK_BRATU2D:
            if (p->jvp_mode == AK_JVP_FD_FUSED) {
                // north_star (2): one pass reads u and v and writes J v ~ (F(u + eps v) - F(u)) / eps
                if (ctx->nranks > 1) { set_error("AK_JVP_FD_FUSED is single-GPU in this version"); return AK_ERR_UNSUPPORTED; }
                AK_REQUIRE(u != nullptr, "AK_JVP_FD_FUSED needs u");
                AK_TRY(ghost_rows(ctx, p, a.in, &a.lo, &a.hi));  // (single GPU: no exchange)
                a.aux = u;
                AK_TRY(ghost_rows(ctx, p, u, &a.aux_lo, &a.aux_hi));
                a.fd_eps = p->fd_eps > 0.0 ? p->fd_eps : 1.4901161193847656e-08;
                rc = launch2d<OP_JVP_BRATU_FD>(ctx, a, scale, red);
                break;
            }
            if (f->halo_given) { a.lo = f->halo_lo; a.hi = f->halo_hi; }
            else AK_TRY(ghost_rows(ctx, p, a.in, &a.lo, &a.hi));
            a.aux = p->coef ? p->coef : u;
            a.coef_from_u = p->coef ? 0 : 1;
            rc = launch2d<OP_JVP_BRATU>(ctx, a, scale, red);
            break;
        case AK_HEAT2D:
            if (f->halo_given) { a.lo = f->halo_lo; a.hi = f->halo_hi; }
            else AK_TRY(ghost_rows(ctx, p, a.in, &a.lo, &a.hi));
            rc = launch2d<OP_JVP_HEAT>(ctx, a, scale, red);
            break;
        default: break;
    }
    AK_TRY(rc);
    if (red == RED_PROJ) {
        if (!proj_p2p) AK_TRY(allreduce_sum(ctx, f->proj_out, f->nproj));
    } else if (red) {
        AK_TRY(allreduce_sum(ctx, red_out, 1));
    }
    return AK_OK;
}

int launch_jvp_batched_bratu(Ctx* ctx, const ak_problem* p, const double* u, const double* V, int64_t ldv, double* Out,
                             int64_t ldo, int32_t ncols) {
    AK_TRY(check_problem(p));
    BatchArgs a{};
    a.nx = p->nx;
    a.ny = p->kind == AK_BRATU2D ? p->ny : 1;
    a.dx2 = p->dx * p->dx;
    a.dy2 = p->dy * p->dy;
    a.lambda = p->lambda;
    a.aux = p->coef ? p->coef : u;
    a.coef_from_u = p->coef ? 0 : 1;
    a.V = V; a.ldv = ldv; a.Out = Out; a.ldo = ldo; a.ncols = ncols;
    int vec = 1;
    auto ok = [&](int v) {
        return a.nx % v == 0 && ldv % v == 0 && ldo % v == 0 && al(a.aux, 8 * v) && al(V, 8 * v) && al(Out, 8 * v);
    };
    if (ok(4)) vec = 4;
    else if (ok(2)) vec = 2;
    ProfScope prof(ctx, PK_JVP);
    if (p->kind == AK_BRATU2D) {
        const int64_t gx = (a.nx + (int64_t)kTX * vec - 1) / ((int64_t)kTX * vec);
        const int64_t gy = (a.ny + kBatchRY - 1) / kBatchRY;
        if (gy > 65535) { set_error("ak_jvp_batched: more than 65535 row tiles"); return AK_ERR_UNSUPPORTED; }
        const int64_t gz = (ncols + kCB - 1) / kCB;
        if (gz > 65535) { set_error("ak_jvp_batched: more than 65535 column groups"); return AK_ERR_UNSUPPORTED; }
        dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
        if (vec == 4) k_bratu2d_jvp_batched<4><<<grid, kTX, 0, ctx->stream>>>(a);
        else if (vec == 2) k_bratu2d_jvp_batched<2><<<grid, kTX, 0, ctx->stream>>>(a);
        else k_bratu2d_jvp_batched<1><<<grid, kTX, 0, ctx->stream>>>(a);
    } else {
        int64_t grid = (a.nx + (int64_t)kT1 * vec - 1) / ((int64_t)kT1 * vec);
        const int64_t cap = (int64_t)ctx->num_sms * 2;
        if (grid > cap) grid = cap;
        if (vec == 4) k_bratu1d_jvp_batched<4><<<(int)grid, kT1, 0, ctx->stream>>>(a);
        else if (vec == 2) k_bratu1d_jvp_batched<2><<<(int)grid, kT1, 0, ctx->stream>>>(a);
        else k_bratu1d_jvp_batched<1><<<(int)grid, kT1, 0, ctx->stream>>>(a);
    }
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

// y[j] = x[j] * s[j % 4]   (node-wise scaling of a DG vector: 4 LGL nodes per element)
__global__ void __launch_bounds__(256) k_scale_nodes4(double* __restrict__ y, const double* __restrict__ x, double s0,
                                                      double s1, double s2, double s3, int64_t n) {
    const double s[4] = {s0, s1, s2, s3};
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += nth) y[j] = __dmul_rn(x[j], s[j & 3]);
}

int launch_jvp_transpose(Ctx* ctx, const ak_problem* p, const double* u, double* v, double* out) {
    AK_TRY(check_problem(p));
    if (p->kind == AK_HEAT1D_DG) {
        // Upwind SBP duality  M D+ + D-^T M = 0  (M = (h/2) diag(w), w = LGL weights 1/6, 5/6, 5/6, 1/6) gives
        // (D- D+)^T = M D- D+ M^-1, hence J^T = W J W^-1 for every time discretisation (J = c D- D+ - I):
        // two node-wise scalings around the forward tangent kernel.
        const int64_t n = p->nx;
        PoolVec t(ctx, n);
        if (!t.p) { set_error("out of device memory for the DG transpose scratch"); return AK_ERR_NOMEM; }
        const int blocks = ew_blocks(ctx, n);
        const double w0 = 1.0 / 6.0, w1 = 5.0 / 6.0;
        k_scale_nodes4<<<blocks, 256, 0, ctx->stream>>>(t.p, v, 1.0 / w0, 1.0 / w1, 1.0 / w1, 1.0 / w0, n);
        ctx->launches++;
        AK_TRY(launch_jvp(ctx, p, u, t.p, out, nullptr));
        k_scale_nodes4<<<blocks, 256, 0, ctx->stream>>>(out, out, w0, w1, w1, w0, n);
        ctx->launches++;
        AK_CUDA(cudaGetLastError());
        return AK_OK;
    }
    switch (p->kind) {
        case AK_SIMPLE2:
            k_simple2<<<1, 32, 0, ctx->stream>>>(u, v, out, nullptr, 2, nullptr, nullptr);
            ctx->launches++;
            AK_CUDA(cudaGetLastError());
            return AK_OK;
        case AK_BRATU1D:
        case AK_BRATU2D:
        case AK_HEAT2D:
            // Laplacian (Dirichlet-0 or periodic) + diagonal: J is symmetric, J^T v == J v
            return launch_jvp(ctx, p, u, v, out, nullptr);
        case AK_HEAT1D:
            if (p->bc == AK_BC_ZERO) return launch_jvp(ctx, p, u, v, out, nullptr);  // zero boundary rows/cols: symmetric
            if (p->scheme != AK_MIDPOINT && ctx->nranks == 1 && p->nx >= 4) {
                // periodic_bc! makes J = (c1 a L - I) B non-symmetric: dedicated adjoint kernel
                const double c1 = (p->scheme == AK_TRAPEZOID) ? p->dt / 2.0 : p->dt;
                k_heat1d_periodic_transpose<<<ew_blocks(ctx, p->nx), 256, 0, ctx->stream>>>(v, out, p->nx, p->a,
                                                                                          p->dx * p->dx, c1);
                ctx->launches++;
                AK_CUDA(cudaGetLastError());
                return AK_OK;
            }
            break;
        default: break;
    }
    set_error("ak_jvp_transpose: not implemented for kind %d bc %d", p->kind, p->bc);
    return AK_ERR_UNSUPPORTED;
}

}  // namespace ak

using namespace ak;

AK_API int64_t ak_problem_size(const ak_problem* p) {
    if (!p) return 0;
    switch (p->kind) {
        case AK_SIMPLE2: return 2;
        case AK_BRATU1D: case AK_HEAT1D: case AK_HEAT1D_DG: case AK_USER: return p->nx;
        default: return p->nx * p->ny;
    }
}

AK_API int ak_residual(ak_ctx* ctx, const ak_problem* p, double* u, double* res, double* nrm_out_host) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && p && u && res, "ak_residual: NULL argument");
    Ctx* c = &ctx->c;
    AK_TRY(launch_residual(c, p, u, res, nrm_out_host ? c->dscal : nullptr));
    if (nrm_out_host) {
        AK_CUDA(cudaMemcpyAsync(c->hscal, c->dscal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        AK_CUDA(cudaStreamSynchronize(c->stream));
        *nrm_out_host = sqrt(c->hscal[0]);
    }
    return AK_OK;
}

AK_API int ak_jvp(ak_ctx* ctx, const ak_problem* p, const double* u, double* v, double* out) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && p && v && out, "ak_jvp: NULL argument");
    return launch_jvp(&ctx->c, p, u, v, out, nullptr);
}

AK_API int ak_jvp_batched(ak_ctx* ctx, const ak_problem* p, const double* u, double* V, int64_t ldv, double* Out,
                          int64_t ldo, int32_t ncols) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && p && V && Out && ncols >= 0, "ak_jvp_batched: bad argument");
    const int64_t n = ak_problem_size(p);
    AK_REQUIRE(ldv >= n && ldo >= n, "ak_jvp_batched: leading dimensions must be >= n");
    Ctx* c_ = &ctx->c;
    // Bratu tangents share lambda e^u between the columns: multi-RHS kernels (one launch, the shared operand read once)
    const bool bratu = (p->kind == AK_BRATU2D || p->kind == AK_BRATU1D) && p->jvp_mode == AK_JVP_ANALYTIC &&
                       c_->nranks == 1 && ncols >= 2 && (p->coef != nullptr || u != nullptr);
    if (bratu) {
        AK_TRY(launch_jvp_batched_bratu(c_, p, u, V, ldv, Out, ldo, ncols));
        return AK_OK;
    }
    for (int32_t c = 0; c < ncols; ++c) AK_TRY(launch_jvp(c_, p, u, V + (int64_t)c * ldv, Out + (int64_t)c * ldo, nullptr));
    return AK_OK;
}

AK_API int ak_jvp_transpose(ak_ctx* ctx, const ak_problem* p, const double* u, double* v, double* out) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && p && v && out, "ak_jvp_transpose: NULL argument");
    return launch_jvp_transpose(&ctx->c, p, u, v, out);
}
