// BLAS-1 / Arnoldi vector kernels (fp64, HBM-bound, no tensor cores).
//
// Replaces, on the hot path of the reference:
//   Krylov.kdot/knorm/kscal!/kaxpy!/kaxpby!/kcopy!/kfill!/kref!  examples/halovector.jl:51-147
//   the modified Gram-Schmidt loop of Krylov.jl's gmres! (R[nr+i] = kdot(V[i], q);
//   kaxpy!(-R[nr+i], V[i], q); Hbis = knorm(q); V[k+1] = q / Hbis), call site
//   src/Ariadne.jl:338.
//
// All kernels stream 256-bit vectors (LDG.E.256) when the operands are 32-byte aligned
// and fall back to scalar accesses otherwise.  Reductions are deterministic: warp
// butterflies, per-block partials, last block sums partials in index order.
#include "ak_internal.h"
#include "common.cuh"

namespace ak {

constexpr int kThreads = 256;

static inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

static inline int stream_blocks(const Ctx* ctx, int64_t n, int per_thread) {
    int64_t need = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
    int64_t cap = (int64_t)ctx->num_sms * 4;  // 4 x 256 threads resident per SM
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// -----------------------------------------------------------------------------------
// Fused modified Gram-Schmidt step:
//   AXPY : w <- w - h * vi            (h read from device memory)
//   RED=1: out = <vnext, w_new>       RED=2: out = <w_new, w_new>      RED=0: no reduction
// Algorithmic bytes per launch: AXPY+dot 32n, AXPY+sumsq 24n, dot only 16n, sumsq only 8n.
// -----------------------------------------------------------------------------------
template <bool AXPY, int RED, bool VEC>
__global__ void __launch_bounds__(kThreads) k_mgs_step(double* __restrict__ w, const double* __restrict__ vi,
                                                       const double* __restrict__ h_in,
                                                       const double* __restrict__ vnext, double* __restrict__ out,
                                                       double* __restrict__ partials, unsigned int* ticket,
                                                       int64_t n, const int* __restrict__ stop) {
    __shared__ double sh[32];
    if (stop != nullptr && *stop != 0) return;
    const double h = AXPY ? -(*h_in) : 0.0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t n4 = n >> 2;
        int64_t i = tid;
        // two independent 256-bit requests per stream in flight per thread
        for (; i + nth < n4; i += 2 * nth) {
            const int64_t j0 = i << 2, j1 = (i + nth) << 2;
            d4 w0 = (AXPY ? ld4(w + j0) : ld4_stream(w + j0));
            d4 w1 = (AXPY ? ld4(w + j1) : ld4_stream(w + j1));
            if (AXPY) {
                d4 x0 = ld4_stream(vi + j0), x1 = ld4_stream(vi + j1);
                w0.x = fma(h, x0.x, w0.x); w0.y = fma(h, x0.y, w0.y); w0.z = fma(h, x0.z, w0.z); w0.w = fma(h, x0.w, w0.w);
                w1.x = fma(h, x1.x, w1.x); w1.y = fma(h, x1.y, w1.y); w1.z = fma(h, x1.z, w1.z); w1.w = fma(h, x1.w, w1.w);
                st4(w + j0, w0);
                st4(w + j1, w1);
            }
            if (RED == 1) {
                d4 y0 = ld4_stream(vnext + j0), y1 = ld4_stream(vnext + j1);
                a0 = fma(y0.x, w0.x, a0); a1 = fma(y0.y, w0.y, a1); a2 = fma(y0.z, w0.z, a2); a3 = fma(y0.w, w0.w, a3);
                a0 = fma(y1.x, w1.x, a0); a1 = fma(y1.y, w1.y, a1); a2 = fma(y1.z, w1.z, a2); a3 = fma(y1.w, w1.w, a3);
            } else if (RED == 2) {
                a0 = fma(w0.x, w0.x, a0); a1 = fma(w0.y, w0.y, a1); a2 = fma(w0.z, w0.z, a2); a3 = fma(w0.w, w0.w, a3);
                a0 = fma(w1.x, w1.x, a0); a1 = fma(w1.y, w1.y, a1); a2 = fma(w1.z, w1.z, a2); a3 = fma(w1.w, w1.w, a3);
            }
        }
        for (; i < n4; i += nth) {
            const int64_t j0 = i << 2;
            d4 w0 = (AXPY ? ld4(w + j0) : ld4_stream(w + j0));
            if (AXPY) {
                d4 x0 = ld4_stream(vi + j0);
                w0.x = fma(h, x0.x, w0.x); w0.y = fma(h, x0.y, w0.y); w0.z = fma(h, x0.z, w0.z); w0.w = fma(h, x0.w, w0.w);
                st4(w + j0, w0);
            }
            if (RED == 1) {
                d4 y0 = ld4_stream(vnext + j0);
                a0 = fma(y0.x, w0.x, a0); a1 = fma(y0.y, w0.y, a1); a2 = fma(y0.z, w0.z, a2); a3 = fma(y0.w, w0.w, a3);
            } else if (RED == 2) {
                a0 = fma(w0.x, w0.x, a0); a1 = fma(w0.y, w0.y, a1); a2 = fma(w0.z, w0.z, a2); a3 = fma(w0.w, w0.w, a3);
            }
        }
        // scalar tail (n % 4 elements), one thread each
        const int64_t j = (n4 << 2) + tid;
        if (j < n) {
            double wj = w[j];
            if (AXPY) { wj = fma(h, vi[j], wj); w[j] = wj; }
            if (RED == 1) a0 = fma(vnext[j], wj, a0);
            else if (RED == 2) a0 = fma(wj, wj, a0);
        }
    } else {
        for (int64_t j = tid; j < n; j += nth) {
            double wj = w[j];
            if (AXPY) { wj = fma(h, vi[j], wj); w[j] = wj; }
            if (RED == 1) a0 = fma(vnext[j], wj, a0);
            else if (RED == 2) a0 = fma(wj, wj, a0);
        }
    }
    if (RED != 0) {
        double s = block_sum((a0 + a1) + (a2 + a3), sh);
        grid_sum_finish(s, partials, ticket, blockIdx.x, gridDim.x, out, sh);
    }
}

template <bool AXPY, int RED>
static int launch_mgs_t(Ctx* ctx, int64_t n, double* w, const double* vi, const double* h_in, const double* vnext,
                        double* out, const int* stop) {
    const bool vec = aligned32(w) && (!AXPY || aligned32(vi)) && (RED != 1 || aligned32(vnext));
    const int blocks = stream_blocks(ctx, n, 8);
    constexpr int cls = AXPY ? (RED == 1 ? PK_MGS_AXPY_DOT : (RED == 2 ? PK_MGS_AXPY_NRM : PK_MGS_AXPY))
                             : (RED == 1 ? PK_DOT : PK_SUMSQ);
    {
        ProfScope prof(ctx, cls);  // brackets the kernel only (not the all-reduce that follows)
        if (vec)
            k_mgs_step<AXPY, RED, true><<<blocks, kThreads, 0, ctx->stream>>>(w, vi, h_in, vnext, out, ctx->partials,
                                                                              ctx->ticket, n, stop);
        else
            k_mgs_step<AXPY, RED, false><<<blocks, kThreads, 0, ctx->stream>>>(w, vi, h_in, vnext, out,
                                                                               ctx->partials, ctx->ticket, n, stop);
    }
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    if (RED != 0) AK_TRY(allreduce_sum(ctx, out, 1));
    return AK_OK;
}

int launch_mgs_step(Ctx* ctx, int64_t n, double* w, const double* vi, const double* h_in, const double* vnext,
                    int want_sumsq, double* out_dev, const int* stop) {
    if (n <= 0) return AK_OK;
    if (vi) {
        if (vnext) return launch_mgs_t<true, 1>(ctx, n, w, vi, h_in, vnext, out_dev, stop);
        if (want_sumsq) return launch_mgs_t<true, 2>(ctx, n, w, vi, h_in, nullptr, out_dev, stop);
        return launch_mgs_t<true, 0>(ctx, n, w, vi, h_in, nullptr, nullptr, stop);
    }
    if (vnext) return launch_mgs_t<false, 1>(ctx, n, w, nullptr, nullptr, vnext, out_dev, stop);
    if (want_sumsq) return launch_mgs_t<false, 2>(ctx, n, w, nullptr, nullptr, nullptr, out_dev, stop);
    return AK_OK;
}

// -----------------------------------------------------------------------------------
// Pair-wise modified Gram-Schmidt pass (fuse level PAIR): two Gram-Schmidt steps per sweep over w.
//   NAX  axpys : w <- w - h_a v_a [- h_b v_b]     with (h_a, h_b) = (t[0], t[1] - t[0] t[2]) from `tin`
//   NRED = 1   : out[0] = <y_a, w_new>
//   NRED = 2   : out = { <y_a,w_new>, <y_b,w_new>, <y_b,y_a> }   (h of y_b follows as d2 - d1 g: algebraically the
//                modified Gram-Schmidt coefficient <y_b, w_new - d1 y_a>, without a second sweep over w)
//   NRED = 3   : out[0] = <w_new, w_new>
// Algorithmic bytes per launch: 8n (2 [w in/out] + NAX + number of y vectors); 48n for NAX = NRED = 2,
// i.e. 24n per Gram-Schmidt step instead of 32n.
// -----------------------------------------------------------------------------------
struct PairP2P {      // fused collective of the pair-wise pass (all zero / null when not used)
    P2PDev pd;
    unsigned long long seq_in;   // record to wait for (reduced sums of the pair being subtracted), 0 = read `tin`
    unsigned long long seq_out;  // record to post (this pass's sums), 0 = store to `out`
    double* tin_store;           // where block 0 leaves the reduced incoming sums for the Givens kernel
    P2PHalo halo;                // boundary-row push (final pass only)
};

template <int NAX, int NRED, bool VEC, bool P2P>
__global__ void __launch_bounds__(kThreads) k_mgs_pair(double* __restrict__ w, const double* __restrict__ va,
                                                       const double* __restrict__ vb, const double* __restrict__ tin,
                                                       const double* __restrict__ ya, const double* __restrict__ yb,
                                                       double* __restrict__ out, double* __restrict__ partials,
                                                       unsigned int* ticket, int64_t n, const int* __restrict__ stop,
                                                       const PairP2P pp) {
    __shared__ double sh[32];
    __shared__ double sh4[4 * kMaxPeers];
    if (stop != nullptr && *stop != 0) return;
    double ha = 0.0, hb = 0.0;
    if (NAX >= 1) {
        double t[3];
        if (P2P && pp.seq_in != 0) {
            mail_wait_sum(pp.pd, pp.seq_in, t, sh4);
            if (blockIdx.x == 0 && threadIdx.x == 0 && pp.tin_store != nullptr) {
                pp.tin_store[0] = t[0]; pp.tin_store[1] = t[1]; pp.tin_store[2] = t[2];
            }
        } else {
            t[0] = tin[0]; t[1] = (NAX == 2) ? tin[1] : 0.0; t[2] = (NAX == 2) ? tin[2] : 0.0;
        }
        ha = -t[0];
        if (NAX == 2) hb = -pair_second_h(t[0], t[1], t[2]);
    }
    double s0a = 0.0, s0b = 0.0, s1a = 0.0, s1b = 0.0, s2a = 0.0, s2b = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const bool push = P2P && NRED == 3 && (pp.halo.down_hi != nullptr || pp.halo.up_lo != nullptr);
    const int64_t hnx = pp.halo.nx;
    auto body = [&](double& wj, double xa, double xb, double a, double b, double& t0, double& t1, double& t2) {
        if (NAX >= 1) wj = fma(ha, xa, wj);   // same order as two successive kaxpy!
        if (NAX == 2) wj = fma(hb, xb, wj);
        if (NRED == 1) t0 = fma(a, wj, t0);
        if (NRED == 2) { t0 = fma(a, wj, t0); t1 = fma(b, wj, t1); t2 = fma(b, a, t2); }
        if (NRED == 3) t0 = fma(wj, wj, t0);
    };
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            const int64_t j = i << 2;
            d4 wv = (NAX > 0) ? ld4(w + j) : ld4_stream(w + j);
            d4 xa = {0, 0, 0, 0}, xb = {0, 0, 0, 0}, a = {0, 0, 0, 0}, b = {0, 0, 0, 0};
            if (NAX >= 1) xa = ld4_stream(va + j);
            if (NAX == 2) xb = ld4_stream(vb + j);
            if (NRED == 1 || NRED == 2) a = ld4_stream(ya + j);
            if (NRED == 2) b = ld4_stream(yb + j);
            body(wv.x, xa.x, xb.x, a.x, b.x, s0a, s1a, s2a);
            body(wv.y, xa.y, xb.y, a.y, b.y, s0b, s1b, s2b);
            body(wv.z, xa.z, xb.z, a.z, b.z, s0a, s1a, s2a);
            body(wv.w, xa.w, xb.w, a.w, b.w, s0b, s1b, s2b);
            if (NAX > 0) st4(w + j, wv);
            if (push) {  // boundary rows of the finished w go straight into the neighbours' ghost rows (NVLink)
                if (pp.halo.down_hi != nullptr && j < hnx) st4(pp.halo.down_hi + j, wv);
                if (pp.halo.up_lo != nullptr && j >= n - hnx) st4(pp.halo.up_lo + (j - (n - hnx)), wv);
            }
        }
        const int64_t j = (n4 << 2) + tid;
        if (j < n) {
            double wj = w[j];
            body(wj, NAX >= 1 ? va[j] : 0.0, NAX == 2 ? vb[j] : 0.0, (NRED == 1 || NRED == 2) ? ya[j] : 0.0,
                 NRED == 2 ? yb[j] : 0.0, s0a, s1a, s2a);
            if (NAX > 0) w[j] = wj;
        }
    } else {
        for (int64_t j = tid; j < n; j += nth) {
            double wj = w[j];
            body(wj, NAX >= 1 ? va[j] : 0.0, NAX == 2 ? vb[j] : 0.0, (NRED == 1 || NRED == 2) ? ya[j] : 0.0,
                 NRED == 2 ? yb[j] : 0.0, s0a, s1a, s2a);
            if (NAX > 0) w[j] = wj;
        }
    }
    if (P2P && pp.seq_out != 0) {
        const double r0 = block_sum(s0a + s0b, sh);
        const double r1 = (NRED == 2) ? block_sum(s1a + s1b, sh) : 0.0;
        const double r2 = (NRED == 2) ? block_sum(s2a + s2b, sh) : 0.0;
        double tot[3];
        if (grid_reduce3(r0, r1, r2, partials, ticket, blockIdx.x, gridDim.x, sh, tot, push))
            mail_post(pp.pd, pp.seq_out, tot[0], tot[1], tot[2]);
    } else if (NRED == 2) {
        const double r0 = block_sum(s0a + s0b, sh);
        const double r1 = block_sum(s1a + s1b, sh);
        const double r2 = block_sum(s2a + s2b, sh);
        grid_sum_finish3(r0, r1, r2, partials, ticket, blockIdx.x, gridDim.x, out, sh);
    } else if (NRED != 0) {
        const double r0 = block_sum(s0a + s0b, sh);
        grid_sum_finish(r0, partials, ticket, blockIdx.x, gridDim.x, out, sh);
    }
}

// va/vb: vectors to subtract (0, 1 or 2 non-null), tin: raw triple of that pair; ya/yb: vectors to project on
// (0, 1 or 2 non-null); want_sumsq: ||w_new||^2 instead.  out receives 3 doubles (NRED = 2) or 1.
// With `p2p` (multi-GPU, peer memory enabled) the sums travel through the peers' mailboxes instead of NCCL.
int launch_mgs_pair(Ctx* ctx, int64_t n, double* w, const double* va, const double* vb, const double* tin,
                    const double* ya, const double* yb, int want_sumsq, double* out, const int* stop,
                    const PairComm* pc) {
    if (n <= 0) return AK_OK;
    const int nax = va ? (vb ? 2 : 1) : 0;
    const int nred = want_sumsq ? 3 : (ya ? (yb ? 2 : 1) : 0);
    const bool vec = aligned32(w) && aligned32(va) && aligned32(vb) && aligned32(ya) && aligned32(yb);
    const int blocks = stream_blocks(ctx, n, 4);
    const int cls = nax == 2 && nred == 2 ? PK_MGS_PAIR : (nax > 0 ? (nred == 3 ? PK_MGS_AXPY_NRM : PK_MGS_PAIR_EDGE)
                                                                       : PK_MGS_PAIR_EDGE);
    const bool p2p = pc != nullptr && ctx->p2p_on && ctx->nranks > 1;
    PairP2P pp{};
    if (p2p) {
        pp.pd = ctx->p2p_dev();
        pp.seq_in = pc->seq_in;
        pp.seq_out = pc->seq_out;
        pp.tin_store = pc->tin_store;
        pp.halo = pc->halo;
        // the vector path pushes whole 256-bit words: rows must be 32-byte multiples
        if (pp.halo.nx % 4 != 0 || n % 4 != 0 || !vec) pp.halo.down_hi = pp.halo.up_lo = nullptr;
    }
#define AK_PAIR(A, R)                                                                                               \
    if (nax == A && nred == R) {                                                                                    \
        ProfScope prof(ctx, cls);                                                                                   \
        if (p2p && vec)                                                                                             \
            k_mgs_pair<A, R, true, true><<<blocks, kThreads, 0, ctx->stream>>>(w, va, vb, tin, ya, yb, out,        \
                                                                               ctx->partials, ctx->ticket, n, stop, pp); \
        else if (p2p)                                                                                               \
            k_mgs_pair<A, R, false, true><<<blocks, kThreads, 0, ctx->stream>>>(w, va, vb, tin, ya, yb, out,       \
                                                                                ctx->partials, ctx->ticket, n, stop, pp); \
        else if (vec)                                                                                               \
            k_mgs_pair<A, R, true, false><<<blocks, kThreads, 0, ctx->stream>>>(w, va, vb, tin, ya, yb, out,       \
                                                                                ctx->partials, ctx->ticket, n, stop, pp); \
        else                                                                                                        \
            k_mgs_pair<A, R, false, false><<<blocks, kThreads, 0, ctx->stream>>>(w, va, vb, tin, ya, yb, out,      \
                                                                                 ctx->partials, ctx->ticket, n, stop, pp); \
    }
    AK_PAIR(0, 1) AK_PAIR(0, 2) AK_PAIR(1, 1) AK_PAIR(1, 2) AK_PAIR(1, 3) AK_PAIR(2, 1) AK_PAIR(2, 2) AK_PAIR(2, 3)
#undef AK_PAIR
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    if (!p2p) {
        if (nred == 2) AK_TRY(allreduce_sum(ctx, out, 3));
        else if (nred != 0) AK_TRY(allreduce_sum(ctx, out, 1));
    }
    return AK_OK;
}

int launch_dot(Ctx* ctx, int64_t n, const double* x, const double* y, double* out_dev) {
    if (n <= 0) return launch_fill(ctx, 1, out_dev, 0.0);
    // <y, x>: w is only read when AXPY == false
    return launch_mgs_t<false, 1>(ctx, n, const_cast<double*>(x), nullptr, nullptr, y, out_dev, nullptr);
}
int launch_sumsq(Ctx* ctx, int64_t n, const double* x, double* out_dev) {
    if (n <= 0) return launch_fill(ctx, 1, out_dev, 0.0);
    return launch_mgs_t<false, 2>(ctx, n, const_cast<double*>(x), nullptr, nullptr, nullptr, out_dev, nullptr);
}

// -----------------------------------------------------------------------------------
// Element-wise streams:  y <- op(x, y)
// -----------------------------------------------------------------------------------
enum { OP_SCAL, OP_AXPY, OP_AXPBY, OP_COPY, OP_FILL, OP_DIVCOPY, OP_REF };

template <int OP>
AK_DEV void ew_apply(double& x, double& y, double s, double t) {
    if (OP == OP_SCAL) y = s * y;
    else if (OP == OP_AXPY) y = fma(s, x, y);
    else if (OP == OP_AXPBY) y = fma(s, x, t * y);
    else if (OP == OP_COPY) y = x;
    else if (OP == OP_FILL) y = s;
    else if (OP == OP_DIVCOPY) y = x / s;
    else if (OP == OP_REF) {  // kref!: x <- c x + s y ; y <- s x - c y   (c = s-arg, s = t-arg)
        const double xi = x, yi = y;
        x = fma(s, xi, t * yi);
        y = fma(t, xi, -(s * yi));
    }
}

// s_dev (optional): scalar taken from device memory, multiplied by `s` (used as a sign).
template <int OP, bool VEC>
__global__ void __launch_bounds__(kThreads) k_ew(double* __restrict__ y, double* x, double s, double t,
                                                 const double* __restrict__ s_dev, int64_t n,
                                                 const int* __restrict__ stop) {
    if (stop != nullptr && *stop != 0) return;
    if (s_dev != nullptr) s = s * (*s_dev);
    constexpr bool READ_X = (OP == OP_AXPY || OP == OP_AXPBY || OP == OP_COPY || OP == OP_DIVCOPY || OP == OP_REF);
    constexpr bool READ_Y = (OP == OP_SCAL || OP == OP_AXPY || OP == OP_AXPBY || OP == OP_REF);
    constexpr bool WRITE_X = (OP == OP_REF);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            const int64_t j = i << 2;
            d4 xv = {0, 0, 0, 0}, yv = {0, 0, 0, 0};
            if (READ_X) xv = WRITE_X ? ld4(x + j) : ld4_stream(x + j);
            if (READ_Y) yv = ld4(y + j);
            ew_apply<OP>(xv.x, yv.x, s, t);
            ew_apply<OP>(xv.y, yv.y, s, t);
            ew_apply<OP>(xv.z, yv.z, s, t);
            ew_apply<OP>(xv.w, yv.w, s, t);
            st4(y + j, yv);
            if (WRITE_X) st4(x + j, xv);
        }
        const int64_t j = (n4 << 2) + tid;
        if (j < n) {
            double xv = READ_X ? x[j] : 0.0, yv = READ_Y ? y[j] : 0.0;
            ew_apply<OP>(xv, yv, s, t);
            y[j] = yv;
            if (WRITE_X) x[j] = xv;
        }
    } else {
        for (int64_t j = tid; j < n; j += nth) {
            double xv = READ_X ? x[j] : 0.0, yv = READ_Y ? y[j] : 0.0;
            ew_apply<OP>(xv, yv, s, t);
            y[j] = yv;
            if (WRITE_X) x[j] = xv;
        }
    }
}

template <int OP>
static int launch_ew(Ctx* ctx, int64_t n, double* y, const double* x, double s, double t, const double* s_dev,
                     const int* stop) {
    if (n <= 0) return AK_OK;
    const bool vec = aligned32(y) && (x == nullptr || aligned32(x));
    const int blocks = stream_blocks(ctx, n, 4);
    ProfScope prof(ctx, PK_ELEMENTWISE);
    if (vec)
        k_ew<OP, true><<<blocks, kThreads, 0, ctx->stream>>>(y, const_cast<double*>(x), s, t, s_dev, n, stop);
    else
        k_ew<OP, false><<<blocks, kThreads, 0, ctx->stream>>>(y, const_cast<double*>(x), s, t, s_dev, n, stop);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

int launch_scal(Ctx* ctx, int64_t n, double s, double* x) { return launch_ew<OP_SCAL>(ctx, n, x, nullptr, s, 0, nullptr, nullptr); }
int launch_axpy(Ctx* ctx, int64_t n, double s, const double* x, double* y) { return launch_ew<OP_AXPY>(ctx, n, y, x, s, 0, nullptr, nullptr); }
int launch_axpy_dev(Ctx* ctx, int64_t n, const double* s_dev, double sign, const double* x, double* y) {
    return launch_ew<OP_AXPY>(ctx, n, y, x, sign, 0, s_dev, nullptr);
}
int launch_axpby(Ctx* ctx, int64_t n, double s, const double* x, double t, double* y) { return launch_ew<OP_AXPBY>(ctx, n, y, x, s, t, nullptr, nullptr); }
int launch_copy(Ctx* ctx, int64_t n, double* y, const double* x) { return launch_ew<OP_COPY>(ctx, n, y, x, 0, 0, nullptr, nullptr); }
int launch_fill(Ctx* ctx, int64_t n, double* x, double v) { return launch_ew<OP_FILL>(ctx, n, x, nullptr, v, 0, nullptr, nullptr); }
int launch_ref(Ctx* ctx, int64_t n, double* x, double* y, double c, double s) { return launch_ew<OP_REF>(ctx, n, y, x, c, s, nullptr, nullptr); }
int launch_divcopy(Ctx* ctx, int64_t n, double* y, const double* x, double s) { return launch_ew<OP_DIVCOPY>(ctx, n, y, x, s, 0, nullptr, nullptr); }
int launch_divcopy_dev(Ctx* ctx, int64_t n, double* y, const double* x, const double* s_dev, const int* stop) {
    return launch_ew<OP_DIVCOPY>(ctx, n, y, x, 1.0, 0, s_dev, stop);
}

// -----------------------------------------------------------------------------------
// x <- [x +] sum_{i<k} y[i] V[i]   in the sequential axpy order of gmres! step 10
// (for i = 1:k  kaxpy!(n, y[i], V[i], xr)).  Algorithmic bytes 8n(k+1) (+8n when accumulating).
// -----------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_basis_combine(double* __restrict__ x, const double* const* __restrict__ V,
                                                            const double* __restrict__ y, int k, int zero_first,
                                                            int64_t n) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            const int64_t j = i << 2;
            d4 acc = {0, 0, 0, 0};
            if (!zero_first) acc = ld4(x + j);
            int c = 0;
            for (; c + 1 < k; c += 2) {  // two basis vectors in flight
                const d4 v0 = ld4_stream(V[c] + j), v1 = ld4_stream(V[c + 1] + j);
                const double y0 = y[c], y1 = y[c + 1];
                acc.x = fma(y0, v0.x, acc.x); acc.y = fma(y0, v0.y, acc.y); acc.z = fma(y0, v0.z, acc.z); acc.w = fma(y0, v0.w, acc.w);
                acc.x = fma(y1, v1.x, acc.x); acc.y = fma(y1, v1.y, acc.y); acc.z = fma(y1, v1.z, acc.z); acc.w = fma(y1, v1.w, acc.w);
            }
            if (c < k) {
                const d4 v0 = ld4_stream(V[c] + j);
                const double y0 = y[c];
                acc.x = fma(y0, v0.x, acc.x); acc.y = fma(y0, v0.y, acc.y); acc.z = fma(y0, v0.z, acc.z); acc.w = fma(y0, v0.w, acc.w);
            }
            st4(x + j, acc);
        }
        const int64_t j = (n4 << 2) + tid;
        if (j < n) {
            double acc = zero_first ? 0.0 : x[j];
            for (int c = 0; c < k; ++c) acc = fma(y[c], V[c][j], acc);
            x[j] = acc;
        }
    } else {
        for (int64_t j = tid; j < n; j += nth) {
            double acc = zero_first ? 0.0 : x[j];
            for (int c = 0; c < k; ++c) acc = fma(y[c], V[c][j], acc);
            x[j] = acc;
        }
    }
}

int launch_basis_combine(Ctx* ctx, int64_t n, double* x, const double* const* V_dev, const double* y_dev, int k,
                         int zero_x_first) {
    if (n <= 0) return AK_OK;
    const int blocks = stream_blocks(ctx, n, 4);
    ProfScope prof(ctx, PK_COMBINE);
    // basis vectors come from the workspace arena (256-byte aligned); x may be caller memory
    if (aligned32(x))
        k_basis_combine<true><<<blocks, kThreads, 0, ctx->stream>>>(x, V_dev, y_dev, k, zero_x_first, n);
    else
        k_basis_combine<false><<<blocks, kThreads, 0, ctx->stream>>>(x, V_dev, y_dev, k, zero_x_first, n);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

}  // namespace ak

// ---------------------------------------------------------------------------------------
// C ABI: Krylov.k* hooks (examples/halovector.jl:51-147)
// ---------------------------------------------------------------------------------------
using namespace ak;

static int scalar_to_host(Ctx* c, const double* dev, double* out_host) {
    AK_CUDA(cudaMemcpyAsync(c->hscal, dev, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    AK_CUDA(cudaStreamSynchronize(c->stream));
    *out_host = c->hscal[0];
    return AK_OK;
}

AK_API int ak_dot(ak_ctx* ctx, int64_t n, const double* x, const double* y, double* out_host) {
    AK_REQUIRE(ctx && out_host && n >= 0, "ak_dot: bad argument");
    AK_TRY(launch_dot(&ctx->c, n, x, y, ctx->c.dscal));
    return scalar_to_host(&ctx->c, ctx->c.dscal, out_host);
}
AK_API int ak_nrm2(ak_ctx* ctx, int64_t n, const double* x, double* out_host) {
    AK_REQUIRE(ctx && out_host && n >= 0, "ak_nrm2: bad argument");
    AK_TRY(launch_sumsq(&ctx->c, n, x, ctx->c.dscal));
    double ss = 0.0;
    AK_TRY(scalar_to_host(&ctx->c, ctx->c.dscal, &ss));
    *out_host = sqrt(ss);
    return AK_OK;
}
AK_API int ak_scal(ak_ctx* ctx, int64_t n, double s, double* x) {
    AK_REQUIRE(ctx && n >= 0, "ak_scal: bad argument");
    return launch_scal(&ctx->c, n, s, x);
}
AK_API int ak_axpy(ak_ctx* ctx, int64_t n, double s, const double* x, double* y) {
    AK_REQUIRE(ctx && n >= 0, "ak_axpy: bad argument");
    return launch_axpy(&ctx->c, n, s, x, y);
}
AK_API int ak_axpby(ak_ctx* ctx, int64_t n, double s, const double* x, double t, double* y) {
    AK_REQUIRE(ctx && n >= 0, "ak_axpby: bad argument");
    return launch_axpby(&ctx->c, n, s, x, t, y);
}
AK_API int ak_copy(ak_ctx* ctx, int64_t n, double* y, const double* x) {
    AK_REQUIRE(ctx && n >= 0, "ak_copy: bad argument");
    return launch_copy(&ctx->c, n, y, x);
}
AK_API int ak_fill(ak_ctx* ctx, int64_t n, double* x, double val) {
    AK_REQUIRE(ctx && n >= 0, "ak_fill: bad argument");
    return launch_fill(&ctx->c, n, x, val);
}
AK_API int ak_ref(ak_ctx* ctx, int64_t n, double* x, double* y, double c, double s) {
    AK_REQUIRE(ctx && n >= 0, "ak_ref: bad argument");
    return launch_ref(&ctx->c, n, x, y, c, s);
}
AK_API int ak_divcopy(ak_ctx* ctx, int64_t n, double* y, const double* x, double s) {
    AK_REQUIRE(ctx && n >= 0, "ak_divcopy: bad argument");
    return launch_divcopy(&ctx->c, n, y, x, s);
}
