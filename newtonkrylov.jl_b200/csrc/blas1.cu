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
#include <atomic>

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
// Blocked modified Gram-Schmidt pass (fuse levels PAIR = blocks of 2, BLOCK4 = blocks of kBlkMax): several
// Gram-Schmidt steps per sweep over w.
//   NAX axpys      : w <- w - sum_{b<NAX} c_b S_b, (h, c) = block_coefficients(projections `tin` of that block,
//                    its cached Gram entries `gram_in`, its scales `rho_in`)
//   NRED 1..kBlkMax: projections on the next block, t[b] = <S'_b, w_new>.  With the cached Gram entries
//                    h_b = <v_b,w> - sum_{a<b} h_a <v_b,v_a> is algebraically the modified Gram-Schmidt coefficient
//                    <v_b, w_new - sum_{a<b} h_a v_a>, obtained without another sweep over w.
//   NRED = final   : out[0] = <w_new, w_new>, out[1 + a] = <S_a, w_new> for the NAX vectors just subtracted: w_new is the
//                    stored basis vector this iteration produces, and these are its Gram entries with the earlier
//                    vectors of its own block (rounding-level numbers: the loss of orthogonality that the
//                    coefficients of later iterations correct for).  They come for free: S_a is in registers.
// Algorithmic bytes per launch: 8n (2 [w in/out] + NAX + number of projection vectors): 48n for blocks of 2 (24n per
// Gram-Schmidt step), 80n for blocks of 4 (20n per step), instead of 32n per step for axpy_i + dot_{i+1}.
// -----------------------------------------------------------------------------------
struct BlkPtrs {
    const double* va[kBlkMax];  // stored vectors to subtract
    const double* ya[kBlkMax];  // stored vectors to project on
};
struct BlockP2P {     // fused collective of the blocked pass (all zero / null when not used)
    P2PDev pd;
    unsigned long long seq_in;   // record to wait for (reduced sums of the block being subtracted), 0 = read `tin`
    unsigned long long seq_out;  // record to post (this pass's sums), 0 = store to `out`
    double* tin_store;           // where block 0 leaves the reduced incoming sums for the Givens kernel
    P2PHalo halo;                // boundary-row push (final pass only)
};
constexpr int kRedFinal = kBlkMax + 1;

// resident blocks per SM the register budget is sized for: wide passes hold up to 16 vectors of 4 doubles per thread
constexpr int mgs_block_min_blocks(int nax, int nred) {
    return (nred == kBlkMax + 1) ? (nax > 4 ? 1 : 2) : ((nax + nred > 8 || nred > 6) ? 1 : 2);
}

template <int NAX, int NRED, bool P2P>
__global__ void __launch_bounds__(kThreads, mgs_block_min_blocks(NAX, NRED)) k_mgs_block(double* __restrict__ w, const BlkPtrs bp,
                                                           const double* __restrict__ tin,
                                                           const double* __restrict__ gram_in,
                                                           const double* __restrict__ rho_in, double* __restrict__ out,
                                                           double* __restrict__ partials, unsigned int* ticket,
                                                           int64_t n, const int* __restrict__ stop, const int vec,
                                                           const BlockP2P pp) {
    constexpr bool FINAL = (NRED == kRedFinal);
    constexpr int NY = FINAL ? 0 : NRED;
    constexpr int NS = FINAL ? 1 + NAX : NRED;
    __shared__ double sh[32];
    __shared__ double shm[kBlkSums * kMaxPeers];
    __shared__ double s_t[kBlkSums], s_g[kBlkMax * kBlkMax], s_h[kBlkMax], s_c[kBlkMax];
    if (stop != nullptr && *stop != 0) return;
    double h[kBlkMax];  // negated multipliers of the stored vectors of the block being subtracted
#pragma unroll
    for (int b = 0; b < kBlkMax; ++b) h[b] = 0.0;
    if (NAX >= 1) {
        if (P2P && pp.seq_in != 0) {
            mail_wait_sum(pp.pd, pp.seq_in, s_t, NAX, shm, const_cast<int*>(stop));
            if (blockIdx.x == 0 && threadIdx.x < NAX && pp.tin_store != nullptr) pp.tin_store[threadIdx.x] = s_t[threadIdx.x];
        } else {
            if (threadIdx.x < NAX) s_t[threadIdx.x] = tin[threadIdx.x];
            __syncthreads();
        }
        // one warp turns the raw projections into the Gram-Schmidt multipliers, the block reads them from shared memory
        if (threadIdx.x < 32) block_coefficients_warp(threadIdx.x, s_t, gram_in, rho_in, NAX, s_g, s_h, s_c);
        __syncthreads();
#pragma unroll
        for (int b = 0; b < NAX; ++b) h[b] = -s_c[b];
    }
    double sA[NS], sB[NS];  // two accumulator sets (even / odd elements of a 256-bit word)
#pragma unroll
    for (int c = 0; c < NS; ++c) sA[c] = sB[c] = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const bool push = P2P && FINAL && (pp.halo.down_hi != nullptr || pp.halo.up_lo != nullptr);
    const int64_t hnx = pp.halo.nx;
    auto body = [&](double& wj, const double (&x)[kBlkMax], const double (&y)[kBlkMax], double (&s)[NS]) {
#pragma unroll
        for (int b = 0; b < NAX; ++b) wj = fma(h[b], x[b], wj);  // same order as successive kaxpy!
        if (FINAL) {
            s[0] = fma(wj, wj, s[0]);
#pragma unroll
            for (int a = 0; a < NAX; ++a) s[1 + a] = fma(x[a], wj, s[1 + a]);
        } else {
#pragma unroll
            for (int b = 0; b < NY; ++b) s[b] = fma(y[b], wj, s[b]);
        }
    };
    auto scalar_elem = [&](int64_t j) {
        double wj = w[j];
        double x[kBlkMax], y[kBlkMax];
#pragma unroll
        for (int b = 0; b < kBlkMax; ++b) x[b] = y[b] = 0.0;
#pragma unroll
        for (int b = 0; b < NAX; ++b) x[b] = bp.va[b][j];
#pragma unroll
        for (int b = 0; b < NY; ++b) y[b] = bp.ya[b][j];
        body(wj, x, y, sA);
        if (NAX > 0) w[j] = wj;
    };
    if (vec) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            const int64_t j = i << 2;
            d4 wv = (NAX > 0) ? ld4(w + j) : ld4_stream(w + j);
            d4 xv[kBlkMax], yv[kBlkMax];
#pragma unroll
            for (int b = 0; b < kBlkMax; ++b) {
                xv[b] = d4{0, 0, 0, 0};
                yv[b] = d4{0, 0, 0, 0};
            }
#pragma unroll
            for (int b = 0; b < NAX; ++b) xv[b] = ld4_stream(bp.va[b] + j);
#pragma unroll
            for (int b = 0; b < NY; ++b) yv[b] = ld4_stream(bp.ya[b] + j);
            {
                double x0[kBlkMax], x1[kBlkMax], x2[kBlkMax], x3[kBlkMax], y0[kBlkMax], y1[kBlkMax], y2[kBlkMax], y3[kBlkMax];
#pragma unroll
                for (int b = 0; b < kBlkMax; ++b) {
                    x0[b] = xv[b].x; x1[b] = xv[b].y; x2[b] = xv[b].z; x3[b] = xv[b].w;
                    y0[b] = yv[b].x; y1[b] = yv[b].y; y2[b] = yv[b].z; y3[b] = yv[b].w;
                }
                body(wv.x, x0, y0, sA);
                body(wv.y, x1, y1, sB);
                body(wv.z, x2, y2, sA);
                body(wv.w, x3, y3, sB);
            }
            if (NAX > 0) st4(w + j, wv);
            if (push) {  // boundary rows of the finished w go straight into the neighbours' ghost rows (NVLink)
                if (pp.halo.down_hi != nullptr && j < hnx) st4(pp.halo.down_hi + j, wv);
                if (pp.halo.up_lo != nullptr && j >= n - hnx) st4(pp.halo.up_lo + (j - (n - hnx)), wv);
            }
        }
        const int64_t j = (n4 << 2) + tid;
        if (j < n) scalar_elem(j);
    } else {
        for (int64_t j = tid; j < n; j += nth) scalar_elem(j);
    }
    double r[NS];
#pragma unroll
    for (int c = 0; c < NS; ++c) r[c] = sA[c] + sB[c];
    double tot[NS];
    if (grid_reduce_n<NS>(r, partials, ticket, blockIdx.x, gridDim.x, sh, tot, push)) {
        if (P2P && pp.seq_out != 0) {
            mail_post(pp.pd, pp.seq_out, tot, NS);
        } else {
#pragma unroll
            for (int c = 0; c < NS; ++c) out[c] = tot[c];
        }
    }
}

template <int NAX, int NRED>
static int launch_mgs_block_t(Ctx* ctx, int64_t n, double* w, const BlkPtrs& bp, const double* tin,
                              const double* gram_in, const double* rho_in, double* out, const int* stop, bool vec,
                              bool p2p, const BlockP2P& pp, int cls) {
    // resident blocks per SM of this instantiation (a property of the sm_100a binary: the same on every B200 of
    // the box, so one process-wide cache per instantiation is enough; atomic because contexts may live on threads)
    static std::atomic<int> occ_cache[2] = {{0}, {0}};
    int occ[2] = {occ_cache[0].load(std::memory_order_relaxed), occ_cache[1].load(std::memory_order_relaxed)};
    if (occ[p2p] == 0) {
        int nb = 0;
        cudaError_t e = p2p ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_mgs_block<NAX, NRED, true>, kThreads, 0)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_mgs_block<NAX, NRED, false>, kThreads, 0);
        AK_CUDA(e);
        occ[p2p] = nb < 1 ? 1 : (nb > 4 ? 4 : nb);
        occ_cache[p2p].store(occ[p2p], std::memory_order_relaxed);
    }
    int64_t need = (n + (int64_t)kThreads * 4 - 1) / ((int64_t)kThreads * 4);
    const int64_t cap = (int64_t)ctx->num_sms * occ[p2p];  // one wave of resident blocks, grid-stride
    const int blocks = (int)(need < 1 ? 1 : (need < cap ? need : cap));
    ProfScope prof(ctx, cls);
    if (p2p)
        k_mgs_block<NAX, NRED, true><<<blocks, kThreads, 0, ctx->stream>>>(w, bp, tin, gram_in, rho_in, out,
                                                                           ctx->partials, ctx->ticket, n, stop,
                                                                           vec ? 1 : 0, pp);
    else
        k_mgs_block<NAX, NRED, false><<<blocks, kThreads, 0, ctx->stream>>>(w, bp, tin, gram_in, rho_in, out,
                                                                            ctx->partials, ctx->ticket, n, stop,
                                                                            vec ? 1 : 0, pp);
    return AK_OK;
}

// compile-time enumeration of the (NAX, NRED) shapes the sweep uses
template <int A, int R>
static int dispatch_block(int nax, int nred, Ctx* ctx, int64_t n, double* w, const BlkPtrs& bp, const double* tin,
                          const double* gram_in, const double* rho_in, double* out, const int* stop, bool vec, bool p2p,
                          const BlockP2P& pp, int cls) {
    if (nax == A && nred == R) {
        // a pass either projects (first pass: nothing to subtract; later passes: any block) or is the final one
        // (projection passes subtract nothing, a full pair or a full block; the final pass subtracts 1..kBlkMax)
        // a second (re-orthogonalisation) sweep also subtracts a ragged last block while it projects on block 0 again:
        // (A, A) when the whole basis is one block, (A <= R, R = 2, 4, kBlkMax) otherwise
        if constexpr (R == kRedFinal ? (A >= 1)
                                     : (A == 0 || A == 2 || A == 4 || A == kBlkMax || R == A ||
                                        ((R == 2 || R == 4 || R == kBlkMax) && A <= R)))
            return launch_mgs_block_t<A, R>(ctx, n, w, bp, tin, gram_in, rho_in, out, stop, vec, p2p, pp, cls);
    }
    if constexpr (R < kRedFinal) {
        return dispatch_block<A, R + 1>(nax, nred, ctx, n, w, bp, tin, gram_in, rho_in, out, stop, vec, p2p, pp, cls);
    } else if constexpr (A < kBlkMax) {
        return dispatch_block<A + 1, 1>(nax, nred, ctx, n, w, bp, tin, gram_in, rho_in, out, stop, vec, p2p, pp, cls);
    } else {
        set_error("launch_mgs_block: no kernel for nax = %d, nred = %d", nax, nred);
        return AK_ERR_ARG;
    }
}

// va[0..nax): stored vectors to subtract, tin: raw projections on that block, gram_in: its cached Gram entries, rho_in
// (may be null): its scales; ya[0..ny): vectors to project on; want_sumsq: final pass instead (||w_new||^2 and the Gram
// entries of w_new with va).  out receives ny doubles, or 1 + nax.
// With `pc` (multi-GPU, peer memory enabled) the sums travel through the peers' mailboxes instead of NCCL.
int launch_mgs_block(Ctx* ctx, int64_t n, double* w, const double* const* va, int nax, const double* tin,
                     const double* gram_in, const double* rho_in, const double* const* ya, int ny, int want_sumsq,
                     double* out, const int* stop, const BlockComm* pc) {
    if (n <= 0) return AK_OK;
    if (nax < 0 || nax > kBlkMax || ny < 0 || ny > kBlkMax || (ny == 0 && !want_sumsq) || (ny > 0 && want_sumsq)) {
        set_error("launch_mgs_block: bad block shape (nax = %d, ny = %d, sumsq = %d)", nax, ny, want_sumsq);
        return AK_ERR_ARG;
    }
    const int nred = want_sumsq ? kRedFinal : ny;
    BlkPtrs bp{};
    bool vec = aligned32(w);
    for (int b = 0; b < nax; ++b) { bp.va[b] = va[b]; vec = vec && aligned32(va[b]); }
    for (int b = 0; b < ny; ++b) { bp.ya[b] = ya[b]; vec = vec && aligned32(ya[b]); }
    const int cls = (nax >= 2 && nax == ny) ? PK_MGS_PAIR
                                            : ((nax > 0 && want_sumsq) ? PK_MGS_BLOCK_FINAL : PK_MGS_PAIR_EDGE);
    const bool p2p = pc != nullptr && ctx->p2p_on && ctx->nranks > 1;
    BlockP2P pp{};
    if (p2p) {
        pp.pd = ctx->p2p_dev();
        pp.seq_in = pc->seq_in;
        pp.seq_out = pc->seq_out;
        pp.tin_store = pc->tin_store;
        pp.halo = pc->halo;
        // the vector path pushes whole 256-bit words: rows must be 32-byte multiples
        if (pp.halo.nx % 4 != 0 || n % 4 != 0 || !vec) pp.halo.down_hi = pp.halo.up_lo = nullptr;
    }
    AK_TRY((dispatch_block<0, 1>(nax, nred, ctx, n, w, bp, tin, gram_in, rho_in, out, stop, vec, p2p, pp, cls)));
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    if (!p2p) AK_TRY(allreduce_sum(ctx, out, want_sumsq ? 1 + nax : ny));
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
// xr = sum_{i<k} y[i] V[i] in the sequential axpy order of gmres! step 10 (for i = 1:k kaxpy!(n, y[i], V[i], xr)),
// formed in registers from zero; then  x <- xr  (accumulate = 0)  or  x <- x + xr  (accumulate = 1: the restart
// update `kaxpy!(n, one, xr, x)` without materialising xr).  k and y live in device memory (k_dev may be null: k_host),
// so the launch needs no host knowledge of how far the pass got.
// Algorithmic bytes 8n(k+1), +8n when accumulating.
// -----------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_basis_combine(double* __restrict__ x, const double* const* __restrict__ V,
                                                            const double* __restrict__ y, const int* __restrict__ k_dev,
                                                            int k_host, int accumulate, int64_t n) {
    const int k = k_dev != nullptr ? *k_dev : k_host;
    if (k <= 0 && accumulate) return;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            const int64_t j = i << 2;
            d4 acc = {0, 0, 0, 0};
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
            if (accumulate) {
                const d4 xo = ld4(x + j);
                acc.x = xo.x + acc.x; acc.y = xo.y + acc.y; acc.z = xo.z + acc.z; acc.w = xo.w + acc.w;
            }
            st4(x + j, acc);
        }
        const int64_t j = (n4 << 2) + tid;
        if (j < n) {
            double acc = 0.0;
            for (int c = 0; c < k; ++c) acc = fma(y[c], V[c][j], acc);
            x[j] = accumulate ? x[j] + acc : acc;
        }
    } else {
        for (int64_t j = tid; j < n; j += nth) {
            double acc = 0.0;
            for (int c = 0; c < k; ++c) acc = fma(y[c], V[c][j], acc);
            x[j] = accumulate ? x[j] + acc : acc;
        }
    }
}

int launch_basis_combine(Ctx* ctx, int64_t n, double* x, const double* const* V_dev, const double* y_dev,
                         const int* k_dev, int k_host, int accumulate) {
    if (n <= 0) return AK_OK;
    const int blocks = stream_blocks(ctx, n, 4);
    ProfScope prof(ctx, PK_COMBINE);
    // basis vectors come from the workspace arena (256-byte aligned); x may be caller memory
    if (aligned32(x))
        k_basis_combine<true><<<blocks, kThreads, 0, ctx->stream>>>(x, V_dev, y_dev, k_dev, k_host, accumulate, n);
    else
        k_basis_combine<false><<<blocks, kThreads, 0, ctx->stream>>>(x, V_dev, y_dev, k_dev, k_host, accumulate, n);
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
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && out_host && n >= 0, "ak_dot: bad argument");
    AK_TRY(launch_dot(&ctx->c, n, x, y, ctx->c.dscal));
    return scalar_to_host(&ctx->c, ctx->c.dscal, out_host);
}
AK_API int ak_nrm2(ak_ctx* ctx, int64_t n, const double* x, double* out_host) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && out_host && n >= 0, "ak_nrm2: bad argument");
    AK_TRY(launch_sumsq(&ctx->c, n, x, ctx->c.dscal));
    double ss = 0.0;
    AK_TRY(scalar_to_host(&ctx->c, ctx->c.dscal, &ss));
    *out_host = sqrt(ss);
    return AK_OK;
}
AK_API int ak_scal(ak_ctx* ctx, int64_t n, double s, double* x) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && n >= 0, "ak_scal: bad argument");
    return launch_scal(&ctx->c, n, s, x);
}
AK_API int ak_axpy(ak_ctx* ctx, int64_t n, double s, const double* x, double* y) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && n >= 0, "ak_axpy: bad argument");
    return launch_axpy(&ctx->c, n, s, x, y);
}
AK_API int ak_axpby(ak_ctx* ctx, int64_t n, double s, const double* x, double t, double* y) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && n >= 0, "ak_axpby: bad argument");
    return launch_axpby(&ctx->c, n, s, x, t, y);
}
AK_API int ak_copy(ak_ctx* ctx, int64_t n, double* y, const double* x) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && n >= 0, "ak_copy: bad argument");
    return launch_copy(&ctx->c, n, y, x);
}
AK_API int ak_fill(ak_ctx* ctx, int64_t n, double* x, double val) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && n >= 0, "ak_fill: bad argument");
    return launch_fill(&ctx->c, n, x, val);
}
AK_API int ak_ref(ak_ctx* ctx, int64_t n, double* x, double* y, double c, double s) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && n >= 0, "ak_ref: bad argument");
    return launch_ref(&ctx->c, n, x, y, c, s);
}
AK_API int ak_divcopy(ak_ctx* ctx, int64_t n, double* y, const double* x, double s) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && n >= 0, "ak_divcopy: bad argument");
    return launch_divcopy(&ctx->c, n, y, x, s);
}
