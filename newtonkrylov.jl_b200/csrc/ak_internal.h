// Internal declarations shared by the translation units of libariadne_b200.so.
// Nothing here is part of the C ABI (that is include/ariadne_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "../../include/ariadne_b200.h"

#define AK_API extern "C" __attribute__((visibility("default")))

namespace ak {

void set_error(const char* fmt, ...);

#define AK_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ak::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return AK_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define AK_TRY(expr)              \
    do {                          \
        int _rc = (expr);         \
        if (_rc < 0) return _rc;  \
    } while (0)

#define AK_REQUIRE(cond, msg)                                          \
    do {                                                               \
        if (!(cond)) {                                                 \
            ak::set_error("%s:%d: %s", __FILE__, __LINE__, msg);       \
            return AK_ERR_ARG;                                         \
        }                                                              \
    } while (0)

// ---- NVLink peer memory (CUDA IPC) for the fused compute + collective kernels -----------------
// One cudaMalloc'ed block per rank, mapped into every other rank of the node:
//   mail : kMailSlots x nranks records {kBlkSums sums, tag, pad}  — partial sums written by peer kernels
//   halo : 2 parities x {lo, hi} x halo_cap doubles       — boundary rows pushed by the neighbours' final Gram-Schmidt pass
//   ghost: 2 parities x {lo, hi} x halo_cap doubles       — ghost values pushed by the stand-alone exchange kernel
constexpr int kMaxPeers = 8;
constexpr int kMailSlots = 8;
constexpr int kBlkMax = 8;    // most Gram-Schmidt steps per sweep over w (largest block of the blocked sweep)
// sums of one pass: up to kBlkMax projections <S_b, w>; final pass: ||w||^2 and the Gram entries <S_a, w_new> of the
// vector it finishes with the vectors of its own block (cached: they do not change in later iterations)
constexpr int kBlkSums = kBlkMax + 1;
constexpr int kMailRec = kBlkSums + 2;  // doubles per mailbox record: kBlkSums sums, tag, pad
struct P2PDev {           // passed by value to kernels
    int nranks, rank;
    double* mail_local;                 // this rank's mailbox
    double* mail_peer[kMaxPeers];       // everybody's mailbox (index = rank; [rank] == mail_local)
    int* err;                           // mapped host flag: set when a spin-wait times out
    long long spin_cycles;              // time-out
};
struct P2PHalo {          // boundary-row push of the final Gram-Schmidt pass
    double* down_hi;      // neighbour (rank-1)'s "hi" ghost row for this parity, or nullptr
    double* up_lo;        // neighbour (rank+1)'s "lo" ghost row for this parity, or nullptr
    int64_t nx;
};

// Upper bound on blocks of any reduction kernel (size of the partials buffer).
constexpr int kMaxPartials = 1 << 16;
constexpr int kNumSMsDefault = 148;

struct Comm;  // NCCL state (context.cu)

// kernel classes for the optional in-stream profiler (ak_profile_*)
enum ProfClass {
    PK_MGS_AXPY_DOT = 0,   // w -= h v_i ; h' = <v_{i+1}, w>      32n bytes
    PK_MGS_AXPY_NRM = 1,   // w -= h v_k ; ||w||^2                24n
    PK_MGS_AXPY = 2,       // w -= h v_i                          24n
    PK_DOT = 3,            // <x, y>                              16n
    PK_SUMSQ = 4,          // ||x||^2                              8n
    PK_JVP = 5,            // J(u) v (+ fused divcopy / dot)
    PK_RESIDUAL = 6,       // F(u) (+ fused norm)
    PK_ELEMENTWISE = 7,    // scal/axpy/axpby/copy/fill/divcopy/ref
    PK_COMBINE = 8,        // x = sum y_i V_i
    PK_SCALAR = 9,         // one-thread Givens / control kernels
    PK_MGS_PAIR = 10,      // full blocked pass: w -= sum_b h_b v_b ; projections on the next block   8n(2+2R): 48n (R=2), 80n (R=4)
    PK_MGS_PAIR_EDGE = 11, // first / ragged passes of the blocked sweep
    PK_MGS_BLOCK_FINAL = 12, // final pass of the blocked sweep: w -= sum_b c_b S_b ; ||w||^2 and the new Gram entries
    PK_SWEEP = 13,           // one-sweep GMRES iteration: update + norms + tangent + all projections   8n(k + 4)
    PK_NUM = 14
};

struct Ctx {
    int device = 0;
    int num_sms = kNumSMsDefault;
    cudaStream_t stream = nullptr;
    // the library's own stream-ordered memory pool (never trimmed): a Krylov workspace (22+ vectors of n doubles) is
    // re-created by every newton_krylov! call like in the reference (src/Ariadne.jl:317-318) but costs no
    // cudaMalloc/cudaFree after the first.  Private, so that the release threshold of the process-wide default pool
    // (shared with torch / CUDA.jl in the same process) is left alone.
    cudaMemPool_t pool = nullptr;
    // grid-wide reduction scratch (kernels on one stream never overlap)
    double* partials = nullptr;     // kMaxPartials doubles
    unsigned int* ticket = nullptr; // last-block-done counter, self-resetting
    // device scalar bank + pinned host mirror for synchronous scalar returns
    double* dscal = nullptr;        // 64 doubles
    double* hscal = nullptr;        // 64 doubles, pinned
    int64_t launches = 0;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    // optional profiler: CUDA-event pairs around every launch of the classes above
    bool prof_on = false;
    struct ProfRec { int cls; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    size_t prof_pool_used = 0;
    // multi-GPU
    Comm* comm = nullptr;
    int rank = 0, nranks = 1;
    double* halo_lo = nullptr;      // nx doubles: row gy0-1 (from rank-1)
    double* halo_hi = nullptr;      // nx doubles: row gy0+ny (from rank+1)
    double* halo_send = nullptr;
    int64_t halo_cap = 0;
    // peer memory (ak_comm_enable_p2p)
    bool p2p_on = false;
    double* p2p_block = nullptr;                 // local IPC-exported block
    void* p2p_peer_block[kMaxPeers] = {};         // mapped blocks of the other ranks
    int64_t p2p_halo_cap = 0;                    // doubles per ghost row
    int* p2p_err = nullptr;                      // pinned, mapped
    uint64_t p2p_seq = 0;                        // collective sequence number (identical on all ranks)
    P2PDev p2p_dev() const;
    uint64_t p2p_xchg = 0;                       // count of stand-alone ghost exchanges (parity of their slots)
    double* p2p_halo_local(int parity, int hi) const;              // this rank's ghost rows
    double* p2p_halo_of(int peer, int parity, int hi) const;       // a neighbour's ghost rows (mapped)
    // second set of ghost slots, used by the stand-alone exchanges (residuals, tangents outside the blocked sweep)
    double* p2p_ghost_local(int parity, int hi) const;
    double* p2p_ghost_of(int peer, int parity, int hi) const;
    // one-sweep GMRES (sweep.cu): mailboxes for its 2k + 2 sums and one pair of ghost rows per basis vector / W buffer
    double* p2p_swmail_of(int peer) const;
    double* p2p_swghost_local(int slot, int hi) const;
    double* p2p_swghost_of(int peer, int slot, int hi) const;
};

// stream-ordered allocation from the context's private pool (freed with cudaFreeAsync on the context's stream)
inline cudaError_t pool_alloc(Ctx* c, void** out, size_t bytes) {
    return cudaMallocFromPoolAsync(out, bytes, c->pool, c->stream);
}
// make the context's device current for the calling thread (a process may hold contexts on several devices)
inline cudaError_t bind_device(const Ctx* c) {
    int cur = -1;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return e;
    return cur == c->device ? cudaSuccess : cudaSetDevice(c->device);
}
#define AK_ENTER(ctxptr)                                           \
    do {                                                           \
        AK_REQUIRE((ctxptr) != nullptr, "NULL context");           \
        AK_CUDA(ak::bind_device(&(ctxptr)->c));                    \
    } while (0)

// RAII: brackets one kernel launch with events when the profiler is on (context.cu)
struct ProfScope {
    Ctx* c;
    int idx;
    ProfScope(Ctx* ctx, int cls);
    ~ProfScope();
};

// --- comm (context.cu) --------------------------------------------------------
// In-stream sum all-reduce of `count` device doubles (no-op on one rank).
int allreduce_sum(Ctx* ctx, double* dev, int count);
// Make a local success/failure verdict collective: *rc becomes an error on every rank when it is one on any rank
// (host-synchronising; used where ranks must take the same control-flow branch, e.g. basis growth).  No-op on one rank.
int collective_verdict(Ctx* ctx, int* rc);
// Fill ctx->halo_lo / halo_hi with the neighbours' boundary rows of `v` (nx*ny slab).
// Returns pointers to use as ghost rows (nullptr => implicit zero).
int exchange_halo_rows(Ctx* ctx, const double* v, int64_t nx, int64_t ny, int32_t bc,
                       const double** lo, const double** hi);
// 1-D segments: receive the last `nlo` values of the left rank (-> *lo) and the first `nhi` values of the right
// rank (-> *hi); nullptr at a physical (non-periodic) end.
int exchange_halo_1d(Ctx* ctx, const double* v, int64_t n, int nlo, int nhi, bool periodic, const double** lo,
                     const double** hi);

// --- user.cu: caller-supplied residual / tangent callbacks, generic finite-difference JVP -------
int user_residual(Ctx* ctx, const ak_problem* p, double* u, double* res);
int user_jvp(Ctx* ctx, const ak_problem* p, const double* u, double* v, double* out);
int launch_jvp_fd(Ctx* ctx, const ak_problem* p, const double* u, double* v, double* out);

// --- precond.cu: Jacobi, tridiagonal LU, caller-supplied apply ---------------------------------
int precond_apply(Ctx* ctx, const ak_problem* p, const double* u, int32_t kind, ak_precond_apply_fn fn, void* user,
                  const double* x, double* y);

// --- blas1.cu -------------------------------------------------------------------
int launch_dot(Ctx* ctx, int64_t n, const double* x, const double* y, double* out_dev);
int launch_sumsq(Ctx* ctx, int64_t n, const double* x, double* out_dev);
int launch_scal(Ctx* ctx, int64_t n, double s, double* x);
int launch_axpy(Ctx* ctx, int64_t n, double s, const double* x, double* y);
// y += (sign * *s_dev) * x   — scalar read from device memory
int launch_axpy_dev(Ctx* ctx, int64_t n, const double* s_dev, double sign, const double* x, double* y);
int launch_axpby(Ctx* ctx, int64_t n, double s, const double* x, double t, double* y);
int launch_copy(Ctx* ctx, int64_t n, double* y, const double* x);
int launch_fill(Ctx* ctx, int64_t n, double* x, double v);
int launch_ref(Ctx* ctx, int64_t n, double* x, double* y, double c, double s);
int launch_divcopy(Ctx* ctx, int64_t n, double* y, const double* x, double s);
// y <- x / (*s_dev)
int launch_divcopy_dev(Ctx* ctx, int64_t n, double* y, const double* x, const double* s_dev, const int* stop_flag);
// Fused modified-Gram-Schmidt step (arnoldi kernels, blas1.cu):
//   if vi:    w <- w - (*h_in) * vi
//   if vnext: *out = <vnext, w_new>     else if want_sumsq: *out = <w_new, w_new>
// `stop_flag` (device int, may be null): kernel is a no-op when *stop_flag != 0.
int launch_mgs_step(Ctx* ctx, int64_t n, double* w, const double* vi, const double* h_in,
                    const double* vnext, int want_sumsq, double* out_dev, const int* stop_flag);
// Blocked Gram-Schmidt pass (see blas1.cu): subtracts up to kBlkMax basis vectors `va[0..nax)` with the
// coefficients recovered from the raw projections `tin`, the cached Gram entries `gram_in` of that block and the
// scales `rho_in`, and projects the result on up to kBlkMax vectors `ya[0..ny)` (raw sums -> out), or (want_sumsq)
// returns ||w||^2 in out[0] and <va[a], w_new> in out[1 + a].
// `pc` (may be null) routes the reduction through the peers' mailboxes (NVLink) instead of NCCL and lets the
// final pass push its boundary rows to the neighbours.
struct BlockComm {
    unsigned long long seq_in = 0, seq_out = 0;
    double* tin_store = nullptr;
    P2PHalo halo = {nullptr, nullptr, 0};
};
int launch_mgs_block(Ctx* ctx, int64_t n, double* w, const double* const* va, int nax, const double* tin,
                     const double* gram_in, const double* rho_in, const double* const* ya, int ny, int want_sumsq,
                     double* out, const int* stop, const BlockComm* pc);
// x <- [x +] sum_{i<k} y[i] V[i]  (sum formed from zero in the sequential axpy order of gmres!, then stored or
// added to x); k = *k_dev when k_dev != null (device-resident pass length), else k_host
int launch_basis_combine(Ctx* ctx, int64_t n, double* x, const double* const* V_dev, const double* y_dev,
                         const int* k_dev, int k_host, int accumulate);

// --- stencil.cu -----------------------------------------------------------------
// res <- F(u); if sumsq_dev != null also *sumsq_dev = ||res||^2 (global when comm is set)
int launch_residual(Ctx* ctx, const ak_problem* p, double* u, double* res, double* sumsq_dev);
// out <- J(u) v; if dot_with != null also *dot_dev = <dot_with, out>
// if scale_src != null: v is first formed as scale_src / (*denom_dev) and written to v (fused divcopy)
struct JvpFusion {
    const double* scale_src = nullptr;  // w_prev: v <- scale_src / denom, written to v
    const double* denom_dev = nullptr;  // device scalar
    const double* inv_denom_dev = nullptr;  // (raw) 1 / *denom_dev, precomputed by the scalar kernels: saves a division per thread
    // un-normalised basis: scale_src stays the stored basis vector and out = J(scale_src) / denom (J is linear; the
    // stencil kernels scale at the store).  `v` is then scratch (n doubles) for the problem kinds that go through a
    // normalised copy (2x2 system, Midpoint, caller-supplied tangents, finite differences), unused otherwise.
    bool raw = false;
    const double* dot_with = nullptr;   // V[0]
    double* dot_dev = nullptr;
    // restart residual of gmres! (w <- b - A x: mul!(w, A, x); kaxpby!(n, one, b, -one, w)) in the same pass:
    // out = rhs_minus - J v, and (sumsq_dev != null, exclusive with dot_with) *sumsq_dev = ||out||^2
    const double* rhs_minus = nullptr;
    double* sumsq_dev = nullptr;
    // first projection pass of the blocked Gram-Schmidt sweep folded into the tangent kernel (2-D analytic tangents):
    // proj_out[b] = <proj[b], out>, b < nproj; with proj_comm (peer memory) the sums go to the ranks' mailboxes
    const double* const* proj = nullptr;
    int nproj = 0;
    double* proj_out = nullptr;
    const struct BlockComm* proj_comm = nullptr;
    const int* stop_flag = nullptr;
    // ghost rows already delivered by the neighbours through peer memory (skips the NCCL exchange)
    bool halo_given = false;
    const double* halo_lo = nullptr;
    const double* halo_hi = nullptr;
};
int launch_jvp(Ctx* ctx, const ak_problem* p, const double* u, double* v, double* out, const JvpFusion* f);
int launch_jvp_transpose(Ctx* ctx, const ak_problem* p, const double* u, double* v, double* out);
// Out[:, c] = J(u) V[:, c] for the Bratu problems with lambda e^u read once for all columns (single GPU)
int launch_jvp_batched_bratu(Ctx* ctx, const ak_problem* p, const double* u, const double* V, int64_t ldv, double* Out,
                             int64_t ldo, int32_t ncols);

// --- krylov.cu ------------------------------------------------------------------
}  // namespace ak

struct ak_ctx {
    ak::Ctx c;
};
