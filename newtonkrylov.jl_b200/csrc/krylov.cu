// Krylov workspace + GMRES / CG drivers.
//
// Replaces `krylov_workspace(algo, KrylovConstructor(res))` and
// `krylov_solve!(workspace, J, copy(res); kwargs...)` at src/Ariadne.jl:317-318,338-340,
// i.e. Krylov.jl's gmres!/cg! (not vendored in the reference; restated from the published
// algorithm) with the operator J = JacobianOperator(F!, res, u, p) of src/Ariadne.jl:34-57.
//
// B200-first structure: the whole Arnoldi iteration lives on the device.  Inner products land
// in a device column `hcol`, a one-thread kernel applies the Givens reflections, updates the
// least-squares right-hand side, evaluates the stopping tests and raises a device-side `stop`
// flag; every vector kernel starts by reading that flag.  The host therefore launches iteration
// k+1 before it knows the outcome of iteration k (it reads a pinned status record one iteration
// late), so the stream never drains inside a pass.  The Krylov basis stays resident in HBM.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ak_internal.h"
#include "common.cuh"
#include "sweep.h"

namespace ak {

struct KrylovCtl {  // device-resident control block
    double eps, rNorm, beta, Hbis, btol, atol, rtol;
    // CG scalars
    double gamma, gamma_next, pAp, pNorm2, alpha, neg_alpha, cg_beta;
    int stop, solved, breakdown, inconsistent, inner_iter, zerocurv;
};
struct KrylovStatus {  // pinned host ring, written by the scalar kernels
    double rNorm, Hbis, beta;
    int iter, stop, solved, breakdown, zerocurv, inconsistent;
};
constexpr int kStatusRing = 8;   // per-iteration records; slot kStatusRing is the pass prologue, kStatusRing + 1 the
                                 // record of the back-substitution (inconsistent flag, read once at the end of a solve)
// How many iterations the host launches ahead of the last verdict it has read.  Kernels of iterations after a stop
// are no-ops on the device, so depth only costs a few empty launches at the end of a pass; with short kernels
// (vectors that fit in L2) a depth of 1 leaves the stream empty while the host wakes up from the event.
constexpr int kSpecDepth = 3;
static_assert(kSpecDepth < kStatusRing, "the status ring must outlive the speculation window");
constexpr int kStatusSlots = kStatusRing + 2;

}  // namespace ak

struct ak_krylov {
    ak_ctx* owner = nullptr;
    ak::Ctx* ctx = nullptr;
    int32_t algo = AK_ALGO_GMRES;
    int64_t n = 0;
    int64_t n_global = 0;  // unknowns over all ranks: Krylov.jl's default itmax = 2n is a property of the whole system
    int32_t mem = 20;
    int64_t max_basis = 0;
    // vectors
    double* x = nullptr;
    double* w[2] = {nullptr, nullptr};
    double* dx = nullptr;  // xr when restart
    std::vector<double*> V;          // basis vectors
    std::vector<double*> Z;          // fgmres: z_k = N v_k
    double* pbuf = nullptr;          // right-preconditioned gmres: p = N v_k
    double* qbuf = nullptr;          // left-preconditioned gmres: A N v_k before M is applied
    ak_krylov* inner = nullptr;      // workspace of the inner GMRES used as preconditioner
    std::vector<double*> chunks;     // cudaMalloc'ed blocks backing V / misc vectors
    const double** V_dev = nullptr;  // device table of basis pointers
    int64_t V_dev_cap = 0;
    // scalars (device)
    int64_t kcap = 0;  // columns of R that fit
    double *R = nullptr, *c = nullptr, *s = nullptr, *z = nullptr, *hcol = nullptr, *hist = nullptr;
    bool last_raw = false;    // the last GMRES solve kept the basis un-normalised (stored V[i] = rho[i] v_i)
    double* ycoef = nullptr;  // solution of R y = z (device back-substitution), already divided by rho for the combine
    double* rho = nullptr;  // un-normalised basis (blocked sweeps): stored V[i] = rho[i] * v_i
    double* rinv = nullptr; // 1 / rho[i], formed once by the scalar kernels (the tangent kernels multiply their result by it)
    double* gram = nullptr; // blocked sweeps: <V[i], V[a]> for the earlier vectors a of V[i]'s own block (kBlkMax per i)
    // one-sweep Gram-Schmidt (sweep.cu): sums of the last sweep, column of H under construction, update multipliers,
    // scaled Gram matrix <v_j, v_a> of the cycle
    double *sw_sums = nullptr, *sw_h = nullptr, *sw_c = nullptr, *sw_gam = nullptr;
    int64_t hist_cap = 0;
    ak::KrylovCtl* ctl = nullptr;
    ak::KrylovStatus* status = nullptr;  // pinned
    cudaEvent_t ev[ak::kStatusSlots] = {};
    // cg
    double *r = nullptr, *p = nullptr, *Ap = nullptr;
};

namespace ak {

// ---------------------------------------------------------------------------------------
// scalar kernels (1 thread)
// ---------------------------------------------------------------------------------------
// start of a solve / of a restart pass: beta = ||r0||, z[0] = beta, stopping tolerance
__global__ void k_gmres_begin(KrylovCtl* ctl, const double* sumsq, double* z, double* hist, int first_pass,
                              double atol, double rtol, KrylovStatus* st, double* rho, double* rinv) {
    if (threadIdx.x != 0) return;
    const double beta = sqrt(*sumsq);
    if (first_pass) {
        ctl->beta = beta;
        ctl->rNorm = beta;
        ctl->eps = atol + rtol * beta;
        ctl->btol = pow(2.220446049250313e-16, 0.75);
        ctl->breakdown = 0;
        ctl->inconsistent = 0;
        if (hist) hist[0] = beta;
    }
    z[0] = beta;
    if (rho) { rho[0] = ctl->rNorm; rinv[0] = __ddiv_rn(1.0, ctl->rNorm); }  // un-normalised basis: V[0] stays r0, v_0 = r0 / rNorm
    ctl->solved = (ctl->rNorm <= ctl->eps) || (beta == 0.0);
    ctl->stop = ctl->solved;
    ctl->inner_iter = 0;
    st->rNorm = ctl->rNorm;
    st->beta = beta;
    st->Hbis = 0.0;
    st->iter = 0;
    st->stop = ctl->stop;
    st->solved = ctl->solved;
    st->breakdown = 0;
}

__device__ __forceinline__ double sgn(double x) { return (double)((x > 0.0) - (x < 0.0)); }

// Tail of one GMRES iteration (gmres! steps 6-8), shared by the Givens kernels of the pass-wise and of the one-sweep
// Gram-Schmidt: the previous reflections applied to the new column R[nr..nr+k), Krylov.jl's sym_givens on
// (R[nr+k-1], Hbis), the update of z and the stopping tests.  One warp; lane 0 does the (sequential) arithmetic, the
// warp prefetches 32 rotation coefficients and column entries at a time.  hh = ||q||^2 (valid on lane 0).
// Returns the stop verdict on lane 0.
__device__ __forceinline__ int givens_tail(int lane, KrylovCtl* ctl, int k, int64_t nr, double* R, double* c, double* s,
                                           double* z, double hh, double* rho_vec, double* rinv_vec, double* hist,
                                           int64_t hist_pos, int inner_limit, KrylovStatus* st, double* s_c, double* s_s,
                                           double* s_r) {
    double cur = 0.0;
    if (lane == 0) cur = R[nr];
    // previous reflections applied to the new column (gmres! step 6), 32 at a time: lane 0 carries the running entry,
    // the warp prefetches c_i, s_i and the untouched column entries r_{i+1}
    for (int base = 0; base + 1 < k; base += 32) {
        const int i = base + lane;
        if (i + 1 < k) {
            s_c[lane] = c[i];
            s_s[lane] = s[i];
            s_r[lane] = R[nr + i + 1];
        }
        __syncwarp();
        if (lane == 0) {
            const int cnt = (k - 1 - base) < 32 ? (k - 1 - base) : 32;
            for (int q = 0; q < cnt; ++q) {
                const double rn = s_r[q];
                const double Rt = s_c[q] * cur + s_s[q] * rn;
                const double nx = s_s[q] * cur - s_c[q] * rn;
                R[nr + base + q] = Rt;
                cur = nx;
            }
        }
        __syncwarp();
    }
    if (lane != 0) return 0;
    const double Hbis = sqrt(hh);
    if (rho_vec) { rho_vec[k] = Hbis; rinv_vec[k] = __ddiv_rn(1.0, Hbis); }  // the finished w of this iteration IS the stored basis vector k
    const double a = cur, b = Hbis;
    double ck, sk, rho;
    if (b == 0.0) {
        ck = (a == 0.0) ? 1.0 : sgn(a);
        sk = 0.0;
        rho = fabs(a);
    } else if (a == 0.0) {
        ck = 0.0;
        sk = sgn(b);
        rho = fabs(b);
    } else if (fabs(b) > fabs(a)) {
        const double t = a / b;
        sk = sgn(b) / sqrt(1.0 + t * t);
        ck = sk * t;
        rho = b / sk;
    } else {
        const double t = b / a;
        ck = sgn(a) / sqrt(1.0 + t * t);
        sk = ck * t;
        rho = a / ck;
    }
    c[k - 1] = ck;
    s[k - 1] = sk;
    R[nr + k - 1] = rho;
    const double zeta = sk * z[k - 1];
    z[k - 1] = ck * z[k - 1];
    const double rNorm = fabs(zeta);
    if (hist) hist[hist_pos] = rNorm;
    const int mach = (rNorm + 1.0 <= 1.0);
    const int solved = (rNorm <= ctl->eps) || mach;
    const int breakdown = (Hbis <= ctl->btol);
    const int tired = (k >= inner_limit);
    ctl->rNorm = rNorm;
    ctl->Hbis = Hbis;
    ctl->solved = solved;
    ctl->breakdown = breakdown;
    ctl->inner_iter = k;
    const int stop = solved || breakdown || tired;
    if (!stop) z[k] = zeta;
    ctl->stop = stop;
    st->rNorm = rNorm;
    st->Hbis = Hbis;
    st->iter = k;
    st->stop = stop;
    st->solved = solved;
    st->breakdown = breakdown;
    return stop;
}

// Krylov.jl sym_givens (real), then the update of iteration k (1-based) — gmres! steps 6-8.
// One warp.  Lane 0 does the (sequential) scalar arithmetic; the other lanes only prefetch: every run of global loads the
// recurrences need (the sums of one block, its cached Gram entries and scales; 32 rotation coefficients and column
// entries at a time) is fetched by the whole warp in one round trip into shared memory.  A single thread walking the
// same data paid one DRAM/L2 latency per load: 19-23 us per iteration at basis size 20, 3.6 % of a step at n = 2^24.
__global__ void __launch_bounds__(32) k_gmres_givens(KrylovCtl* ctl, int k, int64_t nr, double* R, double* c, double* s,
                                                     double* z, double* hcol, int reorth, int blk, double* hist,
                                                     int64_t hist_pos, int inner_limit, KrylovStatus* st, const P2PDev pd,
                                                     unsigned long long seq_in, double* rho_vec, double* gram,
                                                     double* rinv_vec) {
    __shared__ double s_t[kBlkSums], s_g[kBlkMax * kBlkMax], s_hb[kBlkMax], s_cb[kBlkMax];
    __shared__ double s_c[32], s_s[32], s_r[33];
    __shared__ int s_abort;
    const int lane = threadIdx.x;
    if (ctl->stop) {
        // iterations queued behind the verdict that ended the pass repeat that verdict, so that the host can never read
        // a stale record of an earlier lap of the ring as "keep going" (it reads the records in order)
        if (lane == 0) {
            st->rNorm = ctl->rNorm;
            st->Hbis = ctl->Hbis;
            st->iter = ctl->inner_iter;
            st->solved = ctl->solved;
            st->breakdown = ctl->breakdown;
            st->stop = 1;
        }
        return;
    }
    const int nblk = blk > 0 ? (k + blk - 1) / blk : 0;  // blocks of the blocked Gram-Schmidt sweep
    const int mlast = blk > 0 ? k - (nblk - 1) * blk : 0;  // vectors in the last block = subtracted by the final pass
    const int nsweep = reorth ? 2 : 1;
    const int rec_final = nsweep * nblk;  // record of the final pass: ||q||^2 and the new Gram entries
    if (lane == 0) s_abort = 0;
    __syncwarp();
    if (seq_in != 0 && lane == 0) {
        // ||q||^2 arrives through the mailboxes (posted by the final Gram-Schmidt pass of every rank); adding
        // in rank order gives the same bits on every rank, so all ranks take the same decisions below
        const int slot = (int)(seq_in % kMailSlots);
        double tot[kBlkSums];
        for (int q = 0; q < kBlkSums; ++q) tot[q] = 0.0;
        for (int q = 0; q < pd.nranks && !s_abort; ++q) {
            const double* rec = pd.mail_local + ((size_t)slot * pd.nranks + q) * kMailRec;
            const unsigned long long* tag = reinterpret_cast<const unsigned long long*>(rec + kBlkSums);
            const long long t0 = clock64();
            while (ld_acquire_sys_u64(tag) != seq_in) {
                if (clock64() - t0 > pd.spin_cycles) { *pd.err = 1; s_abort = 1; break; }
            }
            if (!s_abort)
                for (int q2 = 0; q2 <= mlast; ++q2) tot[q2] += __ldcv(rec + q2);  // ||q||^2 and the new Gram entries
        }
        if (s_abort) {  // a peer fell out of step: end the pass here, the host reports the error
            ctl->stop = 1;
            st->rNorm = ctl->rNorm; st->Hbis = ctl->Hbis; st->iter = ctl->inner_iter;
            st->solved = 0; st->breakdown = 0; st->stop = 1;
        } else {
            for (int q = 0; q <= mlast; ++q) hcol[kBlkSums * rec_final + q] = tot[q];
        }
    }
    __syncwarp();
    if (s_abort) return;
    // column k of H: h_1k..h_kk from the MGS sweep(s), h_{k+1,k} = ||q||
    if (blk > 0) {  // raw sums of the blocked sweep(s) (one record of kBlkSums per block and sweep), then ||q||^2
        for (int sw = 0; sw < nsweep; ++sw) {
            for (int j = 0; j < nblk; ++j) {
                const int m = (k - j * blk) < blk ? (k - j * blk) : blk;
                // the same warp-cooperative routine the vector kernels ran on these sums (same bits)
                if (lane < kBlkSums) s_t[lane] = hcol[kBlkSums * (sw * nblk + j) + lane];
                __syncwarp();
                block_coefficients_warp(lane, s_t, gram + (size_t)j * blk * kBlkMax, rho_vec ? rho_vec + j * blk : nullptr, m,
                                        s_g, s_hb, s_cb);
                // second sweep: R[nr+i] += Htmp (gmres! step 5)
                if (lane < m) R[nr + j * blk + lane] = sw == 0 ? s_hb[lane] : R[nr + j * blk + lane] + s_hb[lane];
                __syncwarp();
            }
        }
    } else {
        const double* h2 = hcol + (k + 1);
        for (int i = lane; i < k; i += 32) R[nr + i] = reorth ? hcol[i] + h2[i] : hcol[i];
    }
    __syncwarp();
    double hh = 0.0;
    if (lane == 0) {
        hh = blk > 0 ? hcol[kBlkSums * rec_final] : (reorth ? hcol[(k + 1) + k] : hcol[k]);
        // the vector finished by this iteration is stored as basis vector k; when it joins the last block (block not
        // full yet) its Gram entries with that block's vectors were measured by the final pass: cache them
        if (blk > 0 && mlast < blk)
            for (int a = 0; a < mlast; ++a) gram[(size_t)k * kBlkMax + a] = hcol[kBlkSums * rec_final + 1 + a];
    }
    givens_tail(lane, ctl, k, nr, R, c, s, z, hh, rho_vec, rinv_vec, hist, hist_pos, inner_limit, st, s_c, s_s, s_r);
}

// Scalar step of the one-sweep Gram-Schmidt (fuse = AK_FUSE_SWEEP, sweep.cu).  One warp.
//   mode 0  after the opening sweep of a cycle (k = 0): only the multipliers of iteration 1
//   mode 1  after the first sweep of a re-orthogonalised iteration k: multipliers of its second sweep (gmres! step 5)
//   mode 2  after the (last) sweep of iteration k: rho_k, row k of the Gram matrix, column k of R, Givens, stopping
//           tests, then the multipliers of iteration k + 1 from the projections the same sweep measured
// sums (layout sweep.h): [0] ||z||^2, [1] <z,y>, [2 + j] g_j = <S_j, z>, [2 + kSwKMax + j] t_j = <S_j, y>, y = J z raw.
// With v_j = S_j / rho_j and w = y / rho_k:  <v_j, w> = t_j / rho_j / rho_k,  <v_j, v_a> = gam[j][a], and
//   h_j = <v_j, w> - sum_{a<j} h_a gam[j][a]        (modified Gram-Schmidt, forward substitution over the whole cycle)
//   c_j = h_j / rho_j                               (what multiplies the stored vector in the next sweep)
// The substitution is column-oriented (lane j owns h_j; after step a every later lane subtracts h_a gam[j][a]): each h_j
// sees its subtractions in the order a = 0, 1, ... like the serial loop.
// seq_in != 0: the sums arrive through the sweep mailboxes and are added in rank order (same bits on every rank).
__global__ void __launch_bounds__(32) k_gmres_sweep_scalar(KrylovCtl* ctl, int k, int mode, int64_t nr, double* R, double* c,
                                                           double* s, double* z, double* sums, double* hvec, double* cvec,
                                                           double* gam, double* rho_vec, double* rinv_vec, double* gram_blk,
                                                           double* hist, int64_t hist_pos, int inner_limit,
                                                           KrylovStatus* st, const P2PDev pd, const double* swmail_local,
                                                           unsigned long long seq_in) {
    constexpr int LD = kSwKMax + 1;
    __shared__ double s_gam[LD * LD];
    __shared__ double s_h[32], s_sum[kSwSums];
    __shared__ double s_c[32], s_s[32], s_r[33];
    __shared__ int s_abort;
    const int lane = threadIdx.x;
    if (ctl->stop) {
        if (mode == 2 && lane == 0) {  // repeat the verdict that ended the pass (see k_gmres_givens)
            st->rNorm = ctl->rNorm;
            st->Hbis = ctl->Hbis;
            st->iter = ctl->inner_iter;
            st->solved = ctl->solved;
            st->breakdown = ctl->breakdown;
            st->stop = 1;
        }
        return;
    }
    if (lane == 0) s_abort = 0;
    __syncwarp();
    if (seq_in != 0) {
        const int slot = (int)(seq_in % kMailSlots);
        if (lane < pd.nranks) {
            const double* rec = swmail_local + ((size_t)slot * pd.nranks + lane) * kSwMailRec;
            const unsigned long long* tag = reinterpret_cast<const unsigned long long*>(rec + kSwSums);
            const long long t0 = clock64();
            while (ld_acquire_sys_u64(tag) != seq_in) {
                if (clock64() - t0 > pd.spin_cycles) { *pd.err = 1; s_abort = 1; break; }
            }
        }
        __syncwarp();
        if (!s_abort) {
            for (int q = lane; q < kSwSums; q += 32) {
                double a = 0.0;
                for (int r = 0; r < pd.nranks; ++r)
                    a += __ldcv(swmail_local + ((size_t)slot * pd.nranks + r) * kSwMailRec + q);
                s_sum[q] = a;
                sums[q] = a;
            }
        } else if (lane == 0) {  // a peer fell out of step: end the pass here, the host reports the error
            ctl->stop = 1;
            st->rNorm = ctl->rNorm; st->Hbis = ctl->Hbis; st->iter = ctl->inner_iter;
            st->solved = 0; st->breakdown = 0; st->stop = 1;
        }
    } else {
        for (int q = lane; q < kSwSums; q += 32) s_sum[q] = sums[q];
    }
    // the cached (scaled) Gram rows 1..k: one round trip for the whole warp
    for (int q = lane; q < (k + 1) * LD; q += 32) s_gam[q] = gam[q];
    __syncwarp();
    if (s_abort) return;
    const int nh = (mode == 1) ? k : k + 1;  // multipliers to form
    if (mode == 2) {
        const double hh = s_sum[0];
        const double Hbis = sqrt(hh);
        // row k of the Gram matrix, scaled: <v_k, v_a> = <S_k, S_a> / rho_k / rho_a; raw entries of S_k's own block of
        // eight for the pass-wise sweeps (a cycle that outgrows kSwKMax vectors continues with them)
        if (lane < k) {
            const double g = s_sum[2 + lane];
            const double v = __ddiv_rn(__ddiv_rn(g, Hbis), rho_vec[lane]);
            s_gam[k * LD + lane] = v;
            gam[k * LD + lane] = v;
            const int b0 = (k / kBlkMax) * kBlkMax;
            if (lane >= b0) gram_blk[(size_t)k * kBlkMax + (lane - b0)] = g;
            R[nr + lane] = hvec[lane];
        }
        __syncwarp();
        int stop = givens_tail(lane, ctl, k, nr, R, c, s, z, hh, rho_vec, rinv_vec, hist, hist_pos, inner_limit, st, s_c,
                               s_s, s_r);
        stop = __shfl_sync(0xffffffffu, stop, 0);
        if (stop) return;
        __syncwarp();
    }
    // raw projections -> <v_j, w>
    if (lane < nh) {
        double pj;
        if (mode == 1) {
            pj = __ddiv_rn(s_sum[2 + lane], rho_vec[lane]);  // second sweep: w is already in the scale of v
        } else {
            const double tj = (lane == k) ? s_sum[1] : s_sum[2 + kSwKMax + lane];
            pj = __dmul_rn(__ddiv_rn(tj, rho_vec[lane]), rinv_vec[k]);
        }
        s_h[lane] = pj;
    }
    __syncwarp();
    for (int a = 0; a + 1 < nh; ++a) {
        const double ha = s_h[a];
        if (lane > a && lane < nh) s_h[lane] = __dsub_rn(s_h[lane], __dmul_rn(ha, s_gam[lane * LD + a]));
        __syncwarp();
    }
    if (lane < nh) {
        const double h = s_h[lane];
        hvec[lane] = (mode == 1) ? hvec[lane] + h : h;  // R[nr+i] += Htmp
        cvec[lane] = __ddiv_rn(h, rho_vec[lane]);
    }
}

// gmres! step 9: solve R y = z (K x K packed column-major upper triangle, K = inner iterations of the pass, read from
// the control block so that the host can queue this kernel before it knows how far the pass got).  Column-oriented
// back-substitution: y_j = z_j / R_jj, then z_i -= R_ij y_j for all i < j in parallel.  Every y_i therefore sees the
// subtractions in the order j = K, K-1, ..., i+1 with one multiplication and one subtraction each: the same operations
// in the same order as the row-oriented loop of Krylov.jl, |R_ii| <= btol  =>  y_i = 0 and `inconsistent`.
// use_rho: un-normalised basis, the combine kernel multiplies the STORED vectors: y_i <- y_i / rho_i.
__global__ void __launch_bounds__(256) k_gmres_backsolve(KrylovCtl* ctl, const double* __restrict__ R,
                                                         const double* __restrict__ z, const double* __restrict__ rho,
                                                         double* y, int use_rho, KrylovStatus* st) {
    __shared__ double s_yj;
    const int K = ctl->inner_iter;
    const int tid = threadIdx.x;
    const double btol = ctl->btol;
    for (int i = tid; i < K; i += blockDim.x) y[i] = z[i];
    __syncthreads();
    for (int j = K - 1; j >= 0; --j) {
        const int64_t col = (int64_t)j * (j + 1) / 2;  // column j holds R[0..j, j]
        if (tid == 0) {
            const double diag = R[col + j];
            double yj;
            if (fabs(diag) <= btol) { yj = 0.0; ctl->inconsistent = 1; }
            else yj = y[j] / diag;
            y[j] = yj;
            s_yj = yj;
        }
        __syncthreads();
        const double yj = s_yj;
        for (int i = tid; i < j; i += blockDim.x) y[i] = __dsub_rn(y[i], __dmul_rn(R[col + i], yj));
        __syncthreads();
    }
    if (use_rho)
        for (int i = tid; i < K; i += blockDim.x) y[i] = y[i] / rho[i];
    if (tid == 0) st->inconsistent = ctl->inconsistent;
}

// CG scalar updates (Krylov.jl cg!, M = I, radius = 0, linesearch = false)
__global__ void k_cg_begin(KrylovCtl* ctl, const double* sumsq, double* hist, double atol, double rtol,
                           KrylovStatus* st) {
    if (threadIdx.x != 0) return;
    const double gamma = *sumsq;
    const double rNorm = sqrt(gamma);
    ctl->gamma = gamma;
    ctl->pNorm2 = gamma;
    ctl->rNorm = rNorm;
    ctl->beta = rNorm;
    ctl->eps = atol + rtol * rNorm;
    ctl->solved = (rNorm <= ctl->eps) || (gamma == 0.0);
    ctl->zerocurv = 0;
    ctl->inconsistent = 0;
    ctl->stop = ctl->solved;
    ctl->inner_iter = 0;
    if (hist) hist[0] = rNorm;
    st->rNorm = rNorm; st->beta = rNorm; st->iter = 0; st->stop = ctl->stop; st->solved = ctl->solved; st->zerocurv = 0;
}
// after pAp = <p, Ap>
__global__ void k_cg_alpha(KrylovCtl* ctl, const double* pAp_dev, KrylovStatus* st) {
    if (threadIdx.x != 0) return;
    if (ctl->stop) return;
    const double pAp = *pAp_dev;
    const double epsm = 2.220446049250313e-16;
    ctl->pAp = pAp;
    if (pAp <= epsm * ctl->pNorm2 && fabs(pAp) <= epsm * ctl->pNorm2) {
        ctl->zerocurv = 1;
        ctl->inconsistent = 1;
        ctl->stop = 1;
        st->rNorm = ctl->rNorm;
        st->iter = ctl->inner_iter;
        st->solved = 0;
        st->stop = 1;
        st->zerocurv = 1;
        return;
    }
    ctl->alpha = ctl->gamma / pAp;
    ctl->neg_alpha = -ctl->alpha;
}
// after gamma_next = <r, r>
__global__ void k_cg_beta(KrylovCtl* ctl, const double* gnext_dev, int iter, int64_t itmax, double* hist,
                          KrylovStatus* st) {
    if (threadIdx.x != 0) return;
    if (ctl->stop) return;
    const double gn = *gnext_dev;
    const double rNorm = sqrt(gn);
    if (hist) hist[iter] = rNorm;
    const int mach = (rNorm + 1.0 <= 1.0);
    const int solved = (rNorm <= ctl->eps) || mach;
    ctl->rNorm = rNorm;
    ctl->solved = solved;
    if (!solved) {
        const double beta = gn / ctl->gamma;
        ctl->pNorm2 = gn + beta * beta * ctl->pNorm2;
        ctl->gamma = gn;
        ctl->cg_beta = beta;
    }
    ctl->inner_iter = iter;
    const int stop = solved || (iter >= itmax);
    ctl->stop = stop;
    st->rNorm = rNorm; st->iter = iter; st->stop = stop; st->solved = solved; st->zerocurv = 0;
}
// Tail of one CG iteration in one pass (cg!: x += alpha p ... p = r + beta p):
//   x <- x + alpha p      for the iteration that has just been judged (also when it was the last one)
//   p <- r + beta p       unless the solve stopped
// 40n bytes (x and p read + written, r read) instead of 24n + 24n for the two separate updates; 256-bit accesses.
// `k`: the iteration this launch belongs to; when the control block did not advance to k (zero curvature, or a launch
// queued behind the stop) the kernel is a no-op.
template <bool VEC>
__global__ void __launch_bounds__(256) k_cg_update_xp(double* __restrict__ x, double* __restrict__ p,
                                                      const double* __restrict__ r, const KrylovCtl* __restrict__ ctl,
                                                      int k, int64_t n) {
    if (ctl->inner_iter != k) return;
    const bool upd_p = ctl->stop == 0;
    const double alpha = ctl->alpha, beta = ctl->cg_beta;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            const int64_t j = i << 2;
            d4 xv = ld4(x + j), pv = ld4(p + j);
            xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y); xv.z = fma(alpha, pv.z, xv.z); xv.w = fma(alpha, pv.w, xv.w);
            st4(x + j, xv);
            if (upd_p) {
                const d4 rv = ld4_stream(r + j);
                pv.x = fma(beta, pv.x, rv.x); pv.y = fma(beta, pv.y, rv.y); pv.z = fma(beta, pv.z, rv.z); pv.w = fma(beta, pv.w, rv.w);
                st4(p + j, pv);
            }
        }
        const int64_t j = (n4 << 2) + tid;
        if (j < n) {
            const double pj = p[j];
            x[j] = fma(alpha, pj, x[j]);
            if (upd_p) p[j] = fma(beta, pj, r[j]);
        }
    } else {
        for (int64_t j = tid; j < n; j += nth) {
            const double pj = p[j];
            x[j] = fma(alpha, pj, x[j]);
            if (upd_p) p[j] = fma(beta, pj, r[j]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Small-problem regime (BASELINE config 1: 1-D Bratu, N = 10^4 — 80 KB per vector): the multi-kernel CG above is
// launch-bound there (five launches of a few microseconds per iteration for ~100 ns of memory traffic).  This is the
// whole of cg! as ONE persistent block: p and r live in shared memory (the three-point tangent reads p's neighbours
// there), x and lambda e^u in registers, the two inner products are block reductions, and the block iterates until the
// stopping test of cg! fires — no launch, no host round trip inside the solve.  Same operations per element as the
// multi-kernel path (fma forms, exact division by dx^2); only the summation order of the inner products differs.
// Point i = j * kSmallThreads + tid belongs to thread tid (bank-conflict-free, neighbours are neighbouring threads).
// ---------------------------------------------------------------------------------------
constexpr int kSmallThreads = 1024;
constexpr int kSmallNPT = 10;                                   // points per thread (x and lambda e^u: 40 of the 64 registers)
constexpr int64_t kSmallN = (int64_t)kSmallThreads * kSmallNPT; // 10240 unknowns: p, r = 2 x 80 KB of shared memory

AK_DEV double block_sum_all(double v, double* sh, double* bcast) {  // sum over the block, result in every thread
    const double r = block_sum(v, sh);
    if (threadIdx.x == 0) *bcast = r;
    __syncthreads();
    return *bcast;
}

__global__ void __launch_bounds__(kSmallThreads, 1) k_cg_small_bratu1d(int n, double dx2v, double lambda,
                                                                       const double* __restrict__ aux, int coef_from_u,
                                                                       const double* __restrict__ b, double* __restrict__ x_out,
                                                                       KrylovCtl* ctl, KrylovStatus* st, double* hist,
                                                                       double atol, double rtol, long long itmax) {
    extern __shared__ double smem[];
    double* s_p = smem + 1;          // s_p[-1] and s_p[n] are the Dirichlet zeros (bratu.jl:17-18)
    double* s_r = smem + (n + 2);
    __shared__ double sh[32];
    __shared__ double bc;
    const int tid = threadIdx.x;
    const Divisor dx2 = make_divisor(dx2v);
    double x[kSmallNPT], kc[kSmallNPT];
    if (tid == 0) { s_p[-1] = 0.0; s_p[n] = 0.0; }
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < kSmallNPT; ++j) {
        const int i = j * kSmallThreads + tid;
        x[j] = 0.0;
        kc[j] = 0.0;
        if (i < n) {
            const double a = aux[i];
            kc[j] = coef_from_u ? __dmul_rn(lambda, exp(a)) : a;
            const double bi = b[i];
            s_r[i] = bi;          // r = b (x0 = 0)
            s_p[i] = bi;          // p = r
            acc = fma(bi, bi, acc);
        }
    }
    __syncthreads();
    double gamma = block_sum_all(acc, sh, &bc);
    double rNorm = sqrt(gamma);
    const double beta0 = rNorm;
    const double eps = atol + rtol * rNorm;
    double pNorm2 = gamma;
    int solved = (rNorm <= eps) || (gamma == 0.0);
    int zerocurv = 0;
    long long iter = 0;
    if (hist != nullptr && tid == 0) hist[0] = rNorm;
    const double epsm = 2.220446049250313e-16;
    auto tangent = [&](int i, int j) -> double {  // (J p)_i, the 1-D Bratu tangent of stencil.cu
        const double c = s_p[i];
        return __dadd_rn(second_diff(s_p[i + 1], c, s_p[i - 1], dx2), __dmul_rn(kc[j], c));
    };
    while (!solved && iter < itmax) {
        // pAp = <p, A p>
        acc = 0.0;
#pragma unroll
        for (int j = 0; j < kSmallNPT; ++j) {
            const int i = j * kSmallThreads + tid;
            if (i < n) acc = fma(s_p[i], tangent(i, j), acc);
        }
        const double pAp = block_sum_all(acc, sh, &bc);
        if (pAp <= epsm * pNorm2 && fabs(pAp) <= epsm * pNorm2) { zerocurv = 1; break; }
        const double alpha = gamma / pAp;
        // x += alpha p ; r -= alpha A p ; gamma_next = <r, r>
        acc = 0.0;
#pragma unroll
        for (int j = 0; j < kSmallNPT; ++j) {
            const int i = j * kSmallThreads + tid;
            if (i < n) {
                const double ap = tangent(i, j);
                x[j] = fma(alpha, s_p[i], x[j]);
                const double ri = fma(-alpha, ap, s_r[i]);
                s_r[i] = ri;
                acc = fma(ri, ri, acc);
            }
        }
        const double gn = block_sum_all(acc, sh, &bc);  // (its barrier also orders the reads of p before the update below)
        iter += 1;
        rNorm = sqrt(gn);
        if (hist != nullptr && tid == 0) hist[iter] = rNorm;
        solved = (rNorm <= eps) || (rNorm + 1.0 <= 1.0);
        if (solved) break;
        const double beta = gn / gamma;
        pNorm2 = gn + beta * beta * pNorm2;
        gamma = gn;
        // p = r + beta p
#pragma unroll
        for (int j = 0; j < kSmallNPT; ++j) {
            const int i = j * kSmallThreads + tid;
            if (i < n) s_p[i] = fma(beta, s_p[i], s_r[i]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < kSmallNPT; ++j) {
        const int i = j * kSmallThreads + tid;
        if (i < n) x_out[i] = x[j];
    }
    if (tid == 0) {
        ctl->rNorm = rNorm; ctl->beta = beta0; ctl->solved = solved; ctl->zerocurv = zerocurv;
        ctl->inconsistent = zerocurv; ctl->inner_iter = (int)iter; ctl->stop = 1;
        st->rNorm = rNorm; st->beta = beta0; st->iter = (int)iter; st->stop = 1; st->solved = solved; st->zerocurv = zerocurv;
    }
}

// ---------------------------------------------------------------------------------------
// workspace memory
// ---------------------------------------------------------------------------------------
static int ws_alloc_vec(ak_krylov* ws, double** out) {
    double* p = nullptr;
    cudaError_t e = pool_alloc(ws->ctx, (void**)&p, sizeof(double) * (size_t)(ws->n > 0 ? ws->n : 1));
    if (e != cudaSuccess) {
        set_error("krylov workspace: cudaMalloc of %lld doubles failed: %s", (long long)ws->n, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return AK_ERR_NOMEM;
    }
    ws->chunks.push_back(p);
    *out = p;
    return AK_OK;
}

// doubles of the device column `hcol`: two sweeps of k + 1 sums, or one record of kBlkSums per block of the
// blocked sweep (blocks of >= 2) plus the final ||q||^2 record
static inline int64_t hcol_len(int64_t k) {
    // blocks of 2, two sweeps (re-orthogonalisation): 2 ceil(k/2) records + the final one
    const int64_t a = 2 * (k + 1) + 8, b = (int64_t)kBlkSums * (2 * ((k + 1) / 2) + 4);
    return a > b ? a : b;
}

static int ws_grow_scalars(ak_krylov* ws, int64_t kcap_new) {
    Ctx* c = ws->ctx;
    if (kcap_new <= ws->kcap) return AK_OK;
    AK_CUDA(cudaStreamSynchronize(c->stream));
    const int64_t nR = kcap_new * (kcap_new + 1) / 2;
    auto regrow = [&](double** arr, int64_t old_n, int64_t new_n) -> int {
        double* q = nullptr;
        AK_CUDA(pool_alloc(c, (void**)&q, sizeof(double) * (size_t)new_n));
        AK_CUDA(cudaMemsetAsync(q, 0, sizeof(double) * (size_t)new_n, c->stream));
        if (*arr && old_n > 0)
            AK_CUDA(cudaMemcpyAsync(q, *arr, sizeof(double) * (size_t)old_n, cudaMemcpyDeviceToDevice, c->stream));
        if (*arr) AK_CUDA(cudaFreeAsync(*arr, c->stream));
        *arr = q;
        return AK_OK;
    };
    const int64_t oldk = ws->kcap;
    AK_TRY(regrow(&ws->R, oldk * (oldk + 1) / 2, nR));
    AK_TRY(regrow(&ws->c, oldk, kcap_new));
    AK_TRY(regrow(&ws->s, oldk, kcap_new));
    AK_TRY(regrow(&ws->z, oldk ? oldk + 1 : 0, kcap_new + 1));
    AK_TRY(regrow(&ws->rho, oldk ? oldk + 1 : 0, kcap_new + 1));
    AK_TRY(regrow(&ws->rinv, oldk ? oldk + 1 : 0, kcap_new + 1));
    AK_TRY(regrow(&ws->ycoef, oldk ? oldk + 1 : 0, kcap_new + 1));
    AK_TRY(regrow(&ws->gram, oldk ? (oldk + 1) * kBlkMax : 0, (kcap_new + 1) * kBlkMax));
    AK_TRY(regrow(&ws->hcol, oldk ? hcol_len(oldk) : 0, hcol_len(kcap_new)));
    ws->kcap = kcap_new;
    return AK_OK;
}

static int ws_grow_hist(ak_krylov* ws, int64_t need) {
    if (need <= ws->hist_cap) return AK_OK;
    Ctx* c = ws->ctx;
    int64_t nc = ws->hist_cap ? ws->hist_cap * 2 : 256;
    if (nc < need) nc = need;
    AK_CUDA(cudaStreamSynchronize(c->stream));
    double* q = nullptr;
    AK_CUDA(pool_alloc(c, (void**)&q, sizeof(double) * (size_t)nc));
    if (ws->hist) {
        AK_CUDA(cudaMemcpyAsync(q, ws->hist, sizeof(double) * (size_t)ws->hist_cap, cudaMemcpyDeviceToDevice, c->stream));
        AK_CUDA(cudaFreeAsync(ws->hist, c->stream));
    }
    ws->hist = q;
    ws->hist_cap = nc;
    return AK_OK;
}

// scalars of the one-sweep Gram-Schmidt (fixed size: a sweep handles at most kSwKMax basis vectors)
static int ws_ensure_sweep(ak_krylov* ws) {
    if (ws->sw_sums) return AK_OK;
    Ctx* c = ws->ctx;
    const size_t nd = (size_t)kSwSums + 2 * (kSwKMax + 1) + (size_t)(kSwKMax + 1) * (kSwKMax + 1);
    double* q = nullptr;
    AK_CUDA(pool_alloc(c, (void**)&q, sizeof(double) * nd));
    AK_CUDA(cudaMemsetAsync(q, 0, sizeof(double) * nd, c->stream));
    ws->sw_sums = q;
    ws->sw_h = q + kSwSums;
    ws->sw_c = ws->sw_h + (kSwKMax + 1);
    ws->sw_gam = ws->sw_c + (kSwKMax + 1);
    return AK_OK;
}

// make sure basis vectors V[0..count) exist
static int ws_ensure_basis(ak_krylov* ws, int64_t count, bool exact = false) {
    if ((int64_t)ws->V.size() >= count) return AK_OK;
    if (ws->max_basis > 0 && count > ws->max_basis) {
        set_error("krylov workspace: basis would grow past max_basis = %lld", (long long)ws->max_basis);
        return AK_ERR_NOMEM;
    }
    // grow in steps of `mem` vectors (Krylov.jl pushes one at a time; chunking amortises cudaMalloc)
    int64_t target = exact ? count : (int64_t)ws->V.size() + ws->mem;
    if (target < count) target = count;
    if (ws->max_basis > 0 && target > ws->max_basis) target = ws->max_basis;
    while ((int64_t)ws->V.size() < target) {
        double* v = nullptr;
        int rc = ws_alloc_vec(ws, &v);
        if (rc != AK_OK) {
            if ((int64_t)ws->V.size() >= count) break;  // enough for now
            return rc;
        }
        ws->V.push_back(v);
    }
    return AK_OK;
}

static int ws_upload_basis_table(ak_krylov* ws, int64_t k, bool use_z = false) {
    Ctx* c = ws->ctx;
    if (k > ws->V_dev_cap) {
        AK_CUDA(cudaStreamSynchronize(c->stream));
        if (ws->V_dev) AK_CUDA(cudaFreeAsync((void*)ws->V_dev, c->stream));
        int64_t cap = ws->V_dev_cap ? ws->V_dev_cap * 2 : 64;
        if (cap < k) cap = k;
        AK_CUDA(pool_alloc(c, (void**)&ws->V_dev, sizeof(double*) * (size_t)cap));
        ws->V_dev_cap = cap;
    }
    AK_CUDA(cudaMemcpyAsync((void*)ws->V_dev, use_z ? ws->Z.data() : ws->V.data(), sizeof(double*) * (size_t)k,
                            cudaMemcpyHostToDevice, c->stream));
    return AK_OK;
}

// ---------------------------------------------------------------------------------------
// GMRES
// ---------------------------------------------------------------------------------------
static int wait_status(ak_krylov* ws, int slot, KrylovStatus* out) {
    AK_CUDA(cudaEventSynchronize(ws->ev[slot]));
    memcpy(out, (const void*)&ws->status[slot], sizeof(KrylovStatus));
    return AK_OK;
}

static int gmres_solve(ak_krylov* ws, const ak_problem* prob, const double* u, const double* b,
                       const ak_krylov_opts* o, ak_krylov_stats* st, double* hist_host, int64_t hist_cap);

// out <- P in for the right (N) or left (M) preconditioner of this solve.
// AK_PRECOND_INNER_GMRES: `copyto!(y, gmres(P.J, x; P.itmax)[1])` (examples/bratu.jl:146-149): an inner GMRES with
// memory 20, default tolerances, x0 = 0 and at most itmax iterations.  Other kinds: precond.cu.
static int apply_precond(ak_krylov* ws, const ak_problem* prob, const double* u, const ak_krylov_opts* o, bool left,
                         const double* in, double* out) {
    const int32_t kind = left ? o->precond_m : o->precond_n;
    if (kind != AK_PRECOND_INNER_GMRES)
        return precond_apply(ws->ctx, prob, u, kind, left ? o->m_apply : o->n_apply, left ? o->m_user : o->n_user, in,
                             out);
    if (!ws->inner) {
        AK_TRY(ak_krylov_create(ws->owner, AK_ALGO_GMRES, ws->n, 20, 0, &ws->inner));
    }
    ak_krylov_opts io;
    ak_krylov_default_opts(&io);
    io.itmax = left ? o->precond_m_itmax : o->precond_itmax;
    io.fuse = (o->fuse == AK_FUSE_FULL) ? AK_FUSE_MGS : o->fuse;  // the inner solve runs at the outer solve's fusion level
    ak_krylov_stats ist;
    int rc = gmres_solve(ws->inner, prob, u, in, &io, &ist, nullptr, 0);
    if (rc < 0) return rc;
    return launch_copy(ws->ctx, ws->n, out, ws->inner->x);
}
static int apply_precond_n(ak_krylov* ws, const ak_problem* prob, const double* u, const ak_krylov_opts* o,
                           const double* in, double* out) {
    return apply_precond(ws, prob, u, o, false, in, out);
}

static int gmres_solve(ak_krylov* ws, const ak_problem* prob, const double* u, const double* b,
                       const ak_krylov_opts* o, ak_krylov_stats* st, double* hist_host, int64_t hist_cap) {
    Ctx* c = ws->ctx;
    const int64_t n = ws->n;
    const int mem = ws->mem;
    const int restart = o->restart, reorth = o->reorthogonalization;
    int fuse = o->fuse;
    const bool flexible = (ws->algo == AK_ALGO_FGMRES);
    const bool precond = (o->precond_n != AK_PRECOND_NONE);
    const bool lprec = (o->precond_m != AK_PRECOND_NONE);  // q = M A N v_k, r0 = M (b - A x): Krylov.jl solver.q
    // z_k = N v_k needs v_k materialised before the JVP and a host decision per iteration (the preconditioner may be
    // an inner solve with its own verdicts): no JVP fusion.  The blocked sweeps still apply: they orthogonalise
    // against the stored (un-normalised) vectors, only the seed of the preconditioner is a normalised copy.
    const bool hosted = flexible || precond || lprec;
    if (hosted && fuse == AK_FUSE_FULL) fuse = AK_FUSE_MGS;
    // One sweep per iteration (sweep.cu) where it applies: 2-D analytic tangents, no preconditioner, the first kSwKMax
    // iterations of a pass.  Everything else of such a solve runs the eight-step blocked passes on the same
    // (un-normalised) basis.  AK_NO_SWEEP=1: developer switch for A/B timing.
    bool sweep = false;
    if (fuse == AK_FUSE_SWEEP) {
        static const bool no_sweep = getenv("AK_NO_SWEEP") != nullptr;
        sweep = !no_sweep && !hosted && sweep_supported(c, prob, u);
        if (c->nranks > 1 && !hosted) {
            // the ranks must take the same path (their kernels wait for one another), and the local test can differ:
            // segment lengths differ by a point between ranks, so one rank's row length may be odd
            int rc_sw = sweep ? AK_OK : AK_ERR_UNSUPPORTED;
            AK_TRY(collective_verdict(c, &rc_sw));
            sweep = (rc_sw == AK_OK);
        }
        fuse = AK_FUSE_BLOCK8;
    }
    if ((flexible || precond) && !ws->pbuf) AK_TRY(ws_alloc_vec(ws, &ws->pbuf));
    if (lprec && !ws->qbuf) AK_TRY(ws_alloc_vec(ws, &ws->qbuf));
    const int blk = fuse == AK_FUSE_PAIR ? 2 : (fuse == AK_FUSE_BLOCK4 ? 4 : (fuse == AK_FUSE_BLOCK8 ? 8 : 0));  // steps per sweep
    const bool pair = blk > 0;
    // Blocked sweeps keep the basis un-normalised: iteration k works in place on basis slot k, whose finished content
    // IS the stored vector (scale rho[k] = Hbis); the JVP divides by rho[k-1] in registers.  No w buffers, no
    // normalised copy: 24n instead of 32n bytes per JVP.
    const bool raw = pair;
    ws->last_raw = raw;
    // multi-GPU with peer memory: reductions and ghost rows of the blocked sweep go over NVLink stores
    const bool p2p = pair && c->p2p_on && c->nranks > 1;
    const bool is2d = (prob->kind == AK_BRATU2D || prob->kind == AK_HEAT2D);
    const bool p2p_halo = p2p && !hosted && is2d && prob->nx % 4 == 0 && n % 4 == 0 && prob->nx <= c->p2p_halo_cap;
    int nb_down = -1, nb_up = -1;  // owners of the ghost rows below / above this slab
    if (p2p_halo || (sweep && p2p)) {
        const bool per = is2d ? (prob->bc == AK_BC_PERIODIC) : (prob->kind == AK_HEAT1D_DG);
        nb_down = c->rank > 0 ? c->rank - 1 : (per ? c->nranks - 1 : -1);
        nb_up = c->rank < c->nranks - 1 ? c->rank + 1 : (per ? 0 : -1);
    }
    const bool want_hist = (hist_host != nullptr && hist_cap > 0) || o->history;
    cudaStream_t sm = c->stream;

    int64_t itmax = o->itmax == 0 ? 2 * ws->n_global : o->itmax;  // identical on every rank (the ranks must stop together)
    double* x = ws->x;
    // right-preconditioned gmres!: x += N (V y) goes through a separate xr; otherwise the combine kernel adds to x directly
    const bool xr_separate = precond && !flexible;
    double* xr = x;
    if (restart && xr_separate) {
        if (!ws->dx) AK_TRY(ws_alloc_vec(ws, &ws->dx));
        xr = ws->dx;
    }
    if ((fuse == AK_FUSE_FULL || sweep) && !ws->w[1]) AK_TRY(ws_alloc_vec(ws, &ws->w[1]));
    if (sweep) AK_TRY(ws_ensure_sweep(ws));
    AK_TRY(ws_ensure_basis(ws, raw && restart ? mem + 1 : mem, raw));
    // problem kinds without a fused normalise + JVP kernel get the normalised seed in a scratch vector
    const bool raw_needs_seed = raw && !hosted && (prob->kind == AK_SIMPLE2 || prob->kind == AK_USER ||
                                                   prob->scheme == AK_MIDPOINT || prob->jvp_mode == AK_JVP_FD);
    // 2-D analytic tangents: the first projection pass of every iteration is folded into the tangent kernel
    // (AK_NO_PROJ_FUSION=1 keeps the separate pass: a developer switch for A/B timing)
    static const bool no_proj_fusion = getenv("AK_NO_PROJ_FUSION") != nullptr;
    const bool proj_in_jvp = raw && !hosted && is2d && prob->scheme != AK_MIDPOINT && prob->jvp_mode == AK_JVP_ANALYTIC &&
                             !no_proj_fusion;
    if (raw_needs_seed && !ws->pbuf) AK_TRY(ws_alloc_vec(ws, &ws->pbuf));
    AK_TRY(ws_grow_scalars(ws, mem));
    if (want_hist) AK_TRY(ws_grow_hist(ws, 257));

    const bool peer = c->p2p_on && c->nranks > 1;  // some collective of this solve goes through the mailboxes
    if (peer && *c->p2p_err) {
        set_error("the peer-memory path of this context is latched off after a time-out; call ak_comm_use_p2p on every rank");
        return AK_ERR_PEER;
    }
    int wi = 0;  // index of the buffer currently holding w / r0
    double* w = raw ? ws->V[0] : ws->w[wi];
    if (xr_separate) AK_TRY(launch_fill(c, n, x, 0.0));
    if (lprec) AK_TRY(apply_precond(ws, prob, u, o, true, b, w));  // r0 = M b
    else AK_TRY(launch_copy(c, n, w, b));
    AK_TRY(launch_sumsq(c, n, w, ws->hcol));

    memset(st, 0, sizeof(*st));
    int64_t iter = 0, inner_itmax = itmax;
    int npass = 0;
    bool solved = false, tired = false, breakdown = false, inconsistent = false;
    bool x_written = false;  // x holds the sum of the passes so far (false: still conceptually zero)
    double rNorm = 0.0, beta0 = 0.0;
    ws->status[kStatusRing + 1].inconsistent = 0;

    // ---- one-sweep iterations (sweep.cu) -----------------------------------------------------------------------------
    const bool sw_p2p = sweep && p2p;
    // what the neighbours exchange: a boundary row (slabs), two points (1-D segments), one element (DG, periodic mesh)
    const int sw_cnt = is2d ? (int)prob->nx : (prob->kind == AK_HEAT1D_DG ? 4 : 2);
    const bool sw_per = is2d ? (prob->bc == AK_BC_PERIODIC) : (prob->kind == AK_HEAT1D_DG);
    unsigned long long sw_seq = 0;  // record the last sweep posted its sums under (peer memory), 0 otherwise
    auto sw_wslot = [&](int wbuf) -> int { return kSwKMax + 1 + wbuf; };  // ghost-row slot of W buffer `wbuf`
    std::vector<const double*> sw_lo((size_t)kSwKMax, nullptr), sw_hi((size_t)kSwKMax, nullptr);
    if (sw_p2p)
        for (int j = 0; j < kSwKMax; ++j) {
            sw_lo[(size_t)j] = nb_down >= 0 ? c->p2p_swghost_local(j, 0) : nullptr;
            sw_hi[(size_t)j] = nb_up >= 0 ? c->p2p_swghost_local(j, 1) : nullptr;
        }
    // one sweep over S_0..S_{kk-1}: z = zin * in_scale - sum c_j S_j -> zout; (stencil) y = J z -> yout.
    // zin_slot / zout_slot / yout_slot: ghost-row slots of those buffers on the peer-memory path (-1: none)
    auto sweep_launch = [&](int kk, const double* zin, int zin_slot, double* zout, int zout_slot, bool stencil,
                            double* yout, int yout_slot, const double* in_scale) -> int {
        SweepCall sc;
        sc.k = kk;
        sc.S = ws->V.data();
        sc.zin = zin;
        sc.zout = zout;
        sc.stencil = stencil;
        sc.yout = yout;
        sc.cvec = ws->sw_c;
        sc.in_scale = in_scale;
        sc.sums = ws->sw_sums;
        sc.stop = &ws->ctl->stop;
        if (sw_p2p) {
            sc.S_lo = sw_lo.data();
            sc.S_hi = sw_hi.data();
            sc.zin_lo = nb_down >= 0 ? c->p2p_swghost_local(zin_slot, 0) : nullptr;
            sc.zin_hi = nb_up >= 0 ? c->p2p_swghost_local(zin_slot, 1) : nullptr;
            if (zout != nullptr && zout_slot >= 0) {
                sc.push_z_down = nb_down >= 0 ? c->p2p_swghost_of(nb_down, zout_slot, 1) : nullptr;
                sc.push_z_up = nb_up >= 0 ? c->p2p_swghost_of(nb_up, zout_slot, 0) : nullptr;
            }
            if (stencil) {
                sc.push_y_down = nb_down >= 0 ? c->p2p_swghost_of(nb_down, yout_slot, 1) : nullptr;
                sc.push_y_up = nb_up >= 0 ? c->p2p_swghost_of(nb_up, yout_slot, 0) : nullptr;
            }
            sc.seq_out = ++c->p2p_seq;
            sw_seq = sc.seq_out;
        }
        return launch_sweep(c, prob, u, sc);
    };
    auto sweep_scalar = [&](int kk, int mode, int64_t nr, int64_t inner_limit, KrylovStatus* rec) -> int {
        ProfScope prof(c, PK_SCALAR);
        k_gmres_sweep_scalar<<<1, 32, 0, sm>>>(ws->ctl, kk, mode, nr, ws->R, ws->c, ws->s, ws->z, ws->sw_sums, ws->sw_h,
                                               ws->sw_c, ws->sw_gam, ws->rho, ws->rinv, ws->gram,
                                               want_hist ? ws->hist : nullptr, iter + kk, (int)inner_limit, rec,
                                               sw_p2p ? c->p2p_dev() : P2PDev{},
                                               sw_p2p ? c->p2p_swmail_of(c->rank) : nullptr, sw_p2p ? sw_seq : 0ull);
        c->launches++;
        AK_CUDA(cudaGetLastError());
        return AK_OK;
    };

    while (true) {
        // ---- pass prologue -------------------------------------------------------------
        if (restart && npass >= 1) {
            // w <- b - A x with ||w||^2 in the same pass (left-preconditioned: w <- M (b - A x))
            JvpFusion rf;
            rf.rhs_minus = b;
            if (lprec) {
                AK_TRY(launch_jvp(c, prob, u, x, ws->qbuf, &rf));
                AK_TRY(apply_precond(ws, prob, u, o, true, ws->qbuf, w));
                AK_TRY(launch_sumsq(c, n, w, ws->hcol));
            } else {
                rf.sumsq_dev = ws->hcol;
                AK_TRY(launch_jvp(c, prob, u, x, w, &rf));
            }
        }
        k_gmres_begin<<<1, 32, 0, sm>>>(ws->ctl, ws->hcol, ws->z, want_hist ? ws->hist : nullptr, npass == 0, o->atol,
                                        o->rtol, &ws->status[kStatusRing], raw ? ws->rho : nullptr, ws->rinv);
        c->launches++;
        AK_CUDA(cudaGetLastError());
        AK_CUDA(cudaEventRecord(ws->ev[kStatusRing], sm));
        // V[0] <- r0 / rNorm   (no-op when already converged / zero residual; un-normalised basis: V[0] is r0)
        if (!raw) AK_TRY(launch_divcopy_dev(c, n, ws->V[0], w, &ws->ctl->rNorm, &ws->ctl->stop));
        npass += 1;
        int wa = 0;  // one-sweep iterations: the W buffer that holds the raw tangent J S_{k-1} of the next iteration
        if (sweep) {
            // opening of the cycle: W <- J S_0 and <S_0, W>, queued behind the prologue (a no-op after a stop).  There is
            // no basis to sweep over yet, so this is the plain tangent kernel with its fused dot (0.26 ms at 8192^2; the
            // sweep kernel with an empty basis took 0.65 ms: its per-row bookkeeping has nothing to hide behind).
            JvpFusion of;
            of.stop_flag = &ws->ctl->stop;
            of.dot_with = ws->V[0];
            of.dot_dev = ws->sw_sums + 1;
            AK_TRY(launch_jvp(c, prob, u, ws->V[0], ws->w[wa], &of));
            if (sw_p2p) {  // slabs: the sweeps read the neighbours' boundary rows of S_0 and W from their ghost-row slots
                AK_TRY(sweep_push_rows(c, ws->V[0], n, sw_cnt, sw_per, 0));
                AK_TRY(sweep_push_rows(c, ws->w[wa], n, sw_cnt, sw_per, sw_wslot(wa)));
            }
            sw_seq = 0;  // the dot is already summed over the ranks
            AK_TRY(sweep_scalar(0, 0, 0, 0, &ws->status[kStatusRing]));
        }

        const int64_t inner_limit = restart ? (mem < inner_itmax ? mem : inner_itmax) : inner_itmax;
        int64_t K = 0;  // completed inner iterations of this pass
        KrylovStatus hs;
        bool update_queued = false;
        // x (+)= sum_i y_i V_i with y from the device back-substitution; the pass length is read on the device, so this
        // can be queued before the host has seen the last verdict (no stream drain at the cycle boundary)
        auto queue_solution_update = [&](int64_t k_launched) -> int {
            int64_t cnt = k_launched;
            const int64_t have = flexible ? (int64_t)ws->Z.size() : (int64_t)ws->V.size();
            if (cnt > have) cnt = have;
            if (cnt < 1) return AK_OK;
            AK_TRY(ws_upload_basis_table(ws, cnt, flexible));
            { ProfScope prof(c, PK_SCALAR);
            k_gmres_backsolve<<<1, 256, 0, sm>>>(ws->ctl, ws->R, ws->z, ws->rho, ws->ycoef, (raw && !flexible) ? 1 : 0,
                                                 &ws->status[kStatusRing + 1]); }
            c->launches++;
            AK_CUDA(cudaGetLastError());
            const int* kdev = &ws->ctl->inner_iter;
            if (xr_separate) {
                // x_k = N V_k y_k: xr <- V y ; xr <- N xr ; x += xr
                AK_TRY(launch_basis_combine(c, n, xr, ws->V_dev, ws->ycoef, kdev, 0, 0));
                AK_TRY(launch_copy(c, n, ws->pbuf, xr));
                AK_TRY(apply_precond_n(ws, prob, u, o, ws->pbuf, xr));
                if (restart) AK_TRY(launch_axpy(c, n, 1.0, xr, x));
            } else {
                // gmres: V y, fgmres: Z y; first pass stores, restart passes add (kaxpy!(n, one, xr, x))
                AK_TRY(launch_basis_combine(c, n, x, ws->V_dev, ws->ycoef, kdev, 0, x_written ? 1 : 0));
            }
            update_queued = true;
            return AK_OK;
        };
        AK_TRY(wait_status(ws, kStatusRing, &hs));
        if (npass == 1) { beta0 = hs.beta; rNorm = hs.rNorm; }
        if (hs.stop) {
            solved = hs.solved != 0;
            K = 0;
        } else {
            int64_t k = 0;
            bool scale_pending = false;  // FUSE_FULL: V[k] = w/Hbis is folded into the next JVP
            int64_t next_read = 1;       // oldest iteration whose verdict the host has not looked at yet
            // read the verdicts of iterations next_read..upto in order; true as soon as one of them stopped the pass
            auto verdicts_until = [&](int64_t upto, bool* stopped) -> int {
                *stopped = false;
                while (next_read <= upto) {
                    AK_TRY(wait_status(ws, (int)(next_read % kStatusRing), &hs));
                    next_read += 1;
                    if (hs.stop) { *stopped = true; break; }
                }
                return AK_OK;
            };
            bool stopped = false;
            while (true) {
                k += 1;
                // storage for this iteration (basis vector k receives q/Hbis, 0-based)
                if (!restart || k < mem || raw) {
                    if (k + 1 > (int64_t)ws->V.size()) {
                        int rc = ws_ensure_basis(ws, k + 1);
                        // every rank must take the same branch (the peer-memory kernels wait for one another)
                        AK_TRY(collective_verdict(c, &rc));
                        if (rc != AK_OK) {
                            // cannot grow further: treat as out of iterations (reference would keep growing)
                            AK_TRY(verdicts_until(k - 1, &stopped));
                            tired = true;
                            break;
                        }
                    }
                }
                if (k > ws->kcap) AK_TRY(ws_grow_scalars(ws, ws->kcap * 2 > k ? ws->kcap * 2 : k));
                if (want_hist) AK_TRY(ws_grow_hist(ws, iter + k + 1));
                const int64_t nr = k * (k - 1) / 2;
                const int* stop = &ws->ctl->stop;
                double* hcol = ws->hcol;
                unsigned long long givens_seq = 0;

                const int slot = (int)(k % kStatusRing);
                const bool sw_it = sweep && k <= kSwKMax;
                if (sw_it) {
                    // the whole iteration is one sweep over the basis (two with re-orthogonalisation); the tangent of the
                    // next iteration rides along unless this is the last iteration a sweep can serve
                    const bool next_sw = k < inner_limit && k + 1 <= kSwKMax;
                    if (reorth) {
                        // gmres! step 5: z' = w - V h (W[wa] -> W[wa^1]), then z = z' - V h' (-> V[k]) with the tangent -> W[wa]
                        AK_TRY(sweep_launch((int)k, ws->w[wa], sw_wslot(wa), ws->w[wa ^ 1], sw_wslot(wa ^ 1), false, nullptr, -1,
                                            ws->rinv + (k - 1)));
                        AK_TRY(sweep_scalar((int)k, 1, nr, inner_limit, &ws->status[slot]));
                        AK_TRY(sweep_launch((int)k, ws->w[wa ^ 1], sw_wslot(wa ^ 1), ws->V[k], (int)k, next_sw, ws->w[wa],
                                            sw_wslot(wa), nullptr));
                    } else {
                        AK_TRY(sweep_launch((int)k, ws->w[wa], sw_wslot(wa), ws->V[k], (int)k, next_sw, ws->w[wa ^ 1],
                                            sw_wslot(wa ^ 1), ws->rinv + (k - 1)));
                        wa ^= 1;
                    }
                    AK_TRY(sweep_scalar((int)k, 2, nr, inner_limit, &ws->status[slot]));
                    w = ws->V[k];
                } else {
                // fgmres / right preconditioning: z_k = N v_k (kept in Z for fgmres), then w <- A z_k
                double* pv = ws->V[k - 1];
                if (hosted && k > 1) {
                    // the preconditioner solves with host-visible verdicts: this path is not speculative
                    AK_TRY(verdicts_until(k - 1, &stopped));
                    if (stopped) break;
                }
                if (hosted && raw) {
                    // normalised copy of the stored vector: v_k = V[k-1] / rho[k-1] (what gmres! holds in V[k])
                    AK_TRY(launch_divcopy_dev(c, n, ws->w[0], ws->V[k - 1], ws->rho + (k - 1), stop));
                    pv = ws->w[0];
                }
                if (flexible || precond) {
                    double* tgt = ws->pbuf;
                    if (flexible) {
                        while ((int64_t)ws->Z.size() < k) {
                            double* z = nullptr;
                            AK_TRY(ws_alloc_vec(ws, &z));
                            ws->Z.push_back(z);
                        }
                        tgt = ws->Z[k - 1];
                    }
                    if (precond) AK_TRY(apply_precond_n(ws, prob, u, o, pv, tgt));
                    else AK_TRY(launch_copy(c, n, tgt, pv));
                    pv = tgt;
                }
                // w <- A V[k-1]  (+ fused divcopy of V[k-1], + fused first dot)
                JvpFusion jf;
                jf.stop_flag = stop;
                double* wout = w;
                double* seed = pv;
                if (raw && hosted) {
                    wout = ws->V[k];  // plain tangent of the (preconditioned) normalised seed, straight into basis slot k
                } else if (raw) {
                    // w <- J (V[k-1] / rho[k-1]), straight into basis slot k
                    jf.scale_src = ws->V[k - 1];
                    jf.denom_dev = ws->rho + (k - 1);
                    jf.inv_denom_dev = ws->rinv + (k - 1);
                    jf.raw = true;
                    seed = raw_needs_seed ? ws->pbuf : nullptr;
                    wout = ws->V[k];
                    if (p2p_halo && k > 1 && !(sweep && k == kSwKMax + 1)) {  // ghost rows of V[k-1] were pushed by the neighbours' final pass of iteration k-1
                        const int par = (int)((k - 1) & 1);
                        jf.halo_given = true;
                        jf.halo_lo = nb_down >= 0 ? c->p2p_halo_local(par, 0) : nullptr;
                        jf.halo_hi = nb_up >= 0 ? c->p2p_halo_local(par, 1) : nullptr;
                    }
                } else if (fuse == AK_FUSE_FULL) {
                    if (scale_pending) {
                        jf.scale_src = w;
                        jf.denom_dev = &ws->ctl->Hbis;
                        wi ^= 1;
                        wout = ws->w[wi];
                    }
                    jf.dot_with = ws->V[0];
                    jf.dot_dev = hcol;
                }
                // `blk` Gram-Schmidt steps per sweep over w.  Pass list of one iteration: project on block 0; then
                // every pass subtracts the block the previous pass projected on and projects on the next one
                // (re-orthogonalisation: the blocks are visited a second time, gmres! step 5); the final pass
                // subtracts the last block and measures ||w||^2.  Record r of hcol holds the sums of pass r.
                const int64_t P = pair ? (k + blk - 1) / blk : 0;
                const int64_t npj = (reorth ? 2 : 1) * P;  // projection passes
                auto blk_ptr = [&](int64_t j) -> const double* const* { return ws->V.data() + blk * j; };
                auto blk_len = [&](int64_t j) -> int { return (int)((k - blk * j) < blk ? (k - blk * j) : blk); };
                auto blk_rho = [&](int64_t j) -> const double* { return ws->rho + blk * j; };
                auto blk_gram = [&](int64_t j) -> const double* { return ws->gram + blk * j * kBlkMax; };
                BlockComm pc;
                unsigned long long prev_seq = 0;
                if (pair) {
                    if (p2p) { pc.seq_out = ++c->p2p_seq; prev_seq = pc.seq_out; }
                    if (proj_in_jvp) {
                        // the first projection pass rides in the tangent kernel: <S_b, w> for block 0 while w is in registers
                        jf.proj = blk_ptr(0);
                        jf.nproj = blk_len(0);
                        jf.proj_out = hcol;
                        jf.proj_comm = p2p ? &pc : nullptr;
                    }
                }
                if (lprec) {  // w <- M (A N v_k)
                    AK_TRY(launch_jvp(c, prob, u, seed, ws->qbuf, &jf));
                    AK_TRY(apply_precond(ws, prob, u, o, true, ws->qbuf, wout));
                } else {
                    AK_TRY(launch_jvp(c, prob, u, seed, wout, &jf));
                }
                w = wout;
                // modified Gram-Schmidt
                if (pair) {
                    if (!proj_in_jvp)
                        AK_TRY(launch_mgs_block(c, n, w, nullptr, 0, nullptr, nullptr, nullptr, blk_ptr(0), blk_len(0), 0, hcol,
                                                stop, p2p ? &pc : nullptr));
                    for (int64_t r = 1; r < npj; ++r) {
                        const int64_t js = (r - 1) % P, jp = r % P;  // block subtracted / block projected on
                        if (p2p) {
                            pc.seq_in = prev_seq;
                            pc.seq_out = ++c->p2p_seq;
                            pc.tin_store = hcol + kBlkSums * (r - 1);
                            prev_seq = pc.seq_out;
                        }
                        AK_TRY(launch_mgs_block(c, n, w, blk_ptr(js), blk_len(js), hcol + kBlkSums * (r - 1), blk_gram(js),
                                                blk_rho(js), blk_ptr(jp), blk_len(jp), 0, hcol + kBlkSums * r, stop,
                                                p2p ? &pc : nullptr));
                    }
                    if (p2p) {
                        pc.seq_in = prev_seq;
                        pc.seq_out = ++c->p2p_seq;
                        pc.tin_store = hcol + kBlkSums * (npj - 1);
                        givens_seq = pc.seq_out;
                        if (p2p_halo) {
                            const int par = (int)(k & 1);
                            pc.halo.nx = prob->nx;
                            pc.halo.down_hi = nb_down >= 0 ? c->p2p_halo_of(nb_down, par, 1) : nullptr;
                            pc.halo.up_lo = nb_up >= 0 ? c->p2p_halo_of(nb_up, par, 0) : nullptr;
                        }
                    }
                    AK_TRY(launch_mgs_block(c, n, w, blk_ptr(P - 1), blk_len(P - 1), hcol + kBlkSums * (npj - 1),
                                            blk_gram(P - 1), blk_rho(P - 1), nullptr, 0, 1, hcol + kBlkSums * npj, stop,
                                            p2p ? &pc : nullptr));
                } else if (fuse == AK_FUSE_NONE) {
                    for (int64_t i = 0; i < k; ++i) {
                        AK_TRY(launch_mgs_step(c, n, w, nullptr, nullptr, ws->V[i], 0, hcol + i, stop));   // h = <V_i, w>
                        AK_TRY(launch_mgs_step(c, n, w, ws->V[i], hcol + i, nullptr, 0, nullptr, stop));   // w -= h V_i
                    }
                    if (reorth) {
                        double* h2 = hcol + (k + 1);
                        for (int64_t i = 0; i < k; ++i) {
                            AK_TRY(launch_mgs_step(c, n, w, nullptr, nullptr, ws->V[i], 0, h2 + i, stop));
                            AK_TRY(launch_mgs_step(c, n, w, ws->V[i], h2 + i, nullptr, 0, nullptr, stop));
                        }
                        AK_TRY(launch_mgs_step(c, n, w, nullptr, nullptr, nullptr, 1, h2 + k, stop));
                    } else {
                        AK_TRY(launch_mgs_step(c, n, w, nullptr, nullptr, nullptr, 1, hcol + k, stop));
                    }
                } else {
                    if (fuse != AK_FUSE_FULL)
                        AK_TRY(launch_mgs_step(c, n, w, nullptr, nullptr, ws->V[0], 0, hcol, stop));
                    for (int64_t i = 0; i + 1 < k; ++i)
                        AK_TRY(launch_mgs_step(c, n, w, ws->V[i], hcol + i, ws->V[i + 1], 0, hcol + i + 1, stop));
                    if (reorth) {
                        double* h2 = hcol + (k + 1);
                        // last axpy of sweep 1 fused with first dot of sweep 2
                        AK_TRY(launch_mgs_step(c, n, w, ws->V[k - 1], hcol + k - 1, ws->V[0], 0, h2, stop));
                        for (int64_t i = 0; i + 1 < k; ++i)
                            AK_TRY(launch_mgs_step(c, n, w, ws->V[i], h2 + i, ws->V[i + 1], 0, h2 + i + 1, stop));
                        AK_TRY(launch_mgs_step(c, n, w, ws->V[k - 1], h2 + k - 1, nullptr, 1, h2 + k, stop));
                    } else {
                        AK_TRY(launch_mgs_step(c, n, w, ws->V[k - 1], hcol + k - 1, nullptr, 1, hcol + k, stop));
                    }
                }
                { ProfScope prof(c, PK_SCALAR);
                k_gmres_givens<<<1, 32, 0, sm>>>(ws->ctl, (int)k, nr, ws->R, ws->c, ws->s, ws->z, hcol, reorth, blk,
                                                 want_hist ? ws->hist : nullptr, iter + k, (int)inner_limit,
                                                 &ws->status[slot], p2p ? c->p2p_dev() : P2PDev{}, givens_seq,
                                                 raw ? ws->rho : nullptr, ws->gram, ws->rinv); }
                c->launches++;
                AK_CUDA(cudaGetLastError());
                }  // !sw_it
                AK_CUDA(cudaEventRecord(ws->ev[slot], sm));
                // V[k] <- w / Hbis (skipped on the device when this iteration stopped the pass)
                if (k < inner_limit && !raw) {
                    if (fuse == AK_FUSE_FULL) scale_pending = true;
                    else AK_TRY(launch_divcopy_dev(c, n, ws->V[k], w, &ws->ctl->Hbis, stop));
                }
                if (k >= inner_limit) {
                    // last iteration of the pass: queue the solution update behind it, then drain the window (the
                    // record of iteration inner_limit has stop = tired)
                    AK_TRY(queue_solution_update(k));
                    AK_TRY(verdicts_until(k, &stopped));
                    break;
                }
                // look at an older iteration's verdict while the newer ones run
                AK_TRY(verdicts_until(k - kSpecDepth, &stopped));
                if (stopped) break;
            }
            // hs is the record the pass ended on: the stop verdict, or (basis could not grow) the last complete iteration
            solved = hs.solved != 0;
            breakdown = hs.breakdown != 0;
            rNorm = hs.rNorm;
            K = hs.iter;
            if (K > 0 && !update_queued) AK_TRY(queue_solution_update(k));
        }
        if (K > 0) x_written = true;
        if (peer && *c->p2p_err) break;  // a peer fell out of step: reported below
        inner_itmax -= K;
        iter += K;
        if (iter >= itmax) tired = true;
        if (solved || tired || breakdown) break;
        // FUSE_FULL leaves the last residual candidate in w[wi]; the restart recomputes w anyway
        w = raw ? ws->V[0] : ws->w[wi];
    }
    if (!x_written && !xr_separate) AK_TRY(launch_fill(c, n, x, 0.0));  // x0 = 0 and no iteration ran

    // one drain per solve: the verdict of the back-substitution(s) and the history
    if (hist_host && hist_cap > 0 && ws->hist) {
        int64_t m = iter + 1 < hist_cap ? iter + 1 : hist_cap;
        AK_CUDA(cudaMemcpyAsync(hist_host, ws->hist, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, sm));
    }
    AK_CUDA(cudaStreamSynchronize(sm));
    inconsistent = ws->status[kStatusRing + 1].inconsistent != 0;
    if (peer && *c->p2p_err) {
        set_error("peer-memory collective timed out (a rank fell out of step); the peer path of this context is "
                  "latched off until ak_comm_use_p2p is called again on every rank");
        return AK_ERR_PEER;
    }
    st->niter = iter;
    st->solved = solved ? 1 : 0;
    st->inconsistent = inconsistent ? 1 : 0;
    st->breakdown = breakdown ? 1 : 0;
    st->npass = npass;
    st->rnorm = rNorm;
    st->beta = beta0;
    int flags = 0;
    if (!solved) flags |= AK_FLAG_NOT_SOLVED;
    if (breakdown) flags |= AK_FLAG_BREAKDOWN;
    if (inconsistent) flags |= AK_FLAG_INCONSISTENT;
    return flags;
}

// ---------------------------------------------------------------------------------------
// CG (Krylov.jl cg!, M = I) — every call site in examples/bratu.jl:59-108 uses algo = :cg
// ---------------------------------------------------------------------------------------
static int cg_solve(ak_krylov* ws, const ak_problem* prob, const double* u, const double* b, const ak_krylov_opts* o,
                    ak_krylov_stats* st, double* hist_host, int64_t hist_cap) {
    Ctx* c = ws->ctx;
    const int64_t n = ws->n;
    cudaStream_t sm = c->stream;
    const bool want_hist = (hist_host != nullptr && hist_cap > 0) || o->history;
    const int64_t itmax = o->itmax == 0 ? 2 * ws->n_global : o->itmax;
    if (!ws->hcol) AK_TRY(ws_grow_scalars(ws, 4));
    if (want_hist) AK_TRY(ws_grow_hist(ws, 257));
    // kwarg M of cg! (a symmetric positive definite preconditioner): z = M r, gamma = <r, z>, p = z + beta p, the
    // residual norm is measured in the M-norm like Krylov.jl does (rNorm = sqrt(<r, z>))
    const bool lprec = (o->precond_m != AK_PRECOND_NONE);
    if (lprec && !ws->qbuf) AK_TRY(ws_alloc_vec(ws, &ws->qbuf));
    // small-problem regime: the whole solve in one persistent block (AK_NO_SMALL_CG=1: developer switch for A/B runs)
    static const bool no_small = getenv("AK_NO_SMALL_CG") != nullptr;
    if (!no_small && !lprec && prob->kind == AK_BRATU1D && prob->jvp_mode == AK_JVP_ANALYTIC && c->nranks == 1 && n <= kSmallN &&
        itmax < (1ll << 31)) {
        if (want_hist) AK_TRY(ws_grow_hist(ws, itmax + 2));
        const size_t smem = sizeof(double) * (size_t)(2 * n + 2);
        static bool attr_set = false;
        if (!attr_set) {
            AK_CUDA(cudaFuncSetAttribute(k_cg_small_bratu1d, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(double) * (2 * kSmallN + 2))));
            attr_set = true;
        }
        { ProfScope prof(c, PK_SCALAR);
        k_cg_small_bratu1d<<<1, kSmallThreads, smem, sm>>>((int)n, prob->dx * prob->dx, prob->lambda,
                                                          prob->coef ? prob->coef : u, prob->coef ? 0 : 1, b, ws->x, ws->ctl,
                                                          &ws->status[kStatusRing], want_hist ? ws->hist : nullptr, o->atol,
                                                          o->rtol, (long long)itmax); }
        c->launches++;
        AK_CUDA(cudaGetLastError());
        AK_CUDA(cudaStreamSynchronize(sm));
        KrylovStatus hs0;
        memcpy(&hs0, (const void*)&ws->status[kStatusRing], sizeof(hs0));
        memset(st, 0, sizeof(*st));
        st->niter = hs0.iter;
        st->solved = hs0.solved;
        st->inconsistent = hs0.zerocurv;
        st->npass = 1;
        st->rnorm = hs0.rNorm;
        st->beta = hs0.beta;
        if (hist_host && hist_cap > 0 && ws->hist) {
            int64_t m = hs0.iter + 1 < hist_cap ? hs0.iter + 1 : hist_cap;
            AK_CUDA(cudaMemcpy(hist_host, ws->hist, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost));
        }
        int flags0 = 0;
        if (!hs0.solved) flags0 |= AK_FLAG_NOT_SOLVED;
        if (hs0.zerocurv) flags0 |= AK_FLAG_INCONSISTENT;
        return flags0;
    }
    double *x = ws->x, *r = ws->r, *p = ws->p, *Ap = ws->Ap;
    double* z = lprec ? ws->qbuf : r;  // z === r without a preconditioner
    const int* stop = &ws->ctl->stop;
    AK_TRY(launch_fill(c, n, x, 0.0));
    AK_TRY(launch_copy(c, n, r, b));
    if (lprec) AK_TRY(apply_precond(ws, prob, u, o, true, r, z));
    AK_TRY(launch_copy(c, n, p, z));
    if (lprec) AK_TRY(launch_dot(c, n, r, z, ws->hcol));
    else AK_TRY(launch_sumsq(c, n, r, ws->hcol));
    k_cg_begin<<<1, 32, 0, sm>>>(ws->ctl, ws->hcol, want_hist ? ws->hist : nullptr, o->atol, o->rtol, &ws->status[kStatusRing]);
    c->launches++;
    AK_CUDA(cudaGetLastError());
    AK_CUDA(cudaEventRecord(ws->ev[kStatusRing], sm));
    KrylovStatus hs;
    AK_TRY(wait_status(ws, kStatusRing, &hs));
    memset(st, 0, sizeof(*st));
    st->beta = hs.beta;
    int64_t iter = 0;
    bool solved = hs.solved != 0, zerocurv = false;
    double rNorm = hs.rNorm;
    if (!hs.stop && itmax > 0) {
        int64_t k = 0, next_read = 1;
        while (true) {
            k += 1;
            if (want_hist) AK_TRY(ws_grow_hist(ws, k + 1));
            JvpFusion jf;
            jf.stop_flag = stop;
            jf.dot_with = p;
            jf.dot_dev = ws->hcol + 1;
            AK_TRY(launch_jvp(c, prob, u, p, Ap, &jf));  // Ap = A p, pAp = <p, Ap>
            const int slot = (int)(k % kStatusRing);
            k_cg_alpha<<<1, 32, 0, sm>>>(ws->ctl, ws->hcol + 1, &ws->status[slot]);
            c->launches++;
            // r -= alpha Ap (+ gamma_next = <r, r>); the x update rides with the p update below (same values:
            // x += alpha p uses the p of this iteration, which is only overwritten afterwards)
            if (!lprec) {
                AK_TRY(launch_mgs_step(c, n, r, Ap, &ws->ctl->alpha, nullptr, 1, ws->hcol + 2, stop));
            } else {  // r -= alpha Ap ; z = M r ; gamma_next = <r, z>
                AK_TRY(launch_mgs_step(c, n, r, Ap, &ws->ctl->alpha, nullptr, 0, nullptr, stop));
                AK_TRY(apply_precond(ws, prob, u, o, true, r, z));
                AK_TRY(launch_dot(c, n, r, z, ws->hcol + 2));
            }
            k_cg_beta<<<1, 32, 0, sm>>>(ws->ctl, ws->hcol + 2, (int)k, itmax, want_hist ? ws->hist : nullptr,
                                        &ws->status[slot]);
            c->launches++;
            AK_CUDA(cudaGetLastError());
            AK_CUDA(cudaEventRecord(ws->ev[slot], sm));
            // x += alpha p ; p <- r + beta p   (one pass)
            {
                int64_t blocks = (n + 1023) / 1024;
                if (blocks > (int64_t)c->num_sms * 4) blocks = (int64_t)c->num_sms * 4;
                if (blocks < 1) blocks = 1;
                const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(p) |
                                   reinterpret_cast<uintptr_t>(z)) & 31u) == 0;
                ProfScope prof(c, PK_ELEMENTWISE);
                if (vec) k_cg_update_xp<true><<<(int)blocks, 256, 0, sm>>>(x, p, z, ws->ctl, (int)k, n);
                else k_cg_update_xp<false><<<(int)blocks, 256, 0, sm>>>(x, p, z, ws->ctl, (int)k, n);
                c->launches++;
                AK_CUDA(cudaGetLastError());
            }
            bool stopped = false;
            while (next_read <= k - kSpecDepth && !stopped) {  // verdicts of older iterations, in order
                AK_TRY(wait_status(ws, (int)(next_read % kStatusRing), &hs));
                next_read += 1;
                stopped = hs.stop != 0;
            }
            if (stopped) break;
            if (k >= itmax) {  // drain the window; the record of iteration itmax has stop set
                while (next_read <= k) {
                    AK_TRY(wait_status(ws, (int)(next_read % kStatusRing), &hs));
                    next_read += 1;
                    if (hs.stop) break;
                }
                break;
            }
        }
        // the record we stopped on
        solved = hs.solved != 0;
        zerocurv = hs.zerocurv != 0;
        iter = hs.iter;
        rNorm = hs.rNorm;
    }
    AK_CUDA(cudaStreamSynchronize(sm));
    if (c->p2p_on && c->nranks > 1 && *c->p2p_err) {
        set_error("peer-memory collective timed out (a rank fell out of step); the peer path of this context is "
                  "latched off until ak_comm_use_p2p is called again on every rank");
        return AK_ERR_PEER;
    }
    st->niter = iter;
    st->solved = solved;
    st->inconsistent = zerocurv;
    st->npass = 1;
    st->rnorm = rNorm;
    if (hist_host && hist_cap > 0 && ws->hist) {
        int64_t m = iter + 1 < hist_cap ? iter + 1 : hist_cap;
        AK_CUDA(cudaMemcpy(hist_host, ws->hist, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost));
    }
    int flags = 0;
    if (!solved) flags |= AK_FLAG_NOT_SOLVED;
    if (zerocurv) flags |= AK_FLAG_INCONSISTENT;
    return flags;
}

int krylov_solve_internal(ak_krylov* ws, const ak_problem* p, const double* u, const double* b,
                          const ak_krylov_opts* opts, ak_krylov_stats* st, double* hist_host, int64_t hist_cap) {
    if (ws->algo == AK_ALGO_CG) {
        if (opts->precond_n != AK_PRECOND_NONE) {
            set_error("cg! has no right preconditioner (kwarg N); pass the preconditioner as M");
            return AK_ERR_UNSUPPORTED;
        }
        return cg_solve(ws, p, u, b, opts, st, hist_host, hist_cap);
    }
    return gmres_solve(ws, p, u, b, opts, st, hist_host, hist_cap);
}

}  // namespace ak

using namespace ak;

AK_API void ak_krylov_default_opts(ak_krylov_opts* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->atol = sqrt(2.220446049250313e-16);
    o->rtol = sqrt(2.220446049250313e-16);
    o->fuse = AK_FUSE_SWEEP;  // one sweep per iteration where it applies, the eight-step blocked passes everywhere else
}

AK_API int ak_krylov_create(ak_ctx* ctx, int32_t algo, int64_t n, int32_t memory, int64_t max_basis, ak_krylov** out) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && out && n >= 1, "ak_krylov_create: bad argument");
    AK_REQUIRE(algo == AK_ALGO_GMRES || algo == AK_ALGO_CG || algo == AK_ALGO_FGMRES, "ak_krylov_create: unknown algo");
    AK_REQUIRE(memory >= 1, "ak_krylov_create: memory must be >= 1");
    ak_krylov* ws = new ak_krylov();
    ws->owner = ctx;
    ws->ctx = &ctx->c;
    ws->algo = algo;
    ws->n = n;
    ws->n_global = n;
    if (ctx->c.nranks > 1) {
        // collective: every rank creates its workspace at the same point of the program (like the reference's
        // krylov_workspace call inside newton_krylov!); local sizes may differ by a row / a point between ranks
        Ctx* c = &ctx->c;
        const double mine = (double)n;
        double tot = 0.0;
        int rcg = AK_OK;
        if (cudaMemcpyAsync(c->dscal + 60, &mine, sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rcg = AK_ERR_CUDA;
        if (rcg == AK_OK) rcg = allreduce_sum(c, c->dscal + 60, 1);
        if (rcg == AK_OK && cudaMemcpyAsync(&tot, c->dscal + 60, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rcg = AK_ERR_CUDA;
        if (rcg == AK_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rcg = AK_ERR_CUDA;
        if (rcg != AK_OK) {
            set_error("ak_krylov_create: could not sum the problem size over the ranks");
            delete ws;
            return rcg;
        }
        ws->n_global = (int64_t)(tot + 0.5);
    }
    ws->mem = memory;
    ws->max_basis = max_basis;
    int rc = AK_OK;
    do {
        if ((rc = ws_alloc_vec(ws, &ws->x)) != AK_OK) break;
        if (algo == AK_ALGO_GMRES || algo == AK_ALGO_FGMRES) {
            if ((rc = ws_alloc_vec(ws, &ws->w[0])) != AK_OK) break;
            if ((rc = ws_ensure_basis(ws, memory)) != AK_OK) break;
        } else {
            if ((rc = ws_alloc_vec(ws, &ws->r)) != AK_OK) break;
            if ((rc = ws_alloc_vec(ws, &ws->p)) != AK_OK) break;
            if ((rc = ws_alloc_vec(ws, &ws->Ap)) != AK_OK) break;
        }
        if (pool_alloc(&ctx->c, (void**)&ws->ctl, sizeof(KrylovCtl)) != cudaSuccess) { rc = AK_ERR_NOMEM; break; }
        cudaMemsetAsync(ws->ctl, 0, sizeof(KrylovCtl), ctx->c.stream);
        if (cudaHostAlloc((void**)&ws->status, sizeof(KrylovStatus) * kStatusSlots, cudaHostAllocMapped) != cudaSuccess) {
            rc = AK_ERR_NOMEM;
            break;
        }
        memset(ws->status, 0, sizeof(KrylovStatus) * kStatusSlots);
        for (int i = 0; i < kStatusSlots; ++i)
            if (cudaEventCreateWithFlags(&ws->ev[i], cudaEventDisableTiming) != cudaSuccess) { rc = AK_ERR_CUDA; break; }
        if (rc != AK_OK) break;
        rc = ws_grow_scalars(ws, memory > 4 ? memory : 4);
    } while (0);
    if (rc != AK_OK) {
        if (rc == AK_ERR_NOMEM) set_error("ak_krylov_create: out of device memory (n = %lld, memory = %d)", (long long)n, memory);
        ak_krylov_destroy(ws);
        return rc;
    }
    *out = ws;
    return AK_OK;
}

AK_API int ak_krylov_destroy(ak_krylov* ws) {
    if (!ws) return AK_OK;
    if (ws->inner) ak_krylov_destroy(ws->inner);
    cudaStream_t sm = ws->ctx ? ws->ctx->stream : nullptr;
    if (sm) cudaStreamSynchronize(sm);
    auto rel = [&](void* p) { if (p) cudaFreeAsync(p, sm); };
    for (double* p : ws->chunks) rel(p);
    rel((void*)ws->V_dev);
    rel(ws->R); rel(ws->c); rel(ws->s); rel(ws->z); rel(ws->hcol); rel(ws->hist); rel(ws->rho); rel(ws->gram);
    rel(ws->ycoef); rel(ws->rinv); rel(ws->sw_sums);
    rel(ws->ctl);
    if (ws->status) cudaFreeHost(ws->status);
    for (int i = 0; i < kStatusSlots; ++i)
        if (ws->ev[i]) cudaEventDestroy(ws->ev[i]);
    delete ws;
    return AK_OK;
}

AK_API int ak_krylov_solve(ak_krylov* ws, const ak_problem* p, const double* u, const double* b,
                           const ak_krylov_opts* opts, ak_krylov_stats* stats_out, double* hist_host,
                           int64_t hist_cap) {
    AK_REQUIRE(ws && p && b && opts && stats_out, "ak_krylov_solve: NULL argument");
    AK_ENTER(ws->owner);
    AK_REQUIRE(ak_problem_size(p) == ws->n, "ak_krylov_solve: problem size does not match the workspace");
    return krylov_solve_internal(ws, p, u, b, opts, stats_out, hist_host, hist_cap);
}

AK_API double* ak_krylov_x(ak_krylov* ws) { return ws ? ws->x : nullptr; }

AK_API int ak_krylov_basis(ak_krylov* ws, int64_t i, double** stored_dev, double* scale_host, int64_t* count_out) {
    AK_REQUIRE(ws, "ak_krylov_basis: NULL workspace");
    AK_ENTER(ws->owner);
    if (count_out) *count_out = (int64_t)ws->V.size();
    if (stored_dev == nullptr && scale_host == nullptr) return AK_OK;
    AK_REQUIRE(i >= 0 && i < (int64_t)ws->V.size(), "ak_krylov_basis: index out of range");
    if (stored_dev) *stored_dev = ws->V[(size_t)i];
    if (scale_host) {
        *scale_host = 1.0;
        if (ws->last_raw && ws->rho && i <= ws->kcap) {
            AK_CUDA(cudaMemcpyAsync(scale_host, ws->rho + i, sizeof(double), cudaMemcpyDeviceToHost, ws->ctx->stream));
            AK_CUDA(cudaStreamSynchronize(ws->ctx->stream));
        }
    }
    return AK_OK;
}

AK_API int ak_precond_apply(ak_ctx* ctx, const ak_problem* p, const double* u, int32_t kind, int32_t itmax,
                            const double* x, double* y) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && p && x && y, "ak_precond_apply: NULL argument");
    if (kind != AK_PRECOND_INNER_GMRES) return precond_apply(&ctx->c, p, u, kind, nullptr, nullptr, x, y);
    // mul!(y, P::GmresPreconditioner, x) = copyto!(y, gmres(P.J, x; P.itmax)[1])   examples/bratu.jl:146-149
    ak_krylov* ws = nullptr;
    AK_TRY(ak_krylov_create(ctx, AK_ALGO_GMRES, ak_problem_size(p), 20, 0, &ws));
    ak_krylov_opts io;
    ak_krylov_default_opts(&io);
    io.itmax = itmax;
    ak_krylov_stats st;
    int rc = krylov_solve_internal(ws, p, u, x, &io, &st, nullptr, 0);
    if (rc >= 0) rc = launch_copy(&ctx->c, ws->n, y, ws->x);
    cudaStreamSynchronize(ctx->c.stream);
    ak_krylov_destroy(ws);
    return rc < 0 ? rc : AK_OK;
}
