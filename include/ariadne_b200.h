/*
 * ariadne_b200.h — C ABI of libariadne_b200.so
 *
 * B200-native (sm_100a) Jacobian-free Newton–Krylov inner loop behind the API of
 * vchuravy/NewtonKrylov.jl ("Ariadne").  Every entry point below is what a Julia
 * `ccall` (or Python ctypes / C++) binding for that path binds; the comment on each
 * names the reference interface it replaces (file:line under the reference tree).
 *
 * Conventions
 *   - plain C types only: pointers, sizes, PODs.  `double*` arguments are DEVICE
 *     pointers unless the name ends in `_host`.
 *   - every function returns an int status: 0 = ok, <0 = CUDA/NCCL/usage error
 *     (text via ak_last_error()), >0 = numerical flag (AK_FLAG_*).
 *   - one host thread per context; all work is enqueued on the context's stream;
 *     a call blocks only when it returns a host scalar.
 *   - there is NO CPU fallback: without a CUDA device ak_ctx_create fails.
 *
 * Vector layout in HBM ("compact slab"): a grid function on nx × ny points is
 * nx*ny contiguous doubles, x fastest, WITHOUT ghost cells.  Dirichlet ghosts are
 * implicit zeros inside the stencil kernels, periodic ghosts are index wraps, and
 * inter-GPU ghosts are two nx-long halo rows owned by the context.  The reference's
 * (N+2)×(M+2) OffsetArray behind `HaloVector` (examples/halovector.jl:3-45,
 * examples/heat_2D.jl:76,90) maps onto this with ak_halo_pack / ak_halo_unpack.
 */
#ifndef ARIADNE_B200_H
#define ARIADNE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AK_ABI_VERSION 4

/* ---- status / flags ------------------------------------------------------- */
enum {
    AK_OK = 0,
    AK_ERR_CUDA = -1,
    AK_ERR_ARG = -2,
    AK_ERR_NCCL = -3,
    AK_ERR_NOMEM = -4,
    AK_ERR_UNSUPPORTED = -5,
    AK_ERR_USER = -6,          /* a caller-supplied callback returned non-zero */
    AK_ERR_PEER = -7           /* a peer-memory (NVLink mailbox) wait timed out: a rank fell out of step.  The peer path
                                  of the context stays latched off (every solve returns this code) until
                                  ak_comm_use_p2p is called again on every rank                                        */
};
/* positive numerical flags (bit set) */
enum {
    AK_FLAG_NOT_SOLVED = 1,    /* itmax reached before tolerance ("tired") */
    AK_FLAG_BREAKDOWN = 2,     /* Hbis <= eps^(3/4) */
    AK_FLAG_INCONSISTENT = 4,  /* singular R in back-substitution */
    AK_FLAG_NAN = 8            /* residual norm is Inf/NaN: src/Ariadne.jl:353-356 */
};

/* ---- problem description: the residual callback `F!(res,u,p)` -------------- */
/* kinds of residual (reference file:line each one restates) */
enum {
    AK_SIMPLE2 = 0,   /* 2x2 system, test/runtests.jl:4-7, examples/simple.jl:6-9 */
    AK_BRATU1D = 1,   /* examples/bratu.jl:14-24 */
    AK_BRATU2D = 2,   /* tensor extension of bratu.jl on heat_2D's grid (defined here) */
    AK_HEAT1D = 3,    /* examples/heat_1D.jl:12-25 (+ bc! :34-37, periodic_bc! :39-42) */
    AK_HEAT2D = 4,    /* examples/heat_2D.jl:45-62 (+ bc_zero! :28-38, bc_periodic! :15-26) */
    AK_HEAT1D_DG = 5, /* examples/heat_1D_DG.jl:17-36, DG/SBP polydeg 3, periodic */
    AK_USER = 6       /* caller-supplied F!(res,u,p) (src/Ariadne.jl:250-256) and, optionally, its tangent:
                         the generic seam of newton_krylov!(F!, u, p, res), e.g. examples/bvp.jl:10-23 */
};
enum { AK_BC_ZERO = 0, AK_BC_PERIODIC = 1 };
/* time discretisation wrapper around the RHS f!: examples/implicit.jl */
enum {
    AK_STEADY = 0,     /* F(u) = f(u)                      (Bratu, simple)      */
    AK_EULER = 1,      /* G_Euler!     implicit.jl:8-13                         */
    AK_MIDPOINT = 2,   /* G_Midpoint!  implicit.jl:17-25 (alpha = 0.5)          */
    AK_TRAPEZOID = 3   /* G_Trapezoid! implicit.jl:29-37                        */
};
enum {
    AK_JVP_ANALYTIC = 0, /* exact tangent-linear stencil == what Enzyme forward mode
                            yields at src/Ariadne.jl:48-57 (parity path)          */
    AK_JVP_FD_FUSED = 1, /* (F(u+eps v)-F(u))/eps evaluated point-wise in one pass,
                            u+eps v never materialised; the Bratu problems (1-D, 2-D) */
    AK_JVP_FD = 2        /* generic two-evaluation finite difference (F(u+eps v)-F(u))/eps through the
                            residual itself; eps = fd_eps, or sqrt(eps_mach)(1+||u||)/||v|| when 0.
                            Default of AK_USER problems without a tangent callback               */
};

/* AK_USER callbacks.  `stream` is the context's cudaStream_t: the callback ENQUEUES device work on it
 * (or synchronises it before touching the data any other way) and returns 0; all pointers are device
 * pointers to ak_problem_size(p) doubles.  Non-zero return aborts the solve with AK_ERR_USER.       */
/* res <- F(u):  F!(res, u, p), src/Ariadne.jl:252,302,349.  May mutate u like the reference's BC code. */
typedef int (*ak_user_residual_fn)(void* user, uint64_t stream, double* u, double* res);
/* out <- J(u) v: what `mul!(out, J::JacobianOperator, v)` (src/Ariadne.jl:48-57) computes by forward-mode AD */
typedef int (*ak_user_jvp_fn)(void* user, uint64_t stream, const double* u, double* v, double* out);

typedef struct ak_problem {
    int32_t kind;      /* AK_BRATU1D ...                                           */
    int32_t bc;        /* AK_BC_*                                                  */
    int32_t scheme;    /* AK_STEADY / AK_EULER / ...                               */
    int32_t jvp_mode;  /* AK_JVP_*                                                 */
    int64_t nx;        /* points along x (fast axis); 1-D problems: total length;
                          DG: number of elements * 4                               */
    int64_t ny;        /* LOCAL rows on this rank (1 for 1-D problems)             */
    int64_t gny;       /* GLOBAL rows (== ny on one GPU)                           */
    int64_t gy0;       /* global index of local row 0                              */
    double dx, dy;     /* grid spacings (DG: element width h)                      */
    double lambda;     /* Bratu parameter                                          */
    double a;          /* diffusivity                                              */
    double dt;         /* time step for scheme != AK_STEADY                        */
    double fd_eps;     /* epsilon for AK_JVP_FD_FUSED (0 => sqrt(eps_mach))        */
    const double* un;  /* device: previous time level u_n (scheme != AK_STEADY)    */
    double* coef;      /* device scratch, n doubles, or NULL.  Bratu: ak_residual
                          stores lambda*exp(u) here so that every JVP of the same
                          Newton step is a pure 24n-byte stencil.  AK_JVP_FD: ak_residual
                          keeps a copy of F(u) here so that a JVP costs one residual
                          evaluation instead of two                                */
    double* work;      /* device scratch, n doubles, or NULL (midpoint/trapezoid)  */
    /* AK_USER only (n = nx unknowns on this rank; reductions stay global over the communicator) */
    ak_user_residual_fn user_residual;
    ak_user_jvp_fn user_jvp;     /* NULL => AK_JVP_FD                                        */
    void* user_data;             /* the closure `p` of F!(res, u, p)                         */
} ak_problem;

/* number of LOCAL unknowns of a problem (length of u, res, v on this rank) */
int64_t ak_problem_size(const ak_problem* p);

/* ---- context --------------------------------------------------------------- */
typedef struct ak_ctx ak_ctx;

int ak_abi_version(void);
const char* ak_last_error(void);
int ak_ctx_create(int device, ak_ctx** out);
int ak_ctx_destroy(ak_ctx* ctx);
int ak_ctx_sync(ak_ctx* ctx);
/* cudaStream_t of the context as an integer (for torch.cuda.ExternalStream / events) */
uint64_t ak_ctx_stream(ak_ctx* ctx);
/* number of kernels this context has launched since creation / since last reset */
int64_t ak_ctx_launch_count(ak_ctx* ctx, int reset);
/* CUDA-event stopwatch on the context's stream (bench.py uses it for per-kernel times) */
int ak_timer_start(ak_ctx* ctx);
int ak_timer_stop(ak_ctx* ctx, double* ms_out);

/* In-stream profiler: when enabled, every kernel launch is bracketed by CUDA events on the
 * context's stream; ak_profile_read returns the number of launches of one kernel class and the
 * sum of their device times.  Classes: 0 axpy+dot (fused MGS step, 32n bytes), 1 axpy+norm (24n),
 * 2 axpy (24n), 3 dot (16n), 4 sum of squares (8n), 5 JVP, 6 residual, 7 element-wise,
 * 8 basis combine, 9 one-thread scalar kernels, 10 full pass of the blocked Gram-Schmidt sweep (48n for
 * blocks of 2, 80n for blocks of 4, 144n for blocks of 8), 11 first / ragged passes of the blocked sweep, 12 final
 * pass of the blocked sweep (last update + norm + new Gram entries).  Enabling resets the counters.    */
int ak_profile_enable(ak_ctx* ctx, int on);
int ak_profile_read(ak_ctx* ctx, int kernel_class, int64_t* count_out, double* ms_total_out);

/* device memory owned by the library; the library never frees caller memory */
int ak_malloc(ak_ctx* ctx, int64_t n_doubles, double** out);
int ak_free(ak_ctx* ctx, double* p);
int ak_upload(ak_ctx* ctx, double* dst_dev, const double* src_host, int64_t n);   /* blocking */
int ak_download(ak_ctx* ctx, double* dst_host, const double* src_dev, int64_t n); /* blocking */
/* pinned host staging (for the e2e path) */
int ak_host_alloc(int64_t n_doubles, double** out);
int ak_host_free(double* p);

/* HaloVector layout bridge: examples/halovector.jl:3-45, heat_2D.jl:76,90.
 * `padded` is the reference's (nx+2) x (ny+2) column-major array (first index
 * fastest, ghost ring included), `compact` is this library's nx*ny slab.       */
int ak_halo_pack(ak_ctx* ctx, double* compact, const double* padded, int64_t nx, int64_t ny);
int ak_halo_unpack(ak_ctx* ctx, double* padded, const double* compact, int64_t nx, int64_t ny,
                   int32_t bc /* ghost ring filled like bc_zero!/bc_periodic! */);

/* ---- multi-GPU (one process per GPU; examples/halovector.jl is the model) --- */
/* The caller's launcher (torch.distributed / MPI / Julia Distributed) moves the
 * 128-byte NCCL id from rank 0 to everybody; nothing else crosses processes on
 * the host.  After ak_comm_init every reduction below is global and every 2-D
 * stencil exchanges one halo row with rank-1 / rank+1 (slab decomposition along y). */
int ak_comm_unique_id(char id_out[128]);
int ak_comm_init(ak_ctx* ctx, int nranks, int rank, const char id[128]);
/* NVLink peer memory for the fused compute + collective kernels of the pair-wise GMRES sweep:
 * every rank exports one block through CUDA IPC (handles travel in an NCCL all-gather); the
 * Gram-Schmidt kernels then deposit their partial sums directly in the peers' mailboxes and push
 * their boundary rows into the neighbours' ghost rows (st.global over NVLink), so a restart cycle
 * issues no NCCL call.  `halo_doubles` >= nx of the widest 2-D problem.  Collective call.     */
int ak_comm_enable_p2p(ak_ctx* ctx, int64_t halo_doubles);
int ak_comm_p2p_enabled(ak_ctx* ctx);
/* switch between the peer-memory path and the NCCL path (must be done on all ranks alike) */
int ak_comm_use_p2p(ak_ctx* ctx, int on);
int ak_comm_rank(ak_ctx* ctx, int* rank, int* nranks);
int ak_comm_barrier(ak_ctx* ctx);
/* ---- residual callback  F!(res,u,p): src/Ariadne.jl:250-256,302,349 --------- */
/* res <- F(u).  If nrm_out_host != NULL also returns ||res||_2 (the `norm(res)` of
 * src/Ariadne.jl:303,350), reduced in the same pass.  `u` is non-const because the
 * reference's boundary code mutates it (heat_1D.jl:16 bc!(u)).                  */
int ak_residual(ak_ctx* ctx, const ak_problem* p, double* u, double* res, double* nrm_out_host);

/* ---- operator protocol: JacobianOperator, src/Ariadne.jl:34-57 -------------- */
/* out <- J(u) v.  Exact tangent (AK_JVP_ANALYTIC).  `v` is non-const: forward mode
 * through bc!(u) zeroes v's boundary entries (heat_1D.jl:16,34-37).             */
int ak_jvp(ak_ctx* ctx, const ak_problem* p, const double* u, double* v, double* out);
/* out <- J(u)^T v  (src/Ariadne.jl:93-107).  Implemented for the operators that are symmetric in this
 * layout (Bratu 1-D/2-D, heat 2-D, heat 1-D with bc!), the 2x2 test system, the DG operator (J^T = W J W^-1
 * by upwind-SBP duality) and 1-D heat with periodic_bc! (Euler / Trapezoid, one GPU: J = (c1 a L - I) B with the
 * boundary copy B, adjoint kernel B^T (c1 a L^T - I)); AK_USER returns AK_ERR_UNSUPPORTED (GMRES/CG never need J^T). */
int ak_jvp_transpose(ak_ctx* ctx, const ak_problem* p, const double* u, double* v, double* out);
/* Out[:, c] <- J(u) V[:, c], c < ncols: the batched `mul!(Out, J, V)` of src/Ariadne.jl:69-83 for column-major
 * matrices with leading dimensions ldv, ldo >= n.  Bratu 1-D/2-D (the operators that depend on u): ONE multi-RHS
 * launch, lambda e^u read (or computed) once per group of four columns, 18n instead of 24n bytes per column; the u-independent
 * heat / DG tangents and AK_USER: one tangent sweep per column (V's boundary entries may be overwritten like in
 * ak_jvp).  Every column equals ak_jvp's result bit for bit.                                                  */
int ak_jvp_batched(ak_ctx* ctx, const ak_problem* p, const double* u, double* V, int64_t ldv, double* Out,
                   int64_t ldo, int32_t ncols);

/* ---- vector-kernel protocol: Krylov.k* hooks, examples/halovector.jl:51-147 -- */
int ak_dot(ak_ctx* ctx, int64_t n, const double* x, const double* y, double* out_host);  /* kdot  :51-62  */
int ak_nrm2(ak_ctx* ctx, int64_t n, const double* x, double* out_host);                  /* knorm :64-74  */
int ak_scal(ak_ctx* ctx, int64_t n, double s, double* x);                                /* kscal! :76-85 */
int ak_axpy(ak_ctx* ctx, int64_t n, double s, const double* x, double* y);               /* kaxpy! :87-97 */
int ak_axpby(ak_ctx* ctx, int64_t n, double s, const double* x, double t, double* y);    /* kaxpby! :99-109 */
int ak_copy(ak_ctx* ctx, int64_t n, double* y, const double* x);                         /* kcopy! :111-121 */
int ak_fill(ak_ctx* ctx, int64_t n, double* x, double val);                              /* kfill! :123-132 */
int ak_ref(ak_ctx* ctx, int64_t n, double* x, double* y, double c, double s);            /* kref!  :134-147 */
/* y <- x / s  (Krylov.jl kdivcopy!, used for V[k+1] = q / Hbis) */
int ak_divcopy(ak_ctx* ctx, int64_t n, double* y, const double* x, double s);

/* ---- whole linear solve: krylov_workspace / krylov_solve!, src/Ariadne.jl:317-318,338-340 */
typedef struct ak_krylov ak_krylov;

enum { AK_ALGO_GMRES = 0, AK_ALGO_CG = 1, AK_ALGO_FGMRES = 2 /* Krylov.jl fgmres!: examples/bratu.jl:131-157 */ };
/* Preconditioners of `newton_krylov!(...; M = (J) -> ..., N = (J) -> ...)` (src/Ariadne.jl:296-297,324-329): the
 * reference calls M(J) / N(J) once per Newton step and hands the objects to Krylov.jl, which applies them with
 * mul! (ldiv = false) or ldiv! (ldiv = true).  Here the object is named by kind and built natively from the
 * operator J = (problem, u); it is rebuilt implicitly for every u, like N(J) in the reference.            */
enum {
    AK_PRECOND_NONE = 0,
    AK_PRECOND_INNER_GMRES = 1, /* (J) -> GmresPreconditioner(J, itmax): y = gmres(J, x; itmax)
                                   (examples/bratu.jl:141-157, bvp.jl:29-38)                     */
    AK_PRECOND_USER = 2,        /* caller-supplied y <- P x (mul!) or y <- P \ x (ldiv!): ak_precond_apply_fn */
    AK_PRECOND_JACOBI = 3,      /* y = x ./ diag(J(u)); Bratu 1-D/2-D, heat 1-D/2-D (Euler, Trapezoid)     */
    AK_PRECOND_TRIDIAG_LU = 4   /* y = J(u) \ x by tridiagonal LU (parallel partitioned Thomas): what
                                   `ilu(collect(J))` with ldiv = true is for the 1-D Bratu Jacobian, whose LU
                                   factors have no fill-in (examples/bratu.jl:121-139).  AK_BRATU1D, one GPU */
};
/* y <- P x on `stream` (device pointers, n doubles); non-zero return aborts with AK_ERR_USER */
typedef int (*ak_precond_apply_fn)(void* user, uint64_t stream, const double* x, double* y);
/* how aggressively the Arnoldi step is fused (all levels keep modified Gram-Schmidt
 * order, so they differ only by rounding of identical operations)               */
enum {
    AK_FUSE_NONE = 0,  /* reference op list: dot, axpy, ..., nrm2, divcopy        */
    AK_FUSE_MGS = 1,   /* axpy_i + dot_{i+1} in one pass, last axpy + nrm2 fused  */
    AK_FUSE_FULL = 2,  /* + (divcopy + JVP + first dot) in one pass               */
    AK_FUSE_PAIR = 3,  /* two Gram-Schmidt steps per sweep over w: the pass that subtracts
                          h_a v_a + h_b v_b also accumulates <y_a,w>, <y_b,w>, <y_b,y_a>; the
                          coefficient of y_b follows as <y_b,w> - <y_a,w><y_b,y_a>, which is
                          algebraically the modified Gram-Schmidt value (24n bytes per step
                          instead of 32n); falls back to FULL with reorthogonalization    */
    AK_FUSE_BLOCK4 = 4,/* same with four steps per sweep: h_b = <y_b,w> - sum_{a<b} h_a <y_b,y_a>
                          (20n bytes per step); the Gram entries <y_b,y_a> of a block are measured once,
                          by the final pass of the iteration that finishes y_b, and cached          */
    AK_FUSE_BLOCK8 = 5,/* eight steps per sweep (18n bytes per step)                                */
    AK_FUSE_SWEEP = 6  /* ONE pass over the basis per iteration (2-D problems): the Gram-Schmidt update of iteration k,
                        * ||w||, the tangent J w of iteration k+1 and all its projections in the same kernel;
                        * coefficients by forward substitution with the cached Gram matrix (8n(k+4) bytes per
                        * iteration; iterations beyond 24 per pass and all other problems run as BLOCK8)       */
    /* PAIR and BLOCK4 keep the Krylov basis UN-NORMALISED: iteration k works in place on basis slot k, whose
     * finished content is the stored vector rho_k v_k (rho_k = Hbis); gmres!'s `V[k+1] = w / Hbis` is never
     * materialised, the scales enter the Gram-Schmidt coefficients and the JVP divides its result by rho_k
     * (J is linear).  Same algebra, rounding differs in the last bits.                                     */
};

typedef struct ak_krylov_opts {
    double atol;            /* Krylov.jl default sqrt(eps)                         */
    double rtol;            /* Krylov.jl default sqrt(eps); Newton passes eta      */
    int64_t itmax;          /* 0 => 2n (Krylov.jl)                                 */
    int32_t restart;        /* Krylov.jl restart=false default                     */
    int32_t reorthogonalization; /* second MGS sweep (heat_2D.jl:131)              */
    int32_t history;        /* record rNorm per iteration                          */
    int32_t fuse;           /* AK_FUSE_*                                           */
    int32_t precond_n;      /* AK_PRECOND_*: right preconditioner (kwarg N)        */
    int32_t precond_itmax;  /* itmax of the inner GMRES (GmresPreconditioner.itmax) */
    int32_t precond_m;      /* AK_PRECOND_*: left preconditioner (kwarg M); GMRES/FGMRES then
                               iterate on M J N and measure ||M r||, like Krylov.jl          */
    int32_t precond_m_itmax;
    ak_precond_apply_fn n_apply;  /* AK_PRECOND_USER */
    void* n_user;
    ak_precond_apply_fn m_apply;
    void* m_user;
} ak_krylov_opts;

typedef struct ak_krylov_stats {
    int64_t niter;          /* workspace.stats.niter: src/Ariadne.jl:363,367       */
    int32_t solved;
    int32_t inconsistent;
    int32_t breakdown;
    int32_t npass;
    double rnorm;           /* last recurrence residual norm                       */
    double beta;            /* ||b||                                               */
} ak_krylov_stats;

/* Krylov.jl keyword defaults (atol = rtol = sqrt(eps), itmax = 0, no restart, no preconditioners) and
 * fuse = AK_FUSE_SWEEP, the fastest level (falls back by itself to BLOCK8 where it does not apply).        */
void ak_krylov_default_opts(ak_krylov_opts* o);
/* krylov_workspace(algo, KrylovConstructor(res)): `memory` = 20 in Krylov.jl;
 * max_basis caps how far a non-restarted basis may grow (0 => as HBM allows)     */
int ak_krylov_create(ak_ctx* ctx, int32_t algo, int64_t n, int32_t memory, int64_t max_basis,
                     ak_krylov** out);
int ak_krylov_destroy(ak_krylov* ws);
/* krylov_solve!(ws, J, b; kw...) with J = JacobianOperator(F!, res, u, p):
 * solves J(u) x = b, x0 = 0.  hist_host (may be NULL) receives up to hist_cap
 * recurrence residual norms (index 0 = beta).                                    */
int ak_krylov_solve(ak_krylov* ws, const ak_problem* p, const double* u, const double* b,
                    const ak_krylov_opts* opts, ak_krylov_stats* stats_out,
                    double* hist_host, int64_t hist_cap);
/* workspace.x: device pointer to the solution of the last solve */
double* ak_krylov_x(ak_krylov* ws);
/* Inspection of the Krylov basis after a GMRES solve (diagnostics; the loss-of-orthogonality tests): the STORED vector
 * of basis vector i (device pointer, n doubles) and its scale, v_i = stored / scale.  scale = 1 when the basis is kept
 * normalised like Krylov.jl's V[i]; the blocked sweeps keep rho_i v_i.  count_out: vectors the workspace holds.
 * stored_dev / scale_host / count_out may each be NULL.                                                            */
int ak_krylov_basis(ak_krylov* ws, int64_t i, double** stored_dev, double* scale_host, int64_t* count_out);
/* One application of a native preconditioner object built from J = (p, u): `mul!(y, P, x)` for
 * AK_PRECOND_INNER_GMRES / AK_PRECOND_JACOBI, `ldiv!(y, P, x)` for AK_PRECOND_TRIDIAG_LU — what Krylov.jl calls
 * on the object returned by `N(J)` / `M(J)` (src/Ariadne.jl:324-329) when Krylov.jl itself drives the solve. */
int ak_precond_apply(ak_ctx* ctx, const ak_problem* p, const double* u, int32_t kind, int32_t itmax,
                     const double* x, double* y);

/* ---- whole Newton solve: newton_krylov!, src/Ariadne.jl:288-372 ------------- */
enum { AK_FORCING_NONE = 0, AK_FORCING_FIXED = 1, AK_FORCING_EW = 2 };

typedef struct ak_newton_opts {
    double tol_rel;        /* 1e-6   src/Ariadne.jl:290 */
    double tol_abs;        /* 1e-12  :291               */
    int32_t max_niter;     /* 50     :292               */
    int32_t forcing;       /* AK_FORCING_EW  :293       */
    double eta;            /* Fixed.eta = 0.1 :186      */
    double eta_max;        /* 0.999 :198                */
    double gamma;          /* 0.9   :199                */
    int32_t algo;          /* AK_ALGO_GMRES :295        */
    int32_t memory;        /* 20                         */
    int64_t max_basis;     /* growth cap, 0 = auto       */
    ak_krylov_opts krylov; /* krylov_kwargs :298 (rtol is overridden by eta when forcing != NONE,
                              unless krylov_rtol_override != 0: "later keys win", :323-333) */
    int32_t krylov_rtol_override;
    int32_t verbose;
} ak_newton_opts;

typedef struct ak_newton_stats {
    int32_t solved;             /* n_res <= tol   :371 */
    int32_t outer_iterations;   /* Stats :265-269      */
    int64_t inner_iterations;
    double n_res;
    double tol;
    double t_seconds;           /* :301,370            */
    int32_t flags;              /* AK_FLAG_*           */
} ak_newton_stats;

/* callback(u, res, n_res) of src/Ariadne.jl:304,351 — device pointers */
typedef void (*ak_newton_callback)(void* user, const double* u_dev, const double* res_dev,
                                   double n_res);

void ak_newton_default_opts(ak_newton_opts* o);
/* hist_* (may be NULL): per Newton iteration ||F|| (index 0 = initial), GMRES niter, eta used */
int ak_newton_solve(ak_ctx* ctx, const ak_problem* p, double* u, double* res,
                    const ak_newton_opts* opts, ak_newton_stats* stats_out,
                    double* hist_nres_host, int64_t* hist_inner_host, double* hist_eta_host,
                    int32_t hist_cap, ak_newton_callback cb, void* cb_user);
/* Same, but u lives in HOST memory: copies u host->device, solves, copies u back
 * (what a Julia caller holding an Array{Float64} does; bench.py's e2e leg).
 * `un_host` is the previous time level for scheme != AK_STEADY (or NULL).        */
int ak_newton_solve_host(ak_ctx* ctx, const ak_problem* p, double* u_host, const double* un_host,
                         const ak_newton_opts* opts, ak_newton_stats* stats_out,
                         double* hist_nres_host, int64_t* hist_inner_host, int32_t hist_cap);

/* Eisenstat-Walker forcing update, src/Ariadne.jl:207-216 (host scalar code) */
double ak_forcing_ew(double eta_max, double gamma, double eta, double tol, double n_res,
                     double n_res_prior);

/* ---- implicit time stepper: solve(G!, f!, u_n, p, dt, ts), examples/implicit.jl:54-78 */
/* advances un_dev by nsteps steps of p->scheme with tol_abs = 6e-6 (implicit.jl:69);
 * the Krylov workspace is reused across steps.  per_step_* (may be NULL) get one
 * entry per step.                                                               */
int ak_implicit_solve(ak_ctx* ctx, ak_problem* p, double* un_dev, int32_t nsteps,
                      const ak_newton_opts* opts, int32_t* per_step_newton_host,
                      int64_t* per_step_inner_host, int32_t* per_step_solved_host);

#ifdef __cplusplus
}
#endif
#endif /* ARIADNE_B200_H */
