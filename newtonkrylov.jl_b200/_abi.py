"""ctypes mirror of include/ariadne_b200.h (structs and enums only; loads nothing).

Kept separate from `_lib.py` so that test infrastructure (oracle/) can share the
struct layouts without touching the CUDA library.
"""
import ctypes as C

ABI_VERSION = 4

# status / flags
AK_OK = 0
AK_ERR_CUDA, AK_ERR_ARG, AK_ERR_NCCL, AK_ERR_NOMEM, AK_ERR_UNSUPPORTED, AK_ERR_USER, AK_ERR_PEER = -1, -2, -3, -4, -5, -6, -7
AK_FLAG_NOT_SOLVED, AK_FLAG_BREAKDOWN, AK_FLAG_INCONSISTENT, AK_FLAG_NAN = 1, 2, 4, 8

# problem kinds
AK_SIMPLE2, AK_BRATU1D, AK_BRATU2D, AK_HEAT1D, AK_HEAT2D, AK_HEAT1D_DG, AK_USER = range(7)
AK_BC_ZERO, AK_BC_PERIODIC = 0, 1
AK_STEADY, AK_EULER, AK_MIDPOINT, AK_TRAPEZOID = range(4)
AK_JVP_ANALYTIC, AK_JVP_FD_FUSED, AK_JVP_FD = 0, 1, 2
AK_ALGO_GMRES, AK_ALGO_CG, AK_ALGO_FGMRES = 0, 1, 2
AK_PRECOND_NONE, AK_PRECOND_INNER_GMRES, AK_PRECOND_USER, AK_PRECOND_JACOBI, AK_PRECOND_TRIDIAG_LU = 0, 1, 2, 3, 4
AK_FUSE_NONE, AK_FUSE_MGS, AK_FUSE_FULL, AK_FUSE_PAIR, AK_FUSE_BLOCK4, AK_FUSE_BLOCK8, AK_FUSE_SWEEP = 0, 1, 2, 3, 4, 5, 6
AK_FORCING_NONE, AK_FORCING_FIXED, AK_FORCING_EW = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)


class ak_problem(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("bc", C.c_int32),
        ("scheme", C.c_int32),
        ("jvp_mode", C.c_int32),
        ("nx", C.c_int64),
        ("ny", C.c_int64),
        ("gny", C.c_int64),
        ("gy0", C.c_int64),
        ("dx", C.c_double),
        ("dy", C.c_double),
        ("lambda_", C.c_double),
        ("a", C.c_double),
        ("dt", C.c_double),
        ("fd_eps", C.c_double),
        ("un", C.c_void_p),
        ("coef", C.c_void_p),
        ("work", C.c_void_p),
        ("user_residual", C.c_void_p),
        ("user_jvp", C.c_void_p),
        ("user_data", C.c_void_p),
    ]


class ak_krylov_opts(C.Structure):
    _fields_ = [
        ("atol", C.c_double),
        ("rtol", C.c_double),
        ("itmax", C.c_int64),
        ("restart", C.c_int32),
        ("reorthogonalization", C.c_int32),
        ("history", C.c_int32),
        ("fuse", C.c_int32),
        ("precond_n", C.c_int32),
        ("precond_itmax", C.c_int32),
        ("precond_m", C.c_int32),
        ("precond_m_itmax", C.c_int32),
        ("n_apply", C.c_void_p),
        ("n_user", C.c_void_p),
        ("m_apply", C.c_void_p),
        ("m_user", C.c_void_p),
    ]


class ak_krylov_stats(C.Structure):
    _fields_ = [
        ("niter", C.c_int64),
        ("solved", C.c_int32),
        ("inconsistent", C.c_int32),
        ("breakdown", C.c_int32),
        ("npass", C.c_int32),
        ("rnorm", C.c_double),
        ("beta", C.c_double),
    ]


class ak_newton_opts(C.Structure):
    _fields_ = [
        ("tol_rel", C.c_double),
        ("tol_abs", C.c_double),
        ("max_niter", C.c_int32),
        ("forcing", C.c_int32),
        ("eta", C.c_double),
        ("eta_max", C.c_double),
        ("gamma", C.c_double),
        ("algo", C.c_int32),
        ("memory", C.c_int32),
        ("max_basis", C.c_int64),
        ("krylov", ak_krylov_opts),
        ("krylov_rtol_override", C.c_int32),
        ("verbose", C.c_int32),
    ]


class ak_newton_stats(C.Structure):
    _fields_ = [
        ("solved", C.c_int32),
        ("outer_iterations", C.c_int32),
        ("inner_iterations", C.c_int64),
        ("n_res", C.c_double),
        ("tol", C.c_double),
        ("t_seconds", C.c_double),
        ("flags", C.c_int32),
    ]


# AK_USER callbacks: int (*)(void* user, uint64_t stream, double* u, double* res) / (..., const double* u, double* v, double* out)
USER_RESIDUAL = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p)
USER_JVP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p)
# AK_PRECOND_USER: int (*)(void* user, uint64_t stream, const double* x, double* y)
PRECOND_APPLY = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p)

NEWTON_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double)

SQRT_EPS = 2.220446049250313e-16 ** 0.5


def default_krylov_opts(**kw):
    """Krylov.jl gmres!/cg! keyword defaults (atol = rtol = sqrt(eps), itmax = 0 -> 2n)."""
    o = ak_krylov_opts(SQRT_EPS, SQRT_EPS, 0, 0, 0, 0, AK_FUSE_SWEEP, AK_PRECOND_NONE, 0, AK_PRECOND_NONE, 0, None, None, None, None)
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown krylov kwarg {k!r}")
        setattr(o, k, v)
    return o


def default_newton_opts(**kw):
    """newton_krylov! keyword defaults: src/Ariadne.jl:290-299."""
    o = ak_newton_opts()
    o.tol_rel, o.tol_abs, o.max_niter = 1.0e-6, 1.0e-12, 50
    o.forcing, o.eta, o.eta_max, o.gamma = AK_FORCING_EW, 0.1, 0.999, 0.9
    o.algo, o.memory, o.max_basis = AK_ALGO_GMRES, 20, 0
    o.krylov = default_krylov_opts()
    o.krylov_rtol_override, o.verbose = 0, 0
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown newton kwarg {k!r}")
        setattr(o, k, v)
    return o
