"""Python loader of the CPU oracle (oracle/nk_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (newtonkrylov.jl_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import newtonkrylov_jl_b200  # noqa: F401  (struct layouts only; does not load the CUDA library)
from newtonkrylov_jl_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnk_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "nk_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ariadne_b200.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libnk_oracle.so"], check=True, stdout=subprocess.PIPE,
                   stderr=subprocess.STDOUT)
    return LIB_PATH


_dp = A.c_double_p
_P = C.POINTER(A.ak_problem)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    sig = {
        "ok_num_threads": (C.c_int, []),
        "ok_set_num_threads": (C.c_int, [C.c_int]),
        "ok_problem_size": (C.c_int64, [_P]),
        "ok_dot": (C.c_double, [C.c_int64, _dp, _dp]),
        "ok_nrm2": (C.c_double, [C.c_int64, _dp]),
        "ok_scal": (None, [C.c_int64, C.c_double, _dp]),
        "ok_axpy": (None, [C.c_int64, C.c_double, _dp, _dp]),
        "ok_axpby": (None, [C.c_int64, C.c_double, _dp, C.c_double, _dp]),
        "ok_copy": (None, [C.c_int64, _dp, _dp]),
        "ok_fill": (None, [C.c_int64, _dp, C.c_double]),
        "ok_ref": (None, [C.c_int64, _dp, _dp, C.c_double, C.c_double]),
        "ok_divcopy": (None, [C.c_int64, _dp, _dp, C.c_double]),
        "ok_halo_pack": (None, [_dp, _dp, C.c_int64, C.c_int64]),
        "ok_halo_unpack": (None, [_dp, _dp, C.c_int64, C.c_int64, C.c_int32]),
        "ok_dg_matrix": (None, [_dp]),
        "ok_dg_plus": (None, [_dp, _dp, C.c_int64, C.c_double]),
        "ok_dg_minus": (None, [_dp, _dp, C.c_int64, C.c_double]),
        "ok_residual": (None, [_P, _dp, _dp]),
        "ok_jvp": (None, [_P, _dp, _dp, _dp]),
        "ok_jvp_transpose_dense": (None, [_P, _dp, _dp, _dp]),
        "ok_sym_givens": (None, [C.c_double, C.c_double, _dp, _dp, _dp]),
        "ok_krylov_create": (C.c_void_p, [C.c_int32, C.c_int64, C.c_int32]),
        "ok_krylov_destroy": (None, [C.c_void_p]),
        "ok_krylov_x": (_dp, [C.c_void_p]),
        "ok_krylov_basis_size": (C.c_int64, [C.c_void_p]),
        "ok_krylov_solve": (C.c_int, [C.c_void_p, _P, _dp, _dp, C.POINTER(A.ak_krylov_opts),
                                      C.POINTER(A.ak_krylov_stats), _dp, C.c_int64]),
        "ok_forcing_ew": (C.c_double, [C.c_double] * 6),
        "ok_precond_apply": (None, [_P, _dp, C.c_int32, C.c_int32, _dp, _dp]),
        "ok_newton": (C.c_int, [_P, _dp, _dp, C.POINTER(A.ak_newton_opts), C.POINTER(A.ak_newton_stats), _dp,
                                A.c_int64_p, _dp, C.c_int32]),
        "ok_implicit_solve": (C.c_int, [_P, _dp, C.c_int32, C.POINTER(A.ak_newton_opts), A.c_int32_p, A.c_int64_p,
                                        A.c_int32_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _arr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_problem(kind, nx, ny=1, *, bc=A.AK_BC_ZERO, scheme=A.AK_STEADY, dx=0.0, dy=0.0, lam=0.0, a=0.0, dt=0.0,
                 un=None):
    """ak_problem with HOST pointers (oracle side).  Keeps `un` alive on the returned object."""
    p = A.ak_problem()
    p.kind, p.bc, p.scheme, p.jvp_mode = kind, bc, scheme, A.AK_JVP_ANALYTIC
    p.nx, p.ny, p.gny, p.gy0 = nx, ny, ny, 0
    p.dx, p.dy, p.lambda_, p.a, p.dt, p.fd_eps = dx, dy, lam, a, dt, 0.0
    if un is not None:
        un = _arr(un)
        p.un = un.ctypes.data
        p._keep = un
    return p


def make_user_problem(n, F, jvp=None, fd_eps=0.0):
    """AK_USER problem for the oracle: `F(res, u)` / `jvp(out, u, v)` are Python callables on NumPy arrays that alias
    the oracle's host vectors (the same callbacks the CUDA library takes, with host pointers and stream 0)."""
    p = A.ak_problem()
    p.kind, p.bc, p.scheme = A.AK_USER, A.AK_BC_ZERO, A.AK_STEADY
    p.jvp_mode = A.AK_JVP_ANALYTIC if jvp is not None else A.AK_JVP_FD
    p.nx, p.ny, p.gny, p.gy0 = n, 1, 1, 0
    p.fd_eps = fd_eps

    def view(ptr):
        return np.ctypeslib.as_array(C.cast(ptr, _dp), shape=(n,))

    def res_cb(_user, _stream, u, res):
        F(view(res), view(u))
        return 0

    def jvp_cb(_user, _stream, u, v, out):
        jvp(view(out), view(u), view(v))
        return 0

    cb_r = A.USER_RESIDUAL(res_cb)
    cb_j = A.USER_JVP(jvp_cb) if jvp is not None else None
    p.user_residual = C.cast(cb_r, C.c_void_p).value
    p.user_jvp = C.cast(cb_j, C.c_void_p).value if cb_j is not None else None
    p._keep = (cb_r, cb_j)
    return p


def residual(p, u):
    """returns (res, u_after) — u may be mutated by boundary code (1-D heat)."""
    u = _arr(u).copy()
    res = np.empty_like(u)
    load().ok_residual(C.byref(p), _d(u), _d(res))
    return res, u


def jvp(p, u, v):
    """returns (out, v_after)."""
    u = _arr(u)
    v = _arr(v).copy()
    out = np.empty_like(v)
    load().ok_jvp(C.byref(p), _d(u), _d(v), _d(out))
    return out, v


def jvp_transpose_dense(p, u, v):
    u, v = _arr(u), _arr(v)
    out = np.empty_like(v)
    load().ok_jvp_transpose_dense(C.byref(p), _d(u), _d(v), _d(out))
    return out


def dense_jacobian(p, u):
    n = int(load().ok_problem_size(C.byref(p)))
    J = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        J[:, j] = jvp(p, u, e.reshape(np.shape(u)))[0].reshape(-1)
    return J


def sym_givens(a, b):
    c, s, r = C.c_double(), C.c_double(), C.c_double()
    load().ok_sym_givens(a, b, C.byref(c), C.byref(s), C.byref(r))
    return c.value, s.value, r.value


def forcing_ew(eta_max, gamma, eta, tol, n_res, n_res_prior):
    return load().ok_forcing_ew(eta_max, gamma, eta, tol, n_res, n_res_prior)


def user_precond(n, apply):
    """(callback pointer, keep-alive object) for AK_PRECOND_USER on the oracle side: `apply(y, x)` on NumPy views."""
    def cb(_user, _stream, x, y):
        apply(np.ctypeslib.as_array(C.cast(y, _dp), shape=(n,)), np.ctypeslib.as_array(C.cast(x, _dp), shape=(n,)))
        return 0

    f = A.PRECOND_APPLY(cb)
    return C.cast(f, C.c_void_p).value, f


def precond_apply(p, u, kind, x, itmax=0):
    """y = P x for a native preconditioner kind (the oracle's restatement)."""
    u, x = _arr(u), _arr(x)
    y = np.empty_like(x)
    load().ok_precond_apply(C.byref(p), _d(u), kind, itmax, _d(x), _d(y))
    return y


def krylov_solve(p, u, b, *, algo=A.AK_ALGO_GMRES, memory=20, hist_cap=0, **kw):
    """One linear solve J(u) x = b.  Returns (x, stats dict, residual history)."""
    lib = load()
    u, b = _arr(u), _arr(b)
    n = b.size
    ws = lib.ok_krylov_create(algo, n, memory)
    o = A.default_krylov_opts(**kw)
    st = A.ak_krylov_stats()
    hist = np.zeros(max(hist_cap, 1))
    lib.ok_krylov_solve(ws, C.byref(p), _d(u), _d(b), C.byref(o), C.byref(st), _d(hist), hist_cap)
    x = np.ctypeslib.as_array(lib.ok_krylov_x(ws), shape=(n,)).copy().reshape(b.shape)
    nb = int(lib.ok_krylov_basis_size(ws))
    lib.ok_krylov_destroy(ws)
    stats = dict(niter=int(st.niter), solved=bool(st.solved), inconsistent=bool(st.inconsistent),
                 breakdown=bool(st.breakdown), npass=int(st.npass), rnorm=st.rnorm, beta=st.beta, basis=nb)
    return x, stats, hist[: min(hist_cap, st.niter + 1)].copy()


def newton(p, u0, opts=None, hist_cap=64):
    """newton_krylov!(F!, u, p, res).  Returns (u, stats dict, history list of dicts)."""
    lib = load()
    u = _arr(u0).copy()
    res = np.zeros_like(u)
    o = opts if opts is not None else A.default_newton_opts()
    st = A.ak_newton_stats()
    hn, hi, he = np.zeros(hist_cap), np.zeros(hist_cap, dtype=np.int64), np.zeros(hist_cap)
    lib.ok_newton(C.byref(p), _d(u), _d(res), C.byref(o), C.byref(st), _d(hn), hi.ctypes.data_as(A.c_int64_p),
                  _d(he), hist_cap)
    k = min(hist_cap, st.outer_iterations + 1)
    hist = [dict(n_res=hn[i], inner=int(hi[i]), eta=(he[i] if i else None)) for i in range(k)]
    stats = dict(solved=bool(st.solved), outer_iterations=int(st.outer_iterations),
                 inner_iterations=int(st.inner_iterations), n_res=st.n_res, tol=st.tol, t=st.t_seconds,
                 flags=int(st.flags))
    return u, stats, hist


def implicit_solve(p, un0, nsteps, opts=None):
    lib = load()
    un = _arr(un0).copy()
    o = opts if opts is not None else A.default_newton_opts(tol_abs=6.0e-6)
    newt = np.zeros(nsteps, dtype=np.int32)
    inner = np.zeros(nsteps, dtype=np.int64)
    solved = np.zeros(nsteps, dtype=np.int32)
    lib.ok_implicit_solve(C.byref(p), _d(un), nsteps, C.byref(o), newt.ctypes.data_as(A.c_int32_p),
                          inner.ctypes.data_as(A.c_int64_p), solved.ctypes.data_as(A.c_int32_p))
    return un, newt, inner, solved


def num_threads():
    return int(load().ok_num_threads())


def use_all_cores():
    """OpenMP team = the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun exports 1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return int(load().ok_set_num_threads(n))
