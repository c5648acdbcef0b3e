"""Host-side mirror of the reference's interface for the JFNK hot path, over the C ABI.

The reference is a Julia package; Julia is not available in this image, so the host side that
drives the C ABI is written in Python with the reference's names, argument meaning and error
behaviour (Julia's trailing `!` becomes a trailing `_`):

    newton_krylov_, newton_krylov        src/Ariadne.jl:245-263,288-372
    JacobianOperator, mul_, collect      src/Ariadne.jl:34-57,93-107,140-162
    Fixed, EisenstatWalker, inital       src/Ariadne.jl:180-217
    Stats                                src/Ariadne.jl:265-276
    krylov_workspace / krylov_solve_     Krylov.jl call sites src/Ariadne.jl:317-318,338-340
    kdot, knorm, kscal_, kaxpy_, ...     examples/halovector.jl:51-147
    HaloVector                           examples/halovector.jl:3-45
    G_Euler_, solve                      examples/implicit.jl:8-13,54-78
    bratu_, heat_1D_, diffusion_, ...    examples/bratu.jl:14-24, heat_1D.jl:12-25, heat_2D.jl:45-62,
                                         heat_1D_DG.jl:32-36

The Julia wrapper a maintainer would add (julia/AriadneB200.jl) binds exactly the same C
entry points with `ccall`.  All arithmetic on vectors happens in CUDA kernels; this file only
holds scalar logic.  Nothing here falls back to the CPU.
"""
import ctypes as C
import math
import time
from collections import namedtuple

import numpy as np

from . import _abi as A
from . import _lib as L

# ---------------------------------------------------------------------------------------------
# context and device vectors
# ---------------------------------------------------------------------------------------------
_contexts = {}


class Context:
    """One stream on one GPU (include/ariadne_b200.h: ak_ctx)."""

    def __init__(self, device=0):
        self.lib = L.load()
        h = C.c_void_p()
        L.check(self.lib.ak_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self.rank, self.nranks = 0, 1

    def sync(self):
        L.check(self.lib.ak_ctx_sync(self.h))

    def launch_count(self, reset=False):
        return int(self.lib.ak_ctx_launch_count(self.h, 1 if reset else 0))

    def timer_start(self):
        L.check(self.lib.ak_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        L.check(self.lib.ak_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile(self, on=True):
        L.check(self.lib.ak_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, cls):
        """(launch count, total device ms) of one kernel class since profile(True)."""
        cnt, ms = C.c_int64(), C.c_double()
        L.check(self.lib.ak_profile_read(self.h, cls, C.byref(cnt), C.byref(ms)))
        return cnt.value, ms.value

    @property
    def stream(self):
        return int(self.lib.ak_ctx_stream(self.h))

    def init_comm(self, nranks, rank, unique_id):
        L.check(self.lib.ak_comm_init(self.h, nranks, rank, unique_id))
        self.rank, self.nranks = rank, nranks

    def enable_p2p(self, halo_doubles):
        """Map every rank's mailbox / ghost-row block over NVLink (CUDA IPC); collective call."""
        L.check(self.lib.ak_comm_enable_p2p(self.h, int(halo_doubles)))

    @property
    def p2p(self):
        return bool(self.lib.ak_comm_p2p_enabled(self.h))

    def use_p2p(self, on):
        L.check(self.lib.ak_comm_use_p2p(self.h, 1 if on else 0))

    def barrier(self):
        L.check(self.lib.ak_comm_barrier(self.h))

    def close(self):
        if self.h:
            self.lib.ak_ctx_destroy(self.h)
            self.h = None


def get_context(device=0):
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


def comm_unique_id():
    buf = C.create_string_buffer(128)
    L.check(L.load().ak_comm_unique_id(buf))
    return buf.raw


class DeviceVector:
    """fp64 vector resident in HBM ("compact slab": no ghost cells).  `shape` is (n,) or (ny, nx)."""

    def __init__(self, ctx, shape, ptr=None, owner=None):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.n = int(np.prod(self.shape))
        if ptr is None:
            p = C.c_void_p()
            L.check(ctx.lib.ak_malloc(ctx.h, self.n, C.byref(p)))
            self.ptr = p.value
            self._owned = True
        else:
            self.ptr = int(ptr)
            self._owned = False
        self._owner = owner  # keeps e.g. a torch tensor alive

    def __del__(self):
        try:
            if getattr(self, "_owned", False) and self.ctx.h:
                self.ctx.lib.ak_free(self.ctx.h, C.c_void_p(self.ptr))
        except Exception:
            pass

    def __len__(self):
        return self.n

    @classmethod
    def from_numpy(cls, arr, ctx=None):
        ctx = ctx or get_context()
        a = np.ascontiguousarray(arr, dtype=np.float64)
        v = cls(ctx, a.shape)
        L.check(ctx.lib.ak_upload(ctx.h, C.c_void_p(v.ptr), a.ctypes.data_as(C.c_void_p), v.n))
        return v

    @classmethod
    def from_torch(cls, t, ctx=None):
        ctx = ctx or get_context()
        assert t.is_cuda and t.is_contiguous() and str(t.dtype) == "torch.float64"
        return cls(ctx, tuple(t.shape), ptr=t.data_ptr(), owner=t)

    def numpy(self):
        out = np.empty(self.shape, dtype=np.float64)
        L.check(self.ctx.lib.ak_download(self.ctx.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), self.n))
        return out

    def set(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(self.shape)
        L.check(self.ctx.lib.ak_upload(self.ctx.h, C.c_void_p(self.ptr), a.ctypes.data_as(C.c_void_p), self.n))
        return self

    def similar(self):
        return DeviceVector(self.ctx, self.shape)

    def zero(self):
        v = self.similar()
        kfill_(v, 0.0)
        return v

    def copy(self):
        v = self.similar()
        kcopy_(self.n, v, self)
        return v


class HaloVector(DeviceVector):
    """examples/halovector.jl:3-45: logical length = interior only.  The ghost ring of the
    reference's (N+2)x(M+2) OffsetArray is not stored; `from_padded` / `padded` convert."""

    @classmethod
    def from_padded(cls, padded, ctx=None):
        """`padded[i, j]` as in Julia (first index = x, 0:N+1), including the ghost ring."""
        ctx = ctx or get_context()
        p = np.asarray(padded, dtype=np.float64)
        nx, ny = p.shape[0] - 2, p.shape[1] - 2
        # Julia column-major (i fastest) == C-order array indexed [j][i]
        pj = np.ascontiguousarray(p.T)
        dp = DeviceVector.from_numpy(pj, ctx)
        v = cls(ctx, (ny, nx))
        L.check(ctx.lib.ak_halo_pack(ctx.h, C.c_void_p(v.ptr), C.c_void_p(dp.ptr), nx, ny))
        ctx.sync()
        return v

    def padded(self, bc=A.AK_BC_ZERO):
        ny, nx = self.shape
        dp = DeviceVector(self.ctx, (ny + 2, nx + 2))
        L.check(self.ctx.lib.ak_halo_unpack(self.ctx.h, C.c_void_p(dp.ptr), C.c_void_p(self.ptr), nx, ny, bc))
        return dp.numpy().T.copy()

    def similar(self):
        return HaloVector(self.ctx, self.shape)


def _ptr(v):
    return C.c_void_p(v.ptr) if v is not None else None


# ---------------------------------------------------------------------------------------------
# Krylov.k* hooks (examples/halovector.jl:51-147)
# ---------------------------------------------------------------------------------------------
def kdot(n, x, y):
    out = C.c_double()
    L.check(x.ctx.lib.ak_dot(x.ctx.h, n, _ptr(x), _ptr(y), C.byref(out)))
    return out.value


def knorm(n, x):
    out = C.c_double()
    L.check(x.ctx.lib.ak_nrm2(x.ctx.h, n, _ptr(x), C.byref(out)))
    return out.value


def kscal_(n, s, x):
    L.check(x.ctx.lib.ak_scal(x.ctx.h, n, s, _ptr(x)))
    return x


def kaxpy_(n, s, x, y):
    L.check(x.ctx.lib.ak_axpy(x.ctx.h, n, s, _ptr(x), _ptr(y)))
    return y


def kaxpby_(n, s, x, t, y):
    L.check(x.ctx.lib.ak_axpby(x.ctx.h, n, s, _ptr(x), t, _ptr(y)))
    return y


def kcopy_(n, y, x):
    L.check(x.ctx.lib.ak_copy(x.ctx.h, n, _ptr(y), _ptr(x)))
    return y


def kfill_(x, val):
    L.check(x.ctx.lib.ak_fill(x.ctx.h, x.n, _ptr(x), val))
    return x


def kref_(n, x, y, c, s):
    L.check(x.ctx.lib.ak_ref(x.ctx.h, n, _ptr(x), _ptr(y), c, s))
    return x, y


def kdivcopy_(n, y, x, s):
    L.check(x.ctx.lib.ak_divcopy(x.ctx.h, n, _ptr(y), _ptr(x), s))
    return y


# ---------------------------------------------------------------------------------------------
# residual functions F!(res, u, p) — native stencils selected by name
# ---------------------------------------------------------------------------------------------
class NativeResidual:
    """A residual the library has a hand-written kernel for.  Calling it evaluates
    `res <- F(u)` on the device, like the reference's `F!(res, u, p)` (src/Ariadne.jl:250-256)."""

    kind = None
    name = "?"

    def problem(self, u, p, coef=None):
        raise NotImplementedError

    def __call__(self, res, u, p):
        prob = self.problem(u, p)
        L.check(u.ctx.lib.ak_residual(u.ctx.h, C.byref(prob), _ptr(u), _ptr(res), None))
        return None

    def __repr__(self):
        return f"<native residual {self.name}>"


def _base_problem(kind, nx, ny=1, **kw):
    p = A.ak_problem()
    p.kind, p.bc, p.scheme, p.jvp_mode = kind, kw.get("bc", A.AK_BC_ZERO), kw.get("scheme", A.AK_STEADY), A.AK_JVP_ANALYTIC
    p.nx, p.ny = nx, ny
    p.gny, p.gy0 = kw.get("gny", ny), kw.get("gy0", 0)
    p.dx, p.dy = kw.get("dx", 0.0), kw.get("dy", 0.0)
    p.lambda_, p.a, p.dt = kw.get("lam", 0.0), kw.get("a", 0.0), kw.get("dt", 0.0)
    p.fd_eps = 0.0
    p.un = kw.get("un")
    p.coef = kw.get("coef")
    p.work = None
    p.user_residual = p.user_jvp = p.user_data = None
    return p


class _Simple2(NativeResidual):
    """test/runtests.jl:4-7, examples/simple.jl:6-9"""
    kind, name = A.AK_SIMPLE2, "F! (2x2)"

    def problem(self, u, p, coef=None):
        return _base_problem(A.AK_SIMPLE2, 2)


class _Bratu1D(NativeResidual):
    """bratu!(res, y, (dx, lambda)): examples/bratu.jl:14-24"""
    kind, name = A.AK_BRATU1D, "bratu!"

    def problem(self, u, p, coef=None):
        dx, lam = p
        return _base_problem(A.AK_BRATU1D, u.n, dx=dx, lam=lam, coef=coef.ptr if coef is not None else None)


class _Bratu2D(NativeResidual):
    """2-D Bratu on the heat_2D grid, p = (dx, dy, lambda); u has shape (ny, nx).
    A slab of a larger global grid passes p = (dx, dy, lambda, gny, gy0)."""
    kind, name = A.AK_BRATU2D, "bratu2d!"

    def problem(self, u, p, coef=None):
        dx, dy, lam = p[:3]
        ny, nx = u.shape
        gny, gy0 = (p[3], p[4]) if len(p) >= 5 else (ny, 0)
        return _base_problem(A.AK_BRATU2D, nx, ny, dx=dx, dy=dy, lam=lam, gny=gny, gy0=gy0,
                             coef=coef.ptr if coef is not None else None)


simple_F_ = _Simple2()
bratu_ = _Bratu1D()
bratu2d_ = _Bratu2D()


# caller-supplied residuals: the generic seam of newton_krylov!(F!, u, p, res) -----------------------
class _DeviceArrayView:
    """Minimal __cuda_array_interface__ carrier so that torch can alias a raw device pointer."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def as_torch(ptr, shape, device_index=0):
    """torch.float64 CUDA tensor aliasing `shape` doubles at device address `ptr` (no copy)."""
    import torch

    return torch.as_tensor(_DeviceArrayView(ptr, shape), device=torch.device("cuda", device_index))


class UserResidual(NativeResidual):
    """Any `F!(res, u, p)` (contract: src/Ariadne.jl:250-256) written by the caller as device code — the
    seam that lets `newton_krylov!` take arbitrary residuals (examples/bvp.jl:10-23, spring.jl:13-17, ...).

    `F_(res, u, p)` and, optionally, `jvp_(out, u, v, p)` (out <- J(u) v: what the reference obtains from
    Enzyme forward mode, src/Ariadne.jl:48-57) receive torch.float64 CUDA tensors that ALIAS the library's
    device vectors; they run on the library's stream (C ABI: AK_USER, ak_user_residual_fn / ak_user_jvp_fn).
    Without `jvp_` the library forms J v = (F(u + eps v) - F(u)) / eps itself (AK_JVP_FD)."""

    kind = A.AK_USER

    def __init__(self, F_, jvp_=None, name=None, fd_eps=0.0):
        self.F_, self.jvp_, self.fd_eps = F_, jvp_, float(fd_eps)
        self.name = name or getattr(F_, "__name__", "F!")
        self._p, self._shape, self._dev = None, None, 0
        self._streams = {}
        self._cb_res = A.USER_RESIDUAL(self._residual_cb)
        self._cb_jvp = A.USER_JVP(self._jvp_cb) if jvp_ is not None else None

    @classmethod
    def from_function(cls, F, name=None):
        """Out-of-place `F(u, p) -> tensor` (src/Ariadne.jl:245-248 wraps it as `res .= F(u, p)`); the tangent
        comes from forward-mode AD of the same code (torch.func.jvp), like Enzyme's in the reference."""
        import torch

        def F_(res, u, p):
            res.copy_(F(u, p))

        def jvp_(out, u, v, p):
            out.copy_(torch.func.jvp(lambda x: F(x, p), (u,), (v,))[1])

        return cls(F_, jvp_, name=name or getattr(F, "__name__", "F"))

    def _stream(self, handle):
        import torch

        if handle not in self._streams:
            self._streams[handle] = torch.cuda.ExternalStream(handle, device=torch.device("cuda", self._dev))
        return self._streams[handle]

    def _residual_cb(self, _user, stream, u, res):
        try:
            import torch

            with torch.cuda.stream(self._stream(stream)):
                self.F_(as_torch(res, self._shape, self._dev), as_torch(u, self._shape, self._dev), self._p)
            return 0
        except Exception:  # noqa: BLE001 - must not unwind through the C frames
            import traceback

            traceback.print_exc()
            return 1

    def _jvp_cb(self, _user, stream, u, v, out):
        try:
            import torch

            with torch.cuda.stream(self._stream(stream)):
                self.jvp_(as_torch(out, self._shape, self._dev), as_torch(u, self._shape, self._dev),
                          as_torch(v, self._shape, self._dev), self._p)
            return 0
        except Exception:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            return 1

    def problem(self, u, p, coef=None):
        self._p, self._shape, self._dev = p, u.shape, u.ctx.device
        prob = _base_problem(A.AK_USER, u.n, coef=coef.ptr if coef is not None else None)
        prob.jvp_mode = A.AK_JVP_ANALYTIC if self.jvp_ is not None else A.AK_JVP_FD
        prob.fd_eps = self.fd_eps
        prob.user_residual = C.cast(self._cb_res, C.c_void_p).value
        prob.user_jvp = C.cast(self._cb_jvp, C.c_void_p).value if self._cb_jvp is not None else None
        prob.user_data = None
        return prob


# right-hand sides f!(du, u, p, t) of the implicit examples -------------------------------------
class NativeRHS:
    kind = None
    name = "?"

    def fill(self, prob, u, p):
        raise NotImplementedError


class bc_zero_:  # heat_1D.jl:34-37 bc!, heat_2D.jl:28-38 bc_zero!
    code = A.AK_BC_ZERO


class bc_periodic_:  # heat_1D.jl:39-42 periodic_bc!, heat_2D.jl:15-26 bc_periodic!
    code = A.AK_BC_PERIODIC


bc_ = bc_zero_
periodic_bc_ = bc_periodic_


class _Heat1D(NativeRHS):
    """heat_1D!(du, u, (a, dx, bc!), t): examples/heat_1D.jl:12-25"""
    kind, name = A.AK_HEAT1D, "heat_1D!"

    def fill(self, prob, u, p):
        a, dx, bc = p
        prob.kind, prob.nx, prob.ny, prob.gny = self.kind, u.n, 1, 1
        prob.a, prob.dx, prob.bc = a, dx, bc.code


class _Diffusion2D(NativeRHS):
    """diffusion!(du, u, (a, dx, dy, bc!), t): examples/heat_2D.jl:45-62"""
    kind, name = A.AK_HEAT2D, "diffusion!"

    def fill(self, prob, u, p):
        a, dx, dy, bc = p[:4]
        ny, nx = u.shape
        prob.kind, prob.nx, prob.ny = self.kind, nx, ny
        prob.gny, prob.gy0 = (p[4], p[5]) if len(p) >= 6 else (ny, 0)
        prob.a, prob.dx, prob.dy, prob.bc = a, dx, dy, bc.code


class _Heat1DDG(NativeRHS):
    """heat_1D!(du, u, (D1m, D1p), t) with the coupled LGL(4) upwind operators of
    examples/heat_1D_DG.jl:17-36; here p = (h,) is the element width of the periodic mesh."""
    kind, name = A.AK_HEAT1D_DG, "heat_1D! (DG)"

    def fill(self, prob, u, p):
        (h,) = p
        prob.kind, prob.nx, prob.ny, prob.gny = self.kind, u.n, 1, 1
        prob.dx, prob.bc = h, A.AK_BC_PERIODIC


heat_1D_ = _Heat1D()
diffusion_ = _Diffusion2D()
heat_1D_DG_ = _Heat1DDG()


class _Scheme:
    def __init__(self, code, name):
        self.code, self.name = code, name


G_Euler_ = _Scheme(A.AK_EULER, "G_Euler!")          # examples/implicit.jl:8-13
G_Midpoint_ = _Scheme(A.AK_MIDPOINT, "G_Midpoint!")  # examples/implicit.jl:17-25
G_Trapezoid_ = _Scheme(A.AK_TRAPEZOID, "G_Trapezoid!")  # examples/implicit.jl:29-37


class ImplicitResidual(NativeResidual):
    """F!(res, u, (u_n, dt, du, p, t)) = G!(res, u_n, dt, f!, du, u, p, t): examples/implicit.jl:61"""

    def __init__(self, G_, f_):
        self.G_, self.f_ = G_, f_
        self.kind, self.name = f_.kind, f"{G_.name} o {f_.name}"

    def problem(self, u, p, coef=None):
        un, dt, _du, p_inner, _t = p
        prob = _base_problem(self.f_.kind, u.n, scheme=self.G_.code, dt=dt, un=un.ptr)
        self.f_.fill(prob, u, p_inner)
        return prob


# ---------------------------------------------------------------------------------------------
# JacobianOperator: src/Ariadne.jl:34-57
# ---------------------------------------------------------------------------------------------
class JacobianOperator:
    """Matrix-free J(u) of a native residual.  Aliases f, res, u, p (never copies) like the
    reference constructor (src/Ariadne.jl:39-41)."""

    def __init__(self, f, res, u, p, coef=None, jvp_mode="analytic", fd_eps=0.0):
        if not isinstance(f, NativeResidual):
            raise TypeError("JacobianOperator needs a native residual (no Enzyme / no CPU fallback on this path)")
        self.f, self.res, self.u, self.p = f, res, u, p
        self.coef = coef  # lambda*exp(u) cached by the last residual evaluation at this u (Bratu), or None
        # "analytic": exact tangent (what the reference's Enzyme forward mode computes, default);
        # "fd": fused (F(u + eps v) - F(u)) / eps in one pass (BASELINE north_star wording; O(1e-8) error)
        self.jvp_mode, self.fd_eps = jvp_mode, float(fd_eps)

    def size(self):  # src/Ariadne.jl:44
        return (len(self.res), len(self.u))

    @property
    def shape(self):
        return self.size()

    def eltype(self):  # :45
        return np.float64

    def __len__(self):  # :46  length(J) = prod(size(J))
        m, n = self.size()
        return m * n

    def problem(self):
        prob = self.f.problem(self.u, self.p, coef=self.coef)
        if self.jvp_mode == "fd":      # one fused pass for 2-D Bratu, two residual evaluations otherwise
            prob.jvp_mode = A.AK_JVP_FD_FUSED if prob.kind in (A.AK_BRATU1D, A.AK_BRATU2D) else A.AK_JVP_FD
            prob.fd_eps = self.fd_eps
        elif self.jvp_mode == "fd2":   # generic (F(u + eps v) - F(u)) / eps through the residual itself
            prob.jvp_mode, prob.fd_eps = A.AK_JVP_FD, self.fd_eps
        return prob

    @property
    def T(self):
        return TransposeOperator(self)


class TransposeOperator:
    """transpose(J) / adjoint(J): src/Ariadne.jl:87-88"""

    def __init__(self, J):
        self.parent = J

    def size(self):
        m, n = self.parent.size()
        return (n, m)


def transpose(J):
    return TransposeOperator(J)


adjoint = transpose


def mul_(out, J, v):
    """mul!(out, J, v): src/Ariadne.jl:48-57 (and :93-107 for transpose(J)).  `v` may have its
    boundary entries overwritten, exactly as forward mode through bc!(u) does in the reference."""
    if isinstance(J, TransposeOperator):
        Jp = J.parent
        prob = Jp.problem()
        L.check(v.ctx.lib.ak_jvp_transpose(v.ctx.h, C.byref(prob), _ptr(Jp.u), _ptr(v), _ptr(out)))
        return None
    prob = J.problem()
    L.check(v.ctx.lib.ak_jvp(v.ctx.h, C.byref(prob), _ptr(J.u), _ptr(v), _ptr(out)))
    return None


def mul_batched_(Out, J, V):
    """mul!(Out, J, V) for matrices (src/Ariadne.jl:69-83).  `V`, `Out`: DeviceVectors of shape (ncols, n) — the
    column-major n x ncols matrices of the reference, one column per row of the C-ordered array."""
    ncols, n = V.shape[0], int(np.prod(V.shape[1:]))
    prob = J.problem()
    L.check(V.ctx.lib.ak_jvp_batched(V.ctx.h, C.byref(prob), _ptr(J.u), _ptr(V), n, _ptr(Out), n, ncols))
    return None


def _ring_colours(N, s):
    """Colours of N columns on a ring such that equal colours are >= s apart: j mod s on the main range, one
    colour of their own for the N mod s trailing columns.  Returns (colour array, number of colours)."""
    if N <= s:
        return np.arange(N), N
    main = N - (N % s)
    col = np.arange(N) % s
    col[main:] = s + np.arange(N - main)
    return col, s + (N - main)


def collect(JOp, sparse=False, bandwidth=None):
    """Base.collect(J): src/Ariadne.jl:140-162.

    sparse = False: one JVP per column, dense ndarray (small n only).
    sparse = True:  the reference assembles a SparseMatrixCSC with N JVPs; here the stencil structure is used
    (SURVEY §8f-4): columns that cannot share a row are probed together (ring colouring, 3-15 colours in 1-D,
    9-25 on the 2-D five-point grid), so the whole Jacobian costs a handful of JVPs.  Returns scipy.sparse CSR.
    `bandwidth`: half-width for AK_USER operators (|i - j| <= bandwidth on the ring)."""
    if not sparse:
        return _collect_dense(JOp)
    tr = isinstance(JOp, TransposeOperator)
    J = JOp.parent if tr else JOp
    prob = J.problem()
    u = J.u
    n = u.n
    colour, ncol, offsets = probe_plan(prob.kind, prob.bc, u.shape, bandwidth)
    probes = np.zeros((ncol, n))
    probes[colour, np.arange(n)] = 1.0
    V = DeviceVector.from_numpy(probes, u.ctx)
    Out = DeviceVector(u.ctx, (ncol, n))
    kfill_(Out, 0.0)
    mul_batched_(Out, J, V)
    M = assemble_probed(Out.numpy(), colour, offsets, n)
    return M.T.tocsr() if tr else M


def probe_plan(kind, bc, shape, bandwidth=None):
    """(colour of every column, number of colours, candidate column of every row per stencil offset) — pure host logic."""
    n = int(np.prod(shape))
    if kind in (A.AK_BRATU2D, A.AK_HEAT2D):
        ny, nx = shape
        cx, ncx = _ring_colours(nx, 3)
        cy, ncy = _ring_colours(ny, 3)
        colour = (cy[:, None] * ncx + cx[None, :]).reshape(-1)
        ii, jj = np.meshgrid(np.arange(nx), np.arange(ny))
        offsets = [(((jj + dy_) % ny) * nx + (ii + dx_) % nx).reshape(-1)
                   for dx_, dy_ in ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1))]
        return colour, ncx * ncy, offsets
    w = bandwidth
    if w is None:
        if kind == A.AK_BRATU1D:
            w = 1
        elif kind == A.AK_HEAT1D:
            w = 3 if bc == A.AK_BC_PERIODIC else 1   # periodic_bc! copies u[end-1] into u[1]: ring distance 3
        elif kind == A.AK_HEAT1D_DG:
            w = 7                                     # D1m * D1p couples an element with both neighbours
        else:
            raise NotImplementedError("collect(J, sparse=True): pass bandwidth= for this operator")
    colour, ncol = _ring_colours(n, 2 * w + 1)
    rows = np.arange(n)
    offsets = [(rows + d) % n for d in range(-w, w + 1)] if n > 2 * w + 1 else [np.full(n, j) for j in range(n)]
    return colour, ncol, offsets


def assemble_probed(Y, colour, offsets, n):
    """CSR matrix from the probe products Y[c] = J * (sum of the unit vectors of colour c)."""
    import scipy.sparse as sp

    rows = np.arange(n)
    R, Cc, D = [], [], []
    done = np.zeros((0, n), dtype=np.int64)
    for cols in offsets:
        fresh = np.ones(n, dtype=bool)
        for prev in done:            # tiny grids: two offsets can wrap onto the same column of a row
            fresh &= prev != cols
        done = np.vstack([done, cols[None, :]])
        vals = Y[colour[cols], rows]
        nz = (vals != 0.0) & fresh
        R.append(rows[nz]); Cc.append(cols[nz]); D.append(vals[nz])
    return sp.csr_matrix((np.concatenate(D), (np.concatenate(R), np.concatenate(Cc))), shape=(n, n))


def _collect_dense(JOp):
    """One JVP per column (the literal algorithm of src/Ariadne.jl:140-162)."""
    J = JOp.parent if isinstance(JOp, TransposeOperator) else JOp
    N, M = JOp.size()
    v = (J.res if isinstance(JOp, TransposeOperator) else J.u).zero()
    out = (J.u if isinstance(JOp, TransposeOperator) else J.res).zero()
    dense = np.zeros((N, M))
    e = np.zeros(M)
    for j in range(M):
        e[:] = 0.0
        e[j] = 1.0
        v.set(e.reshape(v.shape))
        kfill_(out, 0.0)
        mul_(out, JOp, v)
        dense[:, j] = out.numpy().reshape(-1)
    return dense


# ---------------------------------------------------------------------------------------------
# forcing: src/Ariadne.jl:164-217
# ---------------------------------------------------------------------------------------------
class Forcing:
    pass


class Fixed(Forcing):
    def __init__(self, eta=0.1):
        self.eta = float(eta)

    def __call__(self, *args):
        return self.eta


class EisenstatWalker(Forcing):
    def __init__(self, eta_max=0.999, gamma=0.9):
        self.eta_max, self.gamma = float(eta_max), float(gamma)

    def __call__(self, eta, tol, n_res, n_res_prior):
        # scalar host logic, evaluated by the library so that Julia / Python / C++ drivers agree bit for bit
        return L.load().ak_forcing_ew(self.eta_max, self.gamma, eta, tol, n_res, n_res_prior)


def inital(F):  # sic — the reference spells it this way (src/Ariadne.jl:192,217)
    return F.eta if isinstance(F, Fixed) else F.eta_max


Stats = namedtuple("Stats", "outer_iterations inner_iterations n_res")  # src/Ariadne.jl:265-269


def update(stats, inner_iterations, n_res):  # :270-276
    return Stats(stats.outer_iterations + 1, stats.inner_iterations + inner_iterations, float(n_res))


NewtonResult = namedtuple("NewtonResult", "solved stats t")

# ---------------------------------------------------------------------------------------------
# Krylov workspace: krylov_workspace / krylov_solve!  (src/Ariadne.jl:317-318,338-340,367)
# ---------------------------------------------------------------------------------------------
_ALGOS = {"gmres": A.AK_ALGO_GMRES, "cg": A.AK_ALGO_CG, "fgmres": A.AK_ALGO_FGMRES}
_FUSE = {"none": A.AK_FUSE_NONE, "mgs": A.AK_FUSE_MGS, "full": A.AK_FUSE_FULL, "pair": A.AK_FUSE_PAIR,
         "block4": A.AK_FUSE_BLOCK4, "block8": A.AK_FUSE_BLOCK8, "sweep": A.AK_FUSE_SWEEP}


class GmresPreconditioner:
    """struct GmresPreconditioner{JOp}; J::JOp; itmax::Int  (examples/bratu.jl:141-149, bvp.jl:29-38):
    `mul!(y, P, x) = copyto!(y, gmres(P.J, x; P.itmax)[1])`.  Passed as `N = J -> GmresPreconditioner(J, 5)`;
    the library runs the inner GMRES natively (AK_PRECOND_INNER_GMRES)."""

    kind, ldiv = A.AK_PRECOND_INNER_GMRES, False

    def __init__(self, J, itmax):
        self.J, self.itmax = J, int(itmax)


class TridiagonalLU:
    """What `ilu(collect(J))` is for the 1-D Bratu Jacobian (examples/bratu.jl:121-139): LU factors of a tridiagonal
    matrix have no fill-in, so the incomplete factorisation is the complete one.  Applied with `ldiv = true`
    (`ldiv!(y, P, x)`), natively as a partitioned Thomas solve on the device (AK_PRECOND_TRIDIAG_LU)."""

    kind, ldiv, itmax = A.AK_PRECOND_TRIDIAG_LU, True, 0

    def __init__(self, J):
        self.J = J


def ilu(J):
    """`ilu(collect(J))` of examples/bratu.jl:125,135 for a JacobianOperator whose matrix is tridiagonal (1-D Bratu).
    The O(N) JVPs of `collect` are not needed: the factors are formed from the stencil inside the solve kernel."""
    if not isinstance(J, JacobianOperator) or J.f.kind != A.AK_BRATU1D:
        raise NotImplementedError("ilu(J) is built natively for the tridiagonal 1-D Bratu Jacobian only")
    return TridiagonalLU(J)


class JacobiPreconditioner:
    """y = x ./ diag(J(u)) (AK_PRECOND_JACOBI): Bratu 1-D/2-D, heat 1-D/2-D.  A `mul!`-style object (ldiv = false)."""

    kind, ldiv, itmax = A.AK_PRECOND_JACOBI, False, 0

    def __init__(self, J):
        self.J = J


class UserPreconditioner:
    """Any preconditioner object of the caller: `apply_(y, x)` on torch.float64 CUDA tensors aliasing the library's
    vectors, run on the library's stream (AK_PRECOND_USER).  `ldiv` only records which of `mul!` / `ldiv!` the
    callable stands for (Krylov.jl's `ldiv` keyword); the library just calls it."""

    kind, itmax = A.AK_PRECOND_USER, 0

    def __init__(self, apply_, ldiv=False, device=0):
        self.apply_, self.ldiv, self._dev, self._n = apply_, bool(ldiv), device, None
        self._streams = {}
        self._cb = A.PRECOND_APPLY(self._call)

    def _call(self, _user, stream, x, y):
        try:
            import torch

            if stream not in self._streams:
                self._streams[stream] = torch.cuda.ExternalStream(stream, device=torch.device("cuda", self._dev))
            with torch.cuda.stream(self._streams[stream]):
                self.apply_(as_torch(y, (self._n,), self._dev), as_torch(x, (self._n,), self._dev))
            return 0
        except Exception:  # noqa: BLE001 - must not unwind through the C frames
            import traceback

            traceback.print_exc()
            return 1


def precond_apply_(y, P, x):
    """`mul!(y, P, x)` / `ldiv!(y, P, x)` for a native preconditioner object (what Krylov.jl calls on N(J) / M(J))."""
    if isinstance(P, UserPreconditioner):
        raise TypeError("a UserPreconditioner is applied by calling its own function")
    J = P.J
    prob = J.problem()
    L.check(x.ctx.lib.ak_precond_apply(x.ctx.h, C.byref(prob), _ptr(J.u), P.kind, P.itmax, _ptr(x), _ptr(y)))
    return y


def _precond_fields(P, n, ldiv, side):
    """(kind, itmax, apply pointer) of a preconditioner object for ak_krylov_opts."""
    if P is None:
        return A.AK_PRECOND_NONE, 0, None
    if not hasattr(P, "kind"):
        raise NotImplementedError(f"{side} = {P!r}: not a preconditioner this library can apply "
                                  "(GmresPreconditioner, ilu(J), JacobiPreconditioner, UserPreconditioner)")
    if bool(P.ldiv) != bool(ldiv):
        raise ValueError(f"{side}: {type(P).__name__} is applied with ldiv = {P.ldiv} (Krylov.jl keyword `ldiv`)")
    if isinstance(P, UserPreconditioner):
        P._n = n
        return P.kind, 0, C.cast(P._cb, C.c_void_p).value
    return P.kind, P.itmax, None


class KrylovConstructor:
    def __init__(self, res):
        self.proto = res


class KrylovStats:
    def __init__(self):
        self.niter, self.solved, self.inconsistent, self.breakdown, self.npass = 0, False, False, False, 0
        self.residuals, self.rnorm, self.beta, self.flags = [], 0.0, 0.0, 0


class KrylovWorkspace:
    def __init__(self, algo, kc, memory=20, max_basis=0):
        key = algo if isinstance(algo, str) else str(algo)
        key = key.lstrip(":")
        if key not in _ALGOS:
            raise ValueError(f"algo {algo!r} is not built natively (have: gmres, cg)")
        self.algo, self.proto = key, kc.proto
        self.ctx = kc.proto.ctx
        h = C.c_void_p()
        L.check(self.ctx.lib.ak_krylov_create(self.ctx.h, _ALGOS[key], kc.proto.n, memory, max_basis, C.byref(h)))
        self.h = h
        self.stats = KrylovStats()
        self.memory = memory

    @property
    def x(self):
        p = self.ctx.lib.ak_krylov_x(self.h)
        return DeviceVector(self.ctx, self.proto.shape, ptr=p, owner=self)

    def basis(self, i):
        """(stored basis vector i as a DeviceVector view, its scale): v_i = stored / scale (ak_krylov_basis)."""
        p, sc = C.c_void_p(), C.c_double()
        L.check(self.ctx.lib.ak_krylov_basis(self.h, int(i), C.byref(p), C.byref(sc), None))
        return DeviceVector(self.ctx, self.proto.shape, ptr=p.value, owner=self), sc.value

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.ak_krylov_destroy(self.h)
        except Exception:
            pass


def krylov_workspace(algo, kc, memory=20, max_basis=0):
    return KrylovWorkspace(algo, kc, memory=memory, max_basis=max_basis)


def krylov_solve_(workspace, J, b, atol=A.SQRT_EPS, rtol=A.SQRT_EPS, itmax=0, restart=False,
                  reorthogonalization=False, history=False, fuse="sweep", verbose=0, M=None, N=None, ldiv=False,
                  **unsupported):
    """krylov_solve!(workspace, J, b; kwargs...) — solves J x = b from x0 = 0.
    `M` / `N`: left / right preconditioner objects (GmresPreconditioner, ilu(J), JacobiPreconditioner,
    UserPreconditioner); `ldiv` as in Krylov.jl."""
    if unsupported:
        raise TypeError(f"krylov kwargs not supported on the native path: {sorted(unsupported)}")
    n = workspace.proto.n
    pn, pit, pfn = _precond_fields(N, n, ldiv, "N")
    pm, pmit, pmfn = _precond_fields(M, n, ldiv, "M")
    o = A.default_krylov_opts(atol=atol, rtol=rtol, itmax=itmax, restart=int(bool(restart)),
                              reorthogonalization=int(bool(reorthogonalization)), history=int(bool(history)),
                              fuse=_FUSE[fuse] if isinstance(fuse, str) else int(fuse), precond_n=pn,
                              precond_itmax=pit, precond_m=pm, precond_m_itmax=pmit, n_apply=pfn, m_apply=pmfn)
    st = A.ak_krylov_stats()
    prob = J.problem()
    ctx = workspace.ctx
    hist = None
    cap = 0
    if history:
        cap = int(itmax if itmax else 2 * workspace.proto.n) + 1
        cap = min(cap, 1 << 20)
        hist = np.zeros(cap)
    flags = L.check(ctx.lib.ak_krylov_solve(workspace.h, C.byref(prob), _ptr(J.u), _ptr(b), C.byref(o), C.byref(st),
                                            hist.ctypes.data_as(A.c_double_p) if history else None, cap))
    s = workspace.stats
    s.niter, s.solved, s.inconsistent = int(st.niter), bool(st.solved), bool(st.inconsistent)
    s.breakdown, s.npass, s.rnorm, s.beta, s.flags = bool(st.breakdown), int(st.npass), st.rnorm, st.beta, flags
    s.residuals = list(hist[: min(cap, st.niter + 1)]) if history else []
    return workspace


# ---------------------------------------------------------------------------------------------
# newton_krylov!: src/Ariadne.jl:288-372
# ---------------------------------------------------------------------------------------------
def _wants_coef(F_, jvp_mode):
    """Bratu: lambda*exp(u) cache; finite-difference JVPs: F(u) cache (include/ariadne_b200.h: ak_problem.coef)."""
    if F_.kind in (A.AK_BRATU1D, A.AK_BRATU2D):
        return True
    if isinstance(F_, UserResidual) and F_.jvp_ is None:
        return True
    return jvp_mode in ("fd", "fd2")


def newton_krylov_(F_, u, p=None, res=None, *, tol_rel=1.0e-6, tol_abs=1.0e-12, max_niter=50,
                   forcing=EisenstatWalker(), verbose=0, algo="gmres", M=None, N=None, krylov_kwargs=None,
                   callback=lambda *args: None, memory=20, max_basis=0, history=None, workspace=None,
                   jvp_mode="analytic"):
    """Newton loop driven from the host language, one C-ABI call per arrowed line of
    src/Ariadne.jl:288-372.  Returns `(u, NewtonResult(solved, stats, t))`.

    `history` (optional list) receives one dict per Newton iteration
    (n_res, inner iterations, eta used) — not in the reference; used by the parity tests."""
    krylov_kwargs = dict(krylov_kwargs or {})
    if res is None:  # 3-argument form: res = similar(u0, M); make_zero!(res)   :259-263
        res = u.zero()
    ctx = u.ctx
    lib = ctx.lib
    t0 = time.perf_counter_ns()
    # Bratu: lambda*exp(u) is cached by the residual kernel for the JVPs of the same Newton step
    # (finite-difference JVPs: F(u) is cached instead)
    coef = u.similar() if _wants_coef(F_, jvp_mode) else None
    prob = F_.problem(u, p, coef=coef)
    if jvp_mode == "fd2" or (jvp_mode == "fd" and prob.kind not in (A.AK_BRATU1D, A.AK_BRATU2D)):
        prob.jvp_mode = A.AK_JVP_FD
    nrm = C.c_double()

    def residual_norm():
        L.check(lib.ak_residual(ctx.h, C.byref(prob), _ptr(u), _ptr(res), C.byref(nrm)))  # F!(res,u,p); norm(res)
        return nrm.value

    n_res = residual_norm()                                  # :302-303
    callback(u, res, n_res)                                  # :304
    if history is not None:
        history.append(dict(n_res=n_res, inner=0, eta=None))
    tol = tol_rel * n_res + tol_abs                          # :306
    eta = inital(forcing) if forcing is not None else None   # :308-310
    if verbose > 0:
        print(f"[ Info: Jacobian-Free Newton-Krylov algo={algo} res0={n_res} tol={tol} eta={eta}")
    J = JacobianOperator(F_, res, u, p, coef=coef, jvp_mode=jvp_mode)  # :314
    if workspace is None:
        workspace = krylov_workspace(algo, KrylovConstructor(res), memory=memory, max_basis=max_basis)  # :317-318
    rhs = res.similar()
    stats = Stats(0, 0, n_res)                               # :320
    while n_res > tol and stats.outer_iterations <= max_niter:   # :321
        kwargs = dict(krylov_kwargs)
        if N is not None:
            kwargs = {"N": N(J), **kwargs}                   # kwargs = (; N = N(J), kwargs...)  :324-326
        if M is not None:
            kwargs = {"M": M(J), **kwargs}                   # kwargs = (; M = M(J), kwargs...)  :327-329
        if forcing is not None:
            kwargs = {"rtol": eta, **kwargs}                 # later keys win  :330-333
        kcopy_(len(res), rhs, res)                           # copy(res)      :338
        krylov_solve_(workspace, J, rhs, **kwargs)
        d = workspace.x                                      # :340
        s = 1                                                # :341
        kaxpy_(len(u), -float(s), d, u)                      # u .-= s .* d   :344
        n_res_prior = n_res
        n_res = residual_norm()                              # :349-350
        callback(u, res, n_res)                              # :351
        eta_used = eta
        if math.isinf(n_res) or math.isnan(n_res):           # :353-356
            print(f"[ Error: Inner solver blew up stats={stats}")
            break
        if forcing is not None:
            eta = forcing(eta, tol, n_res, n_res_prior)      # :358-360
        if verbose > 0 and workspace.stats.niter == 0 and forcing is not None:
            print(f"[ Info: Inexact Newton thinks our step is good enough eta={eta} stats={stats}")
        stats = update(stats, workspace.stats.niter, n_res)  # :367
        if history is not None:
            history.append(dict(n_res=n_res, inner=workspace.stats.niter, eta=eta_used))
        if verbose > 0:
            print(f"[ Info: Newton iter={n_res} eta={eta} stats={stats}")
    t = (time.perf_counter_ns() - t0) / 1.0e9
    return u, NewtonResult(n_res <= tol, stats, t)


def newton_krylov(F, u0, p=None, M=None, **kwargs):
    """Out-of-place form, src/Ariadne.jl:245-248: `F!(res, u, p) = (res .= F(u, p))`.  `F` is one of the
    native residuals, a UserResidual, or any callable `F(u, p) -> tensor` on torch CUDA tensors (wrapped with
    UserResidual.from_function: tangent by forward-mode AD of the same code)."""
    if not isinstance(F, NativeResidual):
        if not callable(F):
            raise TypeError("newton_krylov: F must be callable")
        F = UserResidual.from_function(F)
    return newton_krylov_(F, u0, p, None, **kwargs)


def _newton_opts(tol_rel, tol_abs, max_niter, forcing, algo, memory, max_basis, krylov_kwargs, verbose=0, N=None,
                 M=None, J=None, keep=None, native_loop=False):
    kk = dict(krylov_kwargs or {})
    override = "rtol" in kk
    fuse = kk.pop("fuse", "sweep")
    ldiv = bool(kk.pop("ldiv", False))
    n = len(J.u) if J is not None else 0
    def built_once(P, side):
        # The reference calls N(J) / M(J) before EVERY linear solve (src/Ariadne.jl:324-329).  The native kinds read u
        # at apply time, so building them once is equivalent; a caller-supplied object may have captured state of u0
        # (a factorisation of collect(J), a shift ...), which the C++ loop would silently freeze.
        if isinstance(P, UserPreconditioner) and native_loop:
            raise NotImplementedError(
                f"{side} = J -> UserPreconditioner(...) with the C++ Newton loop (ak_newton_solve): the object would be "
                "built once at u0, the reference rebuilds it every Newton step; use newton_krylov_ (host-driven loop)")
        return P

    if N is not None:
        P = built_once(N(J), "N")
        kk["precond_n"], kk["precond_itmax"], kk["n_apply"] = _precond_fields(P, n, ldiv, "N")
        if keep is not None:
            keep.append(P)
    if M is not None:
        P = built_once(M(J), "M")
        kk["precond_m"], kk["precond_m_itmax"], kk["m_apply"] = _precond_fields(P, n, ldiv, "M")
        if keep is not None:
            keep.append(P)
    ko = A.default_krylov_opts(fuse=_FUSE[fuse] if isinstance(fuse, str) else int(fuse),
                               **{k: (int(v) if isinstance(v, bool) else v) for k, v in kk.items()})
    o = A.default_newton_opts(tol_rel=tol_rel, tol_abs=tol_abs, max_niter=max_niter,
                              algo=_ALGOS[algo.lstrip(":")], memory=memory, max_basis=max_basis, verbose=verbose)
    o.krylov = ko
    o.krylov_rtol_override = int(override)
    if forcing is None:
        o.forcing = A.AK_FORCING_NONE
    elif isinstance(forcing, Fixed):
        o.forcing, o.eta = A.AK_FORCING_FIXED, forcing.eta
    else:
        o.forcing, o.eta_max, o.gamma = A.AK_FORCING_EW, forcing.eta_max, forcing.gamma
    return o


def newton_krylov_native_(F_, u, p=None, res=None, *, tol_rel=1.0e-6, tol_abs=1.0e-12, max_niter=50,
                          forcing=EisenstatWalker(), algo="gmres", krylov_kwargs=None, memory=20, max_basis=0,
                          history=None, verbose=0, N=None, M=None):
    """Same solve through the single C entry point ak_newton_solve (the loop runs in C++)."""
    if res is None:
        res = u.zero()
    ctx = u.ctx
    coef = u.similar() if _wants_coef(F_, "analytic") else None
    prob = F_.problem(u, p, coef=coef)
    keep = []  # preconditioner objects (their ctypes callbacks) must outlive the solve
    o = _newton_opts(tol_rel, tol_abs, max_niter, forcing, algo, memory, max_basis, krylov_kwargs, verbose, N=N, M=M,
                     J=JacobianOperator(F_, res, u, p, coef=coef), keep=keep, native_loop=True)
    st = A.ak_newton_stats()
    cap = max_niter + 3
    hn, hi, he = np.zeros(cap), np.zeros(cap, dtype=np.int64), np.zeros(cap)
    L.check(ctx.lib.ak_newton_solve(ctx.h, C.byref(prob), _ptr(u), _ptr(res), C.byref(o), C.byref(st),
                                    hn.ctypes.data_as(A.c_double_p), hi.ctypes.data_as(A.c_int64_p),
                                    he.ctypes.data_as(A.c_double_p), cap, C.cast(None, A.NEWTON_CALLBACK), None))
    if history is not None:
        for i in range(st.outer_iterations + 1):
            history.append(dict(n_res=hn[i], inner=int(hi[i]), eta=he[i] if i else None))
    stats = Stats(int(st.outer_iterations), int(st.inner_iterations), st.n_res)
    return u, NewtonResult(bool(st.solved), stats, st.t_seconds)


# ---------------------------------------------------------------------------------------------
# implicit time stepping: examples/implicit.jl:41-78
# ---------------------------------------------------------------------------------------------
def jacobian(G_, f_, un, p, dt, t):
    """jacobian(G!, f!, u_n, p, dt, t): implicit.jl:41-50 — dense collect(J)."""
    u = un.copy()
    du = un.zero()
    res = un.zero()
    F_ = ImplicitResidual(G_, f_)
    return collect(JacobianOperator(F_, res, u, (un, dt, du, p, t)))


def solve(G_, f_, un, p, dt, ts, callback=lambda u: None, verbose=0, algo="gmres", krylov_kwargs=None,
          step_stats=None):
    """solve(G!, f!, u_n, p, dt, ts; ...): implicit.jl:54-78."""
    u = un.copy()
    du = un.zero()
    res = un.zero()
    F_ = ImplicitResidual(G_, f_)
    ts = list(ts)
    for t in ts:
        if t == ts[0]:
            continue
        _, stats = newton_krylov_(F_, u, (un, dt, du, p, t), res, verbose=verbose, algo=algo, tol_abs=6.0e-6,
                                  krylov_kwargs=krylov_kwargs)
        if not stats.solved:
            print(f"[ Warning: non linear solve failed marching on t={t} stats={stats}")
        if step_stats is not None:
            step_stats.append(stats)
        callback(u)
        kcopy_(len(u), un, u)  # u_n .= u
    return un
