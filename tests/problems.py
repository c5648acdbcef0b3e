"""Shared problem definitions for the parity tests: the same inputs go to the oracle (host
pointers) and to the CUDA library (device vectors).  Inputs follow SURVEY.md §8d."""
import numpy as np

from newtonkrylov_jl_b200 import _abi as A


def bratu1d(N, lam=3.5):
    """examples/bratu.jl:40-46: dx = 1/(N+1), x = LinRange(dx, 1-dx, N), u0 = sin(pi x)."""
    dx = 1.0 / (N + 1)
    x = np.linspace(dx, 1.0 - dx, N)
    return dict(kind=A.AK_BRATU1D, nx=N, ny=1, dx=dx, lam=lam, u0=np.sin(np.pi * x), x=x)


def bratu2d(nx, ny=None, lam=3.5):
    """2-D Bratu (defined by this repo): dx = 1/(nx+1), u0 = sin(pi x) sin(pi y)."""
    ny = ny or nx
    dx, dy = 1.0 / (nx + 1), 1.0 / (ny + 1)
    x = dx * np.arange(1, nx + 1)
    y = dy * np.arange(1, ny + 1)
    u0 = np.sin(np.pi * y)[:, None] * np.sin(np.pi * x)[None, :]
    return dict(kind=A.AK_BRATU2D, nx=nx, ny=ny, dx=dx, dy=dy, lam=lam, u0=u0)


def heat1d(M, a=0.2, dt=0.1, bc=A.AK_BC_ZERO):
    """examples/heat_1D.jl:96-115: dx = 1/(M+1), xs = 0:dx:1 (M+2 points incl. boundaries), IC 4x(1-x)."""
    dx = 1.0 / (M + 1)
    xs = dx * np.arange(0, M + 2)
    return dict(kind=A.AK_HEAT1D, nx=M + 2, ny=1, dx=dx, a=a, dt=dt, bc=bc, scheme=A.AK_EULER,
                u0=4.0 * xs * (1.0 - xs), x=xs)


def heat2d(N, a=0.01, dt_scale=1.0, bc=A.AK_BC_ZERO, ic="sin"):
    """examples/heat_2D.jl:64-91: dx = dy = 1/(N+1), dt = dx^2 dy^2 / (2 a (dx^2+dy^2))."""
    dx = dy = 1.0 / (N + 1)
    dt = dt_scale * dx**2 * dy**2 / (2.0 * a * (dx**2 + dy**2))
    x = dx * np.arange(1, N + 1)
    X, Y = x[None, :], x[:, None]
    if ic == "sin":
        u0 = np.sin(np.pi * X) * np.sin(np.pi * Y)
    else:  # not an eigenfunction of the discrete Laplacian
        u0 = 16.0 * X * (1 - X) * Y * (1 - Y)
    return dict(kind=A.AK_HEAT2D, nx=N, ny=N, dx=dx, dy=dy, a=a, dt=dt, bc=bc, scheme=A.AK_EULER, u0=u0)


LGL = np.array([-1.0, -1.0 / np.sqrt(5.0), 1.0 / np.sqrt(5.0), 1.0])


def heat1d_dg(ne, dt=0.01):
    """examples/heat_1D_DG.jl:17-40: ne elements x 4 LGL nodes on [0,1], periodic, IC sin(pi x)."""
    h = 1.0 / ne
    x = (np.arange(ne)[:, None] * h + (LGL[None, :] + 1.0) * h / 2.0).reshape(-1)
    return dict(kind=A.AK_HEAT1D_DG, nx=4 * ne, ny=1, dx=h, dt=dt, bc=A.AK_BC_PERIODIC, scheme=A.AK_EULER,
                u0=np.sin(np.pi * x), x=x)


def oracle_problem(O, d, un=None):
    return O.make_problem(d["kind"], d["nx"], d["ny"], bc=d.get("bc", A.AK_BC_ZERO), scheme=d.get("scheme", A.AK_STEADY),
                          dx=d.get("dx", 0.0), dy=d.get("dy", 0.0), lam=d.get("lam", 0.0), a=d.get("a", 0.0),
                          dt=d.get("dt", 0.0), un=un)


def device_setup(nk, ctx, d):
    """(F_, u, p, un) for the host mirror: native residual object, device state, parameter tuple."""
    k = d["kind"]
    u = nk.DeviceVector.from_numpy(d["u0"], ctx)
    bc = nk.bc_zero_ if d.get("bc", A.AK_BC_ZERO) == A.AK_BC_ZERO else nk.bc_periodic_
    if k == A.AK_BRATU1D:
        return nk.bratu_, u, (d["dx"], d["lam"]), None
    if k == A.AK_BRATU2D:
        return nk.bratu2d_, u, (d["dx"], d["dy"], d["lam"]), None
    un = nk.DeviceVector.from_numpy(d["u0"], ctx)
    du = un.zero()
    if k == A.AK_HEAT1D:
        F_, pin = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_), (d["a"], d["dx"], bc)
    elif k == A.AK_HEAT2D:
        F_, pin = nk.ImplicitResidual(nk.G_Euler_, nk.diffusion_), (d["a"], d["dx"], d["dy"], bc)
    else:
        F_, pin = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_DG_), (d["dx"],)
    return F_, u, (un, d["dt"], du, pin, 0.0), un


def ulp_diff(a, b):
    """max |a-b| in units of the spacing at max(|a|,|b|)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    sp = np.spacing(np.maximum(np.abs(a), np.abs(b)))
    return float(np.max(np.abs(a - b) / sp)) if a.size else 0.0


def generic(d):
    """Same problem with a non-symmetric smooth perturbation of the initial guess.  The reference's
    own IC (sin(pi x), sin(pi x) sin(pi y)) is an even function on a symmetric grid: the Krylov spaces
    it generates are rank-deficient in exact arithmetic and late GMRES iterations are driven by
    rounding noise, so histories are reproducible only to ~1e-2 there (measured on the oracle itself,
    see tests/test_gpu_solvers.py::oracle_sensitivity)."""
    d = dict(d)
    if d["kind"] == A.AK_BRATU1D:
        x = d["x"]
        d["u0"] = d["u0"] + 1.2 * x * (1 - x) ** 2 * np.exp(x)
    else:
        nx, ny = d["nx"], d["ny"]
        X = (d["dx"] * np.arange(1, nx + 1))[None, :]
        Y = (d["dy"] * np.arange(1, ny + 1))[:, None]
        d["u0"] = d["u0"] + 2.4 * X * (1 - X) ** 2 * Y**2 * (1 - Y) * np.exp(X + 0.5 * Y)
    return d
