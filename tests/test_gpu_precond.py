"""GPU parity tests of the preconditioner hooks `M` / `N` of newton_krylov! (src/Ariadne.jl:296-297,324-329)
against the CPU oracle: the native tridiagonal LU that stands for `ilu(collect(J))` (examples/bratu.jl:121-139),
Jacobi, caller-supplied apply callbacks, left preconditioning."""
import numpy as np
import pytest

from newtonkrylov_jl_b200 import _abi as A
import problems as P

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(11)
EPS = 2.220446049250313e-16


def rel(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300))


def setup(nk, ctx, d):
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    res = u.zero()
    return nk.JacobianOperator(F_, res, u, p), u, res


# block boundaries of the partitioned Thomas solve: kBlk = 32 rows, serial below 2049 unknowns, three levels above 65536
@pytest.mark.parametrize("N", [1, 2, 7, 32, 33, 1000, 2048, 2049, 2081, 5000, 65537, 70001, (1 << 20) + 3])
def test_tridiagonal_lu_solve_matches_thomas(nk, ctx, oracle, N):
    """ldiv!(y, ilu(collect(J)), x) for the 1-D Bratu Jacobian: the partitioned device solve against the oracle's
    sequential Thomas algorithm.  Both are backward stable; the solutions agree to cond(J) * eps ~ N^2 eps."""
    d = P.bratu1d(N)
    J, u, res = setup(nk, ctx, d)
    po = P.oracle_problem(oracle, d)
    y_true = np.sin(3.0 * np.pi * d["x"]) + 0.1 * RNG.standard_normal(N)
    x0, _ = oracle.jvp(po, d["u0"], y_true)
    x = nk.DeviceVector.from_numpy(x0, ctx)
    y = x.zero()
    nk.precond_apply_(y, nk.ilu(J), x)
    yr = oracle.precond_apply(po, d["u0"], A.AK_PRECOND_TRIDIAG_LU, x0)
    tol = max(1e-12, 20.0 * EPS * float(N) ** 2)
    assert rel(y.numpy(), yr) <= tol
    assert rel(y.numpy(), y_true) <= tol
    # residual of the device solution, evaluated by the oracle's operator
    r, _ = oracle.jvp(po, d["u0"], y.numpy())
    assert np.linalg.norm(r - x0) <= 1e-9 * np.linalg.norm(x0)


def test_tridiagonal_lu_uses_cached_coefficient(nk, ctx, oracle):
    """With the lambda*exp(u) cache of the residual kernel (ak_problem.coef) the solve reads no exp."""
    d = P.bratu1d(3000)
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    res, coef = u.zero(), u.zero()
    prob = F_.problem(u, p, coef=coef)
    import ctypes as C
    nk._lib.check(ctx.lib.ak_residual(ctx.h, C.byref(prob), C.c_void_p(u.ptr), C.c_void_p(res.ptr), None))
    x0 = RNG.standard_normal(3000)
    x = nk.DeviceVector.from_numpy(x0, ctx)
    y1, y2 = x.zero(), x.zero()
    nk.precond_apply_(y1, nk.ilu(nk.JacobianOperator(F_, res, u, p, coef=coef)), x)
    nk.precond_apply_(y2, nk.ilu(nk.JacobianOperator(F_, res, u, p)), x)
    assert rel(y1.numpy(), y2.numpy()) < 1e-9


@pytest.mark.parametrize("algo", ["gmres", "fgmres"])
def test_gmres_with_ilu_converges_in_one_iteration(nk, ctx, oracle, algo):
    d = P.bratu1d(1000)
    J, u, res = setup(nk, ctx, d)
    b0 = RNG.standard_normal(1000)
    b = nk.DeviceVector.from_numpy(b0, ctx)
    ws = nk.krylov_workspace(algo, nk.KrylovConstructor(res))
    nk.krylov_solve_(ws, J, b, N=nk.ilu(J), ldiv=True, history=True)
    po = P.oracle_problem(oracle, d)
    xr, sr, hr = oracle.krylov_solve(po, d["u0"], b0, algo=A.AK_ALGO_GMRES if algo == "gmres" else A.AK_ALGO_FGMRES,
                                     precond_n=A.AK_PRECOND_TRIDIAG_LU, hist_cap=10)
    assert (ws.stats.niter, ws.stats.solved) == (sr["niter"], sr["solved"]) == (1, True)
    assert rel(ws.x.numpy(), xr) < 1e-8


def test_ilu_requires_ldiv(nk, ctx):
    d = P.bratu1d(64)
    J, u, res = setup(nk, ctx, d)
    ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res))
    with pytest.raises(ValueError):
        nk.krylov_solve_(ws, J, u.copy(), N=nk.ilu(J))  # examples/bratu.jl:126 passes krylov_kwargs = (; ldiv = true)


CASES = [("bratu2d", lambda: P.bratu2d(24)), ("bratu1d", lambda: P.bratu1d(150)),
         ("heat2d", lambda: P.heat2d(24, dt_scale=64.0, ic="poly")), ("heat1d", lambda: P.heat1d(100))]


@pytest.mark.parametrize("side", ["N", "M", "MN"])
@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_jacobi_left_and_right_match_oracle(nk, ctx, oracle, name, make, side):
    d = make()
    J, u, res = setup(nk, ctx, d)
    b0 = RNG.standard_normal(d["u0"].shape)
    if d["kind"] == A.AK_HEAT1D:
        b0[0] = b0[-1] = 0.0
    b = nk.DeviceVector.from_numpy(b0, ctx)
    kw, okw = {}, {}
    if "N" in side:
        kw["N"], okw["precond_n"] = nk.JacobiPreconditioner(J), A.AK_PRECOND_JACOBI
    if "M" in side:
        kw["M"], okw["precond_m"] = nk.JacobiPreconditioner(J), A.AK_PRECOND_JACOBI
    ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res))
    nk.krylov_solve_(ws, J, b, rtol=1e-9, history=True, **kw)
    po = P.oracle_problem(oracle, d, un=d["u0"] if d.get("scheme") else None)
    xr, sr, hr = oracle.krylov_solve(po, d["u0"], b0, rtol=1e-9, hist_cap=100000, **okw)
    assert (ws.stats.niter, ws.stats.solved, ws.stats.npass) == (sr["niter"], sr["solved"], sr["npass"])
    assert rel(ws.x.numpy(), xr) < 1e-8
    assert np.max(np.abs(np.array(ws.stats.residuals) - hr)) <= 1e-9 * hr[0]


def test_user_preconditioner_callback(nk, ctx, oracle):
    """Any preconditioner object of the caller: here the inverse diagonal written with torch ops on the library's
    stream.  Same iteration as the native Jacobi kernel."""
    import torch

    d = P.bratu2d(24)
    J, u, res = setup(nk, ctx, d)
    b0 = RNG.standard_normal(d["u0"].shape)
    diag = torch.as_tensor((-2.0 / d["dx"] ** 2 - 2.0 / d["dy"] ** 2) + d["lam"] * np.exp(d["u0"].reshape(-1)), device="cuda")
    calls = []

    def apply_(y, x):
        calls.append(1)
        torch.div(x, diag, out=y)

    hist = {}
    for tag, N in (("user", nk.UserPreconditioner(apply_)), ("native", nk.JacobiPreconditioner(J))):
        ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res))
        nk.krylov_solve_(ws, J, nk.DeviceVector.from_numpy(b0, ctx), N=N, rtol=1e-9, history=True)
        hist[tag] = (ws.stats.niter, np.array(ws.stats.residuals), ws.x.numpy())
    assert hist["user"][0] == hist["native"][0] and len(calls) >= hist["user"][0]
    assert np.max(np.abs(hist["user"][1] - hist["native"][1])) <= 1e-10 * hist["native"][1][0]
    assert rel(hist["user"][2], hist["native"][2]) < 1e-10


@pytest.mark.parametrize("algo", ["gmres", "fgmres"])
@pytest.mark.parametrize("drive", ["host", "native"])
def test_newton_with_ilu_bratu_example(nk, ctx, oracle, algo, drive):
    """examples/bratu.jl:121-139: `newton_krylov!(bratu!, u0, (dx, lambda), res; algo = :gmres | :fgmres,
    N = (J) -> ilu(collect(J)), krylov_kwargs = (; ldiv = true))` with N = 10_000, lambda = 3.51382."""
    d = P.bratu1d(10000, lam=3.51382)
    po = P.oracle_problem(oracle, d)
    o = A.default_newton_opts(algo=A.AK_ALGO_GMRES if algo == "gmres" else A.AK_ALGO_FGMRES)
    o.krylov.precond_n = A.AK_PRECOND_TRIDIAG_LU
    ur, sr, hr = oracle.newton(po, d["u0"], o)
    assert sr["solved"]
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    hist = []
    fn = nk.newton_krylov_ if drive == "host" else nk.newton_krylov_native_
    _, r = fn(F_, u, p, u.zero(), algo=algo, N=lambda J: nk.ilu(J), krylov_kwargs=dict(ldiv=True), history=hist)
    assert r.solved and r.stats.outer_iterations == sr["outer_iterations"]
    assert [h["inner"] for h in hist] == [h["inner"] for h in hr]
    for a, b in zip(hist, hr):
        # exact Newton: quadratic convergence, compare down to the rounding floor of ||F|| (cond(J) eps ||F0||)
        assert abs(a["n_res"] - b["n_res"]) <= 1e-8 * b["n_res"] + 1e-7 * hr[0]["n_res"] * 1e-3
    assert rel(u.numpy(), ur) < 1e-8
    theta = 4.79173  # examples/bratu.jl:33-37
    ref = -2.0 * np.log(np.cosh(theta * (d["x"] - 0.5) / 2.0) / np.cosh(theta / 4.0))
    assert np.max(np.abs(u.numpy() - ref)) < 1e-4


def test_full_bratu2d_solve_with_caller_supplied_fast_poisson_preconditioner(nk, ctx, oracle):
    """The `N` hook with user code (AK_PRECOND_USER): y = (Laplacian + mean(lambda e^u))^-1 x by sine transforms
    (examples/python/bratu2d_fast_poisson.py).  With it the 2-D Bratu solve converges in a handful of GMRES
    iterations per Newton step at any grid size; against the oracle driven with the same preconditioner (scipy DST)."""
    import importlib.util
    import os

    import scipy.fft as sfft

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("bratu2d_fast_poisson",
                                                  os.path.join(here, "..", "examples", "python", "bratu2d_fast_poisson.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    N, lam = 192, 3.5
    u, r, hist, _ = ex.solve(N, lam, verbose=False, ctx=ctx)
    assert r.solved and all(h["inner"] <= 12 for h in hist)

    # oracle: same Newton loop, the preconditioner rebuilt per Newton step like N(J)  (src/Ariadne.jl:324-326)
    d = P.bratu2d(N, lam=lam)
    po = P.oracle_problem(oracle, d)
    k = np.arange(1, N + 1)
    mu = -4.0 * np.sin(np.pi * k / (2 * (N + 1))) ** 2 / d["dx"] ** 2
    eig = mu[:, None] + mu[None, :]
    state = {"shift": 0.0}

    def apply(y, x):
        X = sfft.dstn(x.reshape(N, N), type=1) / (eig + state["shift"])
        y[:] = (sfft.idstn(X, type=1)).reshape(-1)

    fn, keep = oracle.user_precond(N * N, apply)
    uo = d["u0"].copy()
    n_res = np.linalg.norm(oracle.residual(po, uo)[0])
    tol, eta, hist_o = 1e-6 * n_res + 1e-12, 0.999, []
    while n_res > tol and len(hist_o) <= 50:
        state["shift"] = lam * float(np.mean(np.exp(uo)))
        res, _ = oracle.residual(po, uo)
        x, st, _ = oracle.krylov_solve(po, uo, res, rtol=eta, precond_n=A.AK_PRECOND_USER, n_apply=fn)
        uo = uo - x.reshape(uo.shape)
        prior, n_res = n_res, np.linalg.norm(oracle.residual(po, uo)[0])
        eta = oracle.forcing_ew(0.999, 0.9, eta, tol, n_res, prior)
        hist_o.append(dict(n_res=n_res, inner=st["niter"]))
    assert [h["inner"] for h in hist[1:]] == [h["inner"] for h in hist_o]
    for a, b in zip(hist[1:], hist_o):
        assert abs(a["n_res"] - b["n_res"]) <= 1e-7 * b["n_res"] + 1e-11 * hist[0]["n_res"]
    assert rel(u.numpy(), uo) < 1e-8
    del keep


def test_user_preconditioner_is_refused_by_the_cpp_newton_loop(nk, ctx):
    """The reference rebuilds N(J) before every linear solve (src/Ariadne.jl:324-329); the single-call C++ loop builds
    the object once, which is only equivalent for the native kinds — a caller-supplied object is refused there."""
    d = P.bratu2d(12)
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    with pytest.raises(NotImplementedError):
        nk.newton_krylov_native_(F_, u, p, None, N=lambda J: nk.UserPreconditioner(lambda y, x: y.copy_(x)))
    # the host-driven loop takes it (and calls the hook once per Newton step)
    built = []
    _, r = nk.newton_krylov_(F_, u, p, None, N=lambda J: built.append(1) or nk.UserPreconditioner(lambda y, x: y.copy_(x)))
    assert r.solved and len(built) == r.stats.outer_iterations


def test_preconditioned_cg_matches_oracle(nk, ctx, oracle):
    """cg! with kwarg M (Krylov.jl: z = M r, gamma = <r, z>, residual norm in the M-norm).  J of 1-D Bratu is negative
    definite and CG is applied to it as it is (like the reference does at examples/bratu.jl:59-108); M = -1 / diag(J) is
    symmetric positive definite.  A right preconditioner (N) does not exist for cg!: refused."""
    import torch

    d = P.bratu1d(12000, lam=1.0)          # above the one-block regime: the multi-kernel path
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    res = u.zero()
    J = nk.JacobianOperator(F_, res, u, p)
    diag = -2.0 / d["dx"] ** 2 + d["lam"] * np.exp(d["u0"])
    minv = torch.as_tensor(-1.0 / diag, device="cuda")
    b0 = RNG.standard_normal(d["nx"])
    kw = dict(rtol=1e-30, atol=0.0, itmax=200)
    ws = nk.krylov_workspace("cg", nk.KrylovConstructor(res))
    nk.krylov_solve_(ws, J, nk.DeviceVector.from_numpy(b0, ctx), history=True,
                     M=nk.UserPreconditioner(lambda y, x: torch.mul(x, minv, out=y)), **kw)
    mh = -1.0 / diag

    def apply(y, x):
        y[:] = x * mh

    fn, keep = oracle.user_precond(d["nx"], apply)
    po = P.oracle_problem(oracle, d)
    xr, sr, hr = oracle.krylov_solve(po, d["u0"], b0, algo=A.AK_ALGO_CG, hist_cap=1000, precond_m=A.AK_PRECOND_USER,
                                     m_apply=fn, **kw)
    st = ws.stats
    assert (st.niter, st.solved) == (sr["niter"], sr["solved"]) and st.niter == 200
    assert hr[-1] < 0.5 * hr[0]  # Jacobi does something on this operator
    assert np.max(np.abs(np.array(st.residuals) - hr)) <= 1e-9 * hr[0]
    assert rel(ws.x.numpy(), xr) < 1e-8
    with pytest.raises(nk.AriadneError):
        nk.krylov_solve_(ws, J, nk.DeviceVector.from_numpy(b0, ctx), N=nk.JacobiPreconditioner(J), **kw)
