"""GPU parity tests of the generic residual seam (AK_USER): caller-supplied `F!(res, u, p)` and tangent callbacks
through the C ABI, against the CPU oracle driven with the same callbacks on host arrays.

Reference cases: the 2x2 system of test/runtests.jl:4-7 / examples/simple.jl:6-9 written as user code, and the
Kelley boundary-value problem of examples/bvp.jl:10-60 (FGMRES + GmresPreconditioner(J, 30))."""
import numpy as np
import pytest

from newtonkrylov_jl_b200 import _abi as A
import problems as P

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300))


# ---- examples/simple.jl:6-9 as user code ---------------------------------------------------------------------
def simple_torch(x, p):
    import torch

    return torch.stack([x[0] ** 2 + x[1] ** 2 - 2.0, torch.exp(x[0] - 1.0) + x[1] ** 2 - 2.0])


def simple_np(res, x):
    res[0] = x[0] ** 2 + x[1] ** 2 - 2.0
    res[1] = np.exp(x[0] - 1.0) + x[1] ** 2 - 2.0


def simple_jvp_np(out, x, v):
    out[0] = 2.0 * x[0] * v[0] + 2.0 * x[1] * v[1]
    out[1] = np.exp(x[0] - 1.0) * v[0] + 2.0 * x[1] * v[1]


@pytest.mark.parametrize("x0", [[2.0, 0.5], [3.0, 5.0]])
def test_simple_system_as_user_code(nk, ctx, oracle, x0):
    """test/runtests.jl:15-23 through the generic seam: out-of-place F, tangent by forward-mode AD (torch.func.jvp,
    the analogue of the reference's Enzyme call); same Newton / GMRES counts as the native kernel and the oracle."""
    hist = []
    u = nk.DeviceVector.from_numpy(np.array(x0), ctx)
    _, r = nk.newton_krylov(simple_torch, u, None, history=hist)
    assert r.solved
    # native 2x2 kernel
    hist_n = []
    un = nk.DeviceVector.from_numpy(np.array(x0), ctx)
    _, rn = nk.newton_krylov_(nk.simple_F_, un, None, history=hist_n)
    assert [h["inner"] for h in hist] == [h["inner"] for h in hist_n]
    assert rel(u.numpy(), un.numpy()) < 1e-12
    # oracle with the same callbacks on host arrays
    po = oracle.make_user_problem(2, simple_np, simple_jvp_np)
    ur, sr, hr = oracle.newton(po, np.array(x0))
    assert r.stats.outer_iterations == sr["outer_iterations"] and r.stats.inner_iterations == sr["inner_iterations"]
    for a, b in zip(hist, hr):
        assert abs(a["n_res"] - b["n_res"]) <= 1e-10 * hr[0]["n_res"]
    assert rel(u.numpy(), ur) < 1e-10


def test_jacobian_operator_known_answer_user_code(nk, ctx):
    """test/runtests.jl:36-38 with F supplied as user code: J([3,5]) * [1,0] == [6.0, 7.38905609893065]."""
    F_ = nk.UserResidual.from_function(simple_torch)
    u = nk.DeviceVector.from_numpy(np.array([3.0, 5.0]), ctx)
    res = u.zero()
    J = nk.JacobianOperator(F_, res, u, None)
    assert J.size() == (2, 2) and len(J) == 4
    v = nk.DeviceVector.from_numpy(np.array([1.0, 0.0]), ctx)
    out = u.zero()
    nk.mul_(out, J, v)
    assert np.allclose(out.numpy(), [6.0, 7.38905609893065], rtol=1e-15, atol=0)


# ---- examples/bvp.jl ---------------------------------------------------------------------------------------------
def bvp_setup(n=801):
    h = 20.0 / (n - 1)
    tv = h * np.arange(n)
    tvdag = tv.copy()
    tvdag[1:] = 1.0 / tv[1:]
    U0 = np.zeros(2 * n)
    U0[0::2] = np.exp(-0.1 * tv * tv)          # BVP_U0!  bvp.jl:25-28
    U0[1::2] = -0.2 * U0[0::2] * tv
    return n, h, tv, tvdag, U0


def bvp_np(n, h, tv, tvdag):
    h2 = 0.5 * h

    def F(res, U):  # Fbvp!  bvp.jl:10-23
        v, vp = U[0::2], U[1::2]
        force = 4.0 * tvdag * vp + (tv * v - 1.0) * v
        res[0] = U[1]
        res[2 * n - 1] = U[2 * n - 2]
        res[2:2 * n - 1:2] = v[1:] - v[:-1] - h2 * (vp[:-1] + vp[1:])
        res[1:2 * n - 2:2] = vp[1:] - vp[:-1] + h2 * (force[:-1] + force[1:])

    def jvp(out, U, dU):
        v = U[0::2]
        dv, dvp = dU[0::2], dU[1::2]
        dforce = 4.0 * tvdag * dvp + (2.0 * tv * v - 1.0) * dv
        out[0] = dU[1]
        out[2 * n - 1] = dU[2 * n - 2]
        out[2:2 * n - 1:2] = dv[1:] - dv[:-1] - h2 * (dvp[:-1] + dvp[1:])
        out[1:2 * n - 2:2] = dvp[1:] - dvp[:-1] + h2 * (dforce[:-1] + dforce[1:])

    return F, jvp


def bvp_torch(n, h, tv, tvdag):
    import torch

    h2 = 0.5 * h
    tvt = torch.as_tensor(tv, device="cuda")
    tdt = torch.as_tensor(tvdag, device="cuda")

    def F_(res, U, p):
        v, vp = U[0::2], U[1::2]
        force = 4.0 * tdt * vp + (tvt * v - 1.0) * v
        res[0] = U[1]
        res[2 * n - 1] = U[2 * n - 2]
        res[2:2 * n - 1:2] = v[1:] - v[:-1] - h2 * (vp[:-1] + vp[1:])
        res[1:2 * n - 2:2] = vp[1:] - vp[:-1] + h2 * (force[:-1] + force[1:])

    def jvp_(out, U, dU, p):
        v = U[0::2]
        dv, dvp = dU[0::2], dU[1::2]
        dforce = 4.0 * tdt * dvp + (2.0 * tvt * v - 1.0) * dv
        out[0] = dU[1]
        out[2 * n - 1] = dU[2 * n - 2]
        out[2:2 * n - 1:2] = dv[1:] - dv[:-1] - h2 * (dvp[:-1] + dvp[1:])
        out[1:2 * n - 2:2] = dvp[1:] - dvp[:-1] + h2 * (dforce[:-1] + dforce[1:])

    return F_, jvp_


def test_bvp_kernels_match_oracle(nk, ctx, oracle):
    n, h, tv, tvdag, U0 = bvp_setup(201)
    Fn, Jn = bvp_np(n, h, tv, tvdag)
    Ft, Jt = bvp_torch(n, h, tv, tvdag)
    po = oracle.make_user_problem(2 * n, Fn, Jn)
    F_ = nk.UserResidual(Ft, Jt, name="Fbvp!")
    u = nk.DeviceVector.from_numpy(U0, ctx)
    res = u.zero()
    F_(res, u, None)
    rr, _ = oracle.residual(po, U0)
    assert rel(res.numpy(), rr) < 1e-14
    v0 = np.random.default_rng(0).standard_normal(2 * n)
    v = nk.DeviceVector.from_numpy(v0, ctx)
    out = u.zero()
    nk.mul_(out, nk.JacobianOperator(F_, res, u, None), v)
    jr, _ = oracle.jvp(po, U0, v0)
    assert rel(out.numpy(), jr) < 1e-14


def test_bvp_first_newton_step(nk, ctx, oracle):
    """BVP_solve of examples/bvp.jl:40-60: algo = :fgmres, N = J -> GmresPreconditioner(J, 30), through the host-driven
    Newton loop and through ak_newton_solve, against the oracle with the same callbacks.

    Only the first Newton step is compared.  This operator is so far from normal that GMRES itself is not
    reproducible beyond it: two IEEE-correct implementations of Fbvp!'s tangent (NumPy / torch, equal to 1e-14)
    give GMRES histories that differ by 2.5e-4 after 60 iterations and inner-GMRES preconditioner outputs that
    differ by 8e-7 (measured, tests/debug/dbg_user.py); the full solve takes 34 Newton steps / 3095 FGMRES iterations at
    n = 101 in the oracle and 23 / 2012 after a 1-ulp change of U0."""
    n, h, tv, tvdag, U0 = bvp_setup(101)
    Fn, Jn = bvp_np(n, h, tv, tvdag)
    Ft, Jt = bvp_torch(n, h, tv, tvdag)
    po = oracle.make_user_problem(2 * n, Fn, Jn)
    o = A.default_newton_opts(algo=A.AK_ALGO_FGMRES, max_niter=0)
    o.krylov.precond_n, o.krylov.precond_itmax = A.AK_PRECOND_INNER_GMRES, 30
    ur, sr, hr = oracle.newton(po, U0, o)
    assert [x["inner"] for x in hr] == [0, 1]
    F_ = nk.UserResidual(Ft, Jt, name="Fbvp!")
    for drive in (nk.newton_krylov_, nk.newton_krylov_native_):
        hist = []
        u = nk.DeviceVector.from_numpy(U0, ctx)
        _, r = drive(F_, u, None, u.zero(), algo="fgmres", N=lambda J: nk.GmresPreconditioner(J, 30), history=hist,
                     max_niter=0)
        assert [x["inner"] for x in hist] == [0, 1]
        for a, b in zip(hist, hr):
            assert abs(a["n_res"] - b["n_res"]) <= 1e-5 * b["n_res"]
        assert rel(u.numpy(), ur) < 1e-5


def test_user_residual_without_tangent_uses_finite_differences(nk, ctx, oracle):
    """No tangent callback: J v = (F(u + eps v) - F(u)) / eps (BASELINE north_star wording), F(u) cached by the
    residual of the same Newton step.  O(sqrt(eps)) accurate; compared with the exact tangent at that level and
    with the oracle's restatement of the same finite difference on the first two Newton steps."""
    n, h, tv, tvdag, U0 = bvp_setup(101)
    Fn, _ = bvp_np(n, h, tv, tvdag)
    Ft, Jt = bvp_torch(n, h, tv, tvdag)
    F_fd = nk.UserResidual(Ft, None, name="Fbvp! (fd)")
    F_an = nk.UserResidual(Ft, Jt, name="Fbvp!")
    u = nk.DeviceVector.from_numpy(U0, ctx)
    res = u.zero()
    v0 = np.random.default_rng(2).standard_normal(2 * n)
    out_fd, out_an = u.zero(), u.zero()
    nk.mul_(out_fd, nk.JacobianOperator(F_fd, res, u, None), nk.DeviceVector.from_numpy(v0, ctx))
    nk.mul_(out_an, nk.JacobianOperator(F_an, res, u, None), nk.DeviceVector.from_numpy(v0, ctx))
    assert 1e-12 < rel(out_fd.numpy(), out_an.numpy()) < 1e-5
    po = oracle.make_user_problem(2 * n, Fn, None)
    o = A.default_newton_opts(algo=A.AK_ALGO_FGMRES, max_niter=1)
    o.krylov.precond_n, o.krylov.precond_itmax = A.AK_PRECOND_INNER_GMRES, 30
    ur, sr, hr = oracle.newton(po, U0, o)
    hf = []
    uf = nk.DeviceVector.from_numpy(U0, ctx)
    nk.newton_krylov_(F_fd, uf, None, algo="fgmres", N=lambda J: nk.GmresPreconditioner(J, 30), max_niter=1, history=hf)
    assert [x["inner"] for x in hf] == [x["inner"] for x in hr]
    for a, b in zip(hf, hr):  # finite-difference noise (1e-8 per product) through an ill-conditioned solve
        assert abs(a["n_res"] - b["n_res"]) <= 5e-3 * b["n_res"]


def bratu_torch(dx, lam):
    import torch

    def F(y, p):  # bratu!  examples/bratu.jl:14-24, out of place
        z = torch.zeros(1, dtype=y.dtype, device=y.device)
        yl, yr = torch.cat([z, y[:-1]]), torch.cat([y[1:], z])
        return ((yr - 2.0 * y) + yl) / dx**2 + lam * torch.exp(y)

    return F


@pytest.mark.parametrize("algo,N", [("fgmres", 5), ("cg", None), ("gmres", None)], ids=["fgmres+gmres5", "cg", "gmres"])
def test_user_coded_bratu_walks_the_native_newton_path(nk, ctx, oracle, algo, N):
    """bratu! written by the caller (torch ops; tangent by forward-mode AD of the same code) through
    `newton_krylov(F, u0, p; ...)` (src/Ariadne.jl:245-248) for the call sites of examples/bratu.jl:82-87,151-157:
    same Newton / Krylov counts and residual history as the oracle's Bratu, to the oracle's own reproducibility
    (non-symmetric initial guess, see problems.generic)."""
    from test_gpu_solvers import assert_newton_parity, oracle_sensitivity

    d = P.generic(P.bratu1d(200, lam=1.0 if algo == "cg" else 3.5))
    po = P.oracle_problem(oracle, d)
    o = A.default_newton_opts(algo={"fgmres": A.AK_ALGO_FGMRES, "cg": A.AK_ALGO_CG, "gmres": A.AK_ALGO_GMRES}[algo])
    if N:
        o.krylov.precond_n, o.krylov.precond_itmax = A.AK_PRECOND_INNER_GMRES, N
    sens = oracle_sensitivity(oracle, po, d["u0"], o)
    hist = []
    u = nk.DeviceVector.from_numpy(d["u0"], ctx)
    kw = dict(N=lambda J: nk.GmresPreconditioner(J, N)) if N else {}
    _, r = nk.newton_krylov(bratu_torch(d["dx"], d["lam"]), u, None, algo=algo, history=hist, **kw)
    assert r.solved
    assert_newton_parity(u.numpy(), r, hist, sens)


@pytest.mark.parametrize("make", [lambda: P.bratu1d(300), lambda: P.bratu2d(20), lambda: P.heat2d(16, dt_scale=8.0, ic="poly")],
                         ids=["bratu1d", "bratu2d", "heat2d"])
def test_generic_fd_jvp_on_native_residuals(nk, ctx, oracle, make):
    """AK_JVP_FD works for every problem kind: two evaluations of the native residual kernel."""
    d = make()
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    res = u.zero()
    v0 = np.random.default_rng(3).standard_normal(d["u0"].shape)
    out_fd, out_an = u.zero(), u.zero()
    nk.mul_(out_an, nk.JacobianOperator(F_, res, u, p), nk.DeviceVector.from_numpy(v0, ctx))
    nk.mul_(out_fd, nk.JacobianOperator(F_, res, u, p, jvp_mode="fd2"), nk.DeviceVector.from_numpy(v0, ctx))
    assert rel(out_fd.numpy(), out_an.numpy()) < 2e-6
    # the oracle's restatement of the same finite difference agrees to rounding amplified by 1/eps
    po = P.oracle_problem(oracle, d, un=d["u0"] if d.get("scheme") else None)
    po.jvp_mode = A.AK_JVP_FD
    jr, _ = oracle.jvp(po, d["u0"], v0)
    assert rel(out_fd.numpy(), jr) < 1e-6


def test_failing_user_callback_is_reported(nk, ctx):
    def bad(res, u, p):
        raise RuntimeError("boom")

    F_ = nk.UserResidual(bad, None)
    u = nk.DeviceVector.from_numpy(np.ones(8), ctx)
    with pytest.raises(nk.AriadneError) as e:
        F_(u.zero(), u, None)
    assert e.value.code == A.AK_ERR_USER


def test_blow_up_guard(nk, ctx, oracle, capsys):
    """`isinf(n_res) || isnan(n_res)` -> `@error "Inner solver blew up"; break` (src/Ariadne.jl:353-356): the full Newton
    step of F(u) = sqrt(u) - 1 from u0 = (9, 16, 25, 36) lands on negative values, the next residual is NaN, and the
    loop stops before the step is counted — in the host-driven loop, in ak_newton_solve and in the oracle alike."""
    import torch

    u0 = np.array([9.0, 16.0, 25.0, 36.0])

    def F_np(res, u):
        with np.errstate(all="ignore"):
            res[:] = np.sqrt(u) - 1.0

    def J_np(out, u, v):
        with np.errstate(all="ignore"):
            out[:] = v / (2.0 * np.sqrt(u))

    o = A.default_newton_opts(forcing=A.AK_FORCING_NONE)
    ur, sr, hr = oracle.newton(oracle.make_user_problem(4, F_np, J_np), u0, o)
    assert sr["flags"] & A.AK_FLAG_NAN and not sr["solved"] and sr["outer_iterations"] == 0

    F_ = nk.UserResidual(lambda res, u, p: res.copy_(torch.sqrt(u) - 1.0),
                         lambda out, u, v, p: out.copy_(v / (2.0 * torch.sqrt(u))))
    for drive in (nk.newton_krylov_, nk.newton_krylov_native_):
        u = nk.DeviceVector.from_numpy(u0, ctx)
        _, r = drive(F_, u, None, u.zero(), forcing=None)
        assert not r.solved and r.stats.outer_iterations == 0
        assert np.allclose(u.numpy(), ur, rtol=1e-9)      # the overshot iterate (-3, -8, -15, -24)
    assert "Inner solver blew up" in capsys.readouterr().out
