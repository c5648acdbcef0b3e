"""One-sweep GMRES iterations (fuse = sweep, csrc/sweep.cu) piece by piece: the stored basis vectors, the recurrence
residuals and the solution against the reference op list (fuse = none) and the CPU oracle, on grids that exercise one
strip, several strips with a ragged last one, segments that split rows between blocks, periodic wrap in both
directions, re-orthogonalisation (two sweeps per iteration) and passes that outgrow the 24 vectors a sweep handles
(the pass continues with the eight-step blocked kernels).  Replaces kaxpy!/kdot/knorm/mul! of Krylov.jl's gmres!
(src/Ariadne.jl:338) like every other fusion level; tolerances as in tests/test_gpu_solvers.py."""
import numpy as np
import pytest

from newtonkrylov_jl_b200 import _abi as A
import ledger
import problems as P

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(77)


def rel(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300))


def run(nk, ctx, d, b0, fuse, coef_cached=False, **kw):
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    res = u.zero()
    coef = None
    if coef_cached:  # lambda e^u cached by the residual kernel, like inside newton_krylov!
        coef = u.similar()
        prob = F_.problem(u, p, coef=coef)
        import ctypes as C
        nk._lib.check(ctx.lib.ak_residual(ctx.h, C.byref(prob), C.c_void_p(u.ptr), C.c_void_p(res.ptr), None))
    J = nk.JacobianOperator(F_, res, u, p, coef=coef) if coef_cached else nk.JacobianOperator(F_, res, u, p)
    b = nk.DeviceVector.from_numpy(b0, ctx)
    ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res), memory=kw.pop("memory", 20))
    ctx.profile(True)
    nk.krylov_solve_(ws, J, b, history=True, fuse=fuse, **kw)
    nsweep = ctx.profile_read(13)[0]
    ctx.profile(False)
    return ws, nsweep


GRIDS = [
    ("one_strip", lambda: P.generic(P.bratu2d(40, 36))),
    ("three_strips_ragged", lambda: P.generic(P.bratu2d(600, 50))),
    ("exact_strips", lambda: P.generic(P.bratu2d(504, 24))),
    ("tall", lambda: P.generic(P.bratu2d(16, 3000))),
    ("tiny", lambda: P.generic(P.bratu2d(6, 5))),
    ("heat", lambda: P.heat2d(48, dt_scale=64.0, ic="poly")),
    ("heat_periodic", lambda: P.heat2d(36, dt_scale=64.0, bc=A.AK_BC_PERIODIC, ic="poly")),
    ("heat_periodic_wide", lambda: P.heat2d(300, dt_scale=64.0, bc=A.AK_BC_PERIODIC, ic="poly")),
    # 1-D problems: chunks of 240 points (192 for DG) play the role of the rows
    ("bratu1d_one_chunk", lambda: P.generic(P.bratu1d(200))),
    ("bratu1d_ragged", lambda: P.generic(P.bratu1d(250))),
    ("bratu1d_many_blocks", lambda: P.generic(P.bratu1d(100000))),
    ("heat1d", lambda: P.heat1d(998)),
    ("heat1d_two_chunks", lambda: P.heat1d(300)),
    ("dg_small", lambda: P.heat1d_dg(16, dt=1e-4)),
    ("dg_ragged", lambda: P.heat1d_dg(1000, dt=4e-7)),
    ("dg_many_blocks", lambda: P.heat1d_dg(40000, dt=2e-10)),
]


@pytest.mark.parametrize("name,make", GRIDS, ids=[g[0] for g in GRIDS])
def test_sweep_basis_and_history_match_the_reference_op_list(nk, ctx, oracle, name, make):
    d = make()
    b0 = RNG.standard_normal(d["u0"].shape)
    if d["kind"] == A.AK_HEAT1D:
        b0[0] = b0[-1] = 0.0  # consistent with the zero boundary rows of J (heat_1D.jl:57-89)
    kw = dict(rtol=1e-30, atol=0.0, restart=True, itmax=12, memory=6)
    ws_s, nsweep = run(nk, ctx, d, b0, "sweep", **kw)
    assert nsweep > 0, "the sweep kernel did not run"
    ws_n, _ = run(nk, ctx, d, b0, "none", **kw)
    hs, hn = np.array(ws_s.stats.residuals), np.array(ws_n.stats.residuals)
    assert ws_s.stats.niter == ws_n.stats.niter == 12
    # basis of the last cycle: stored (un-normalised) vector / scale == the normalised vector of the reference op list
    worst = 0.0
    for i in range(6):
        vs, ss = ws_s.basis(i)
        vn, sn = ws_n.basis(i)
        a, b = vs.numpy().reshape(-1) / ss, vn.numpy().reshape(-1) / sn
        worst = max(worst, float(np.max(np.abs(a - b))))
    ledger.record("sweep_vs_reference_op_list", name, max_basis_entry_dev=worst,
                  max_hist_dev_rel_beta=float(np.max(np.abs(hs - hn)) / hn[0]), x_rel_dev=rel(ws_s.x.numpy(), ws_n.x.numpy()))
    assert worst < 1e-11, worst
    assert np.max(np.abs(hs - hn)) <= 1e-12 * hn[0]
    assert rel(ws_s.x.numpy(), ws_n.x.numpy()) < 1e-11
    # and against the oracle
    po = P.oracle_problem(oracle, d, un=d["u0"] if d.get("scheme") else None)
    xr, sr, hr = oracle.krylov_solve(po, d["u0"], b0, memory=6, hist_cap=64, rtol=1e-30, atol=0.0, restart=True, itmax=12)
    assert sr["niter"] == 12
    assert np.max(np.abs(hs - hr)) <= 1e-10 * hr[0]
    assert rel(ws_s.x.numpy(), xr) < 1e-9


@pytest.mark.parametrize("opts", [dict(restart=True, itmax=30, memory=24), dict(itmax=40, memory=20),
                                  dict(restart=True, reorthogonalization=True, itmax=17, memory=7),
                                  dict(reorthogonalization=True, itmax=31, memory=20), dict(restart=True, itmax=50, memory=30)],
                         ids=["memory24", "outgrows_sweeps", "reorth_restart", "reorth_outgrows", "memory30"])
def test_sweep_options_against_the_oracle(nk, ctx, oracle, opts):
    d = P.generic(P.bratu2d(260, 70))
    b0 = RNG.standard_normal(d["u0"].shape)
    opts = dict(opts)
    mem = opts["memory"]
    ws, nsweep = run(nk, ctx, d, b0, "sweep", coef_cached=True, rtol=1e-30, atol=0.0, **opts)
    assert nsweep > 0
    po = P.oracle_problem(oracle, d)
    okw = {k: v for k, v in opts.items() if k != "memory"}
    xr, sr, hr = oracle.krylov_solve(po, d["u0"], b0, memory=mem, hist_cap=128, rtol=1e-30, atol=0.0, **okw)
    h = np.array(ws.stats.residuals)
    ledger.record("sweep_options_vs_oracle", "_".join(f"{k}={v}" for k, v in sorted(opts.items())), niter_gpu=ws.stats.niter,
                  niter_oracle=sr["niter"], max_hist_dev_rel_beta=float(np.max(np.abs(h - hr)) / hr[0]),
                  x_rel_dev=rel(ws.x.numpy(), xr), sweep_launches=nsweep)
    assert (ws.stats.niter, ws.stats.npass) == (sr["niter"], sr["npass"])
    assert np.max(np.abs(h - hr)) <= 1e-10 * hr[0]
    assert rel(ws.x.numpy(), xr) < 1e-9


def test_sweep_stops_on_convergence_like_the_oracle(nk, ctx, oracle):
    """A solve that converges in the middle of a cycle: the sweep queued behind the verdict is a no-op."""
    d = P.heat2d(64, dt_scale=4.0, ic="poly")
    b0 = RNG.standard_normal(d["u0"].shape)
    ws, nsweep = run(nk, ctx, d, b0, "sweep", rtol=1e-9, restart=True)
    po = P.oracle_problem(oracle, d, un=d["u0"])
    xr, sr, hr = oracle.krylov_solve(po, d["u0"], b0, hist_cap=256, rtol=1e-9, restart=True)
    assert nsweep > 0 and ws.stats.solved and sr["solved"]
    assert (ws.stats.niter, ws.stats.npass) == (sr["niter"], sr["npass"])
    assert np.max(np.abs(np.array(ws.stats.residuals) - hr)) <= 1e-10 * hr[0]
    assert rel(ws.x.numpy(), xr) < 1e-9


def test_newton_with_sweeps_matches_oracle(nk, ctx, oracle):
    """newton_krylov! with defaults (non-restarted GMRES, Eisenstat-Walker) on 2-D Bratu through the C++ loop."""
    d = P.generic(P.bratu2d(64, 48))
    po = P.oracle_problem(oracle, d)
    ur, sr, hr = oracle.newton(po, d["u0"])
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    hist = []
    _, r = nk.newton_krylov_native_(F_, u, p, None, history=hist, krylov_kwargs=dict(fuse="sweep"))
    assert r.solved and sr["solved"] and r.stats.outer_iterations == sr["outer_iterations"]
    assert [h["inner"] for h in hist] == [h["inner"] for h in hr]
    for a, b in zip(hist, hr):
        assert abs(a["n_res"] - b["n_res"]) <= 1e-8 * b["n_res"] + 1e-13 * hr[0]["n_res"]
    assert rel(u.numpy(), ur) < 1e-8
