"""GPU parity tests, kernel level: every call goes through the C ABI (ctypes) and is compared
with the CPU oracle on the same seeded inputs.

Tolerances: the heat / DG stencils use only +,-,*,/ in the reference's order -> bit-exact.
Bratu involves exp(): CUDA libdevice and glibc each stay within 1 ulp of the true value, so the
coefficient lambda*exp(u) may differ by a few ulp -> compared relative to the row scale.
Reductions differ in summation order -> 1e-13 relative.
"""
import ctypes as C

import numpy as np
import pytest

from newtonkrylov_jl_b200 import _abi as A
import problems as P

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(0)

CASES = [
    ("bratu1d", lambda: P.bratu1d(1000)),
    ("bratu1d_odd", lambda: P.bratu1d(1001)),
    ("bratu1d_tiny", lambda: P.bratu1d(3)),
    ("bratu1d_single_point", lambda: P.bratu1d(1)),
    ("bratu1d_two_points", lambda: P.bratu1d(2)),
    ("bratu2d_1x1", lambda: P.bratu2d(1, 1)),
    ("bratu2d_single_row", lambda: P.bratu2d(257, 1)),
    ("bratu2d_single_column", lambda: P.bratu2d(1, 300)),
    ("bratu2d_two_columns", lambda: P.bratu2d(2, 65)),
    ("bratu2d_warp_edge", lambda: P.bratu2d(4 * 32 + 4, 19)),
    ("bratu2d_block_edge", lambda: P.bratu2d(4 * 128 + 8, 11)),
    ("bratu2d", lambda: P.bratu2d(64)),
    ("bratu2d_rect", lambda: P.bratu2d(130, 37)),
    ("bratu2d_odd", lambda: P.bratu2d(33, 31)),
    ("bratu2d_wide", lambda: P.bratu2d(1024, 8)),
    ("heat1d", lambda: P.heat1d(100)),
    ("heat1d_periodic", lambda: P.heat1d(101, bc=A.AK_BC_PERIODIC)),
    ("heat1d_min", lambda: P.heat1d(1)),
    ("heat2d", lambda: P.heat2d(40)),
    ("heat2d_periodic", lambda: P.heat2d(40, bc=A.AK_BC_PERIODIC, ic="poly")),
    ("heat2d_odd", lambda: P.heat2d(37, ic="poly")),
    ("heat2d_1x1", lambda: P.heat2d(1, ic="poly")),
    ("heat2d_2x2_periodic", lambda: P.heat2d(2, bc=A.AK_BC_PERIODIC, ic="poly")),
    ("heat2d_block_edge_periodic", lambda: P.heat2d(516, bc=A.AK_BC_PERIODIC, ic="poly")),
    ("dg", lambda: P.heat1d_dg(40)),
    ("dg_unaligned_warp", lambda: P.heat1d_dg(77)),
    ("dg_min", lambda: P.heat1d_dg(2)),
]
EXACT = {A.AK_HEAT1D, A.AK_HEAT2D, A.AK_HEAT1D_DG}


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_residual_matches_oracle(nk, ctx, oracle, name, make):
    d = make()
    u0 = d["u0"] + 0.01 * RNG.standard_normal(d["u0"].shape)
    d = dict(d, u0=u0)
    un_host = d["u0"] * 0.9 if d["kind"] in EXACT else None
    F_, u, p, un = P.device_setup(nk, ctx, d)
    if un is not None:
        un.set(un_host)
    res = u.zero()
    F_(res, u, p)
    po = P.oracle_problem(oracle, d, un=un_host)
    ref, u_after = oracle.residual(po, u0)
    got = res.numpy()
    if d["kind"] in EXACT:
        assert np.array_equal(got, ref), f"max ulp {P.ulp_diff(got, ref)}"
    else:
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(got - ref)) <= 4e-16 * scale * 8
    # boundary side effect of bc!(u) (heat_1D.jl:16)
    assert np.array_equal(u.numpy(), u_after)


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_jvp_matches_oracle(nk, ctx, oracle, name, make):
    d = make()
    v0 = RNG.standard_normal(d["u0"].shape)
    F_, u, p, un = P.device_setup(nk, ctx, d)
    res = u.zero()
    v = nk.DeviceVector.from_numpy(v0, ctx)
    out = u.zero()
    J = nk.JacobianOperator(F_, res, u, p)
    assert J.size() == (u.n, u.n) and len(J) == u.n * u.n and J.eltype() == np.float64
    nk.mul_(out, J, v)
    po = P.oracle_problem(oracle, d, un=d["u0"] if d["kind"] in EXACT else None)
    ref, v_after = oracle.jvp(po, d["u0"], v0)
    got = out.numpy()
    if d["kind"] in EXACT:
        assert np.array_equal(got, ref), f"max ulp {P.ulp_diff(got, ref)}"
    else:
        assert np.max(np.abs(got - ref)) <= 1e-15 * np.max(np.abs(ref)) * 8
    # tangent BC: v's boundary entries are overwritten like forward mode through bc!(u) does
    assert np.array_equal(v.numpy(), v_after)


SCHEME_CASES = [(nm, mk, sc) for nm, mk in [("heat1d", lambda: P.heat1d(50)), ("heat1d_periodic", lambda: P.heat1d(33, bc=A.AK_BC_PERIODIC)),
                                              ("heat2d", lambda: P.heat2d(24, ic="poly")),
                                              ("heat2d_periodic", lambda: P.heat2d(20, bc=A.AK_BC_PERIODIC, ic="poly")),
                                              ("dg", lambda: P.heat1d_dg(19))]
                for sc in (A.AK_MIDPOINT, A.AK_TRAPEZOID)]


@pytest.mark.parametrize("name,make,scheme", SCHEME_CASES, ids=[f"{c[0]}-{'midpoint' if c[2] == A.AK_MIDPOINT else 'trapezoid'}" for c in SCHEME_CASES])
def test_midpoint_trapezoid_match_oracle(nk, ctx, oracle, name, make, scheme):
    """G_Midpoint! / G_Trapezoid! (examples/implicit.jl:17-37): residual, tangent and their BC side effects,
    bit for bit (no transcendental functions involved)."""
    d = dict(make(), scheme=scheme)
    G_ = nk.G_Midpoint_ if scheme == A.AK_MIDPOINT else nk.G_Trapezoid_
    u0 = d["u0"] + 0.01 * RNG.standard_normal(d["u0"].shape)
    un0 = d["u0"] * 0.9 + 0.05
    F_e, u, p, un = P.device_setup(nk, ctx, dict(d, u0=u0))
    F_ = nk.ImplicitResidual(G_, F_e.f_)
    un.set(un0)
    res = u.zero()
    F_(res, u, p)
    po = P.oracle_problem(oracle, d, un=un0)
    ref, u_after = oracle.residual(po, u0)
    assert np.array_equal(res.numpy(), ref), P.ulp_diff(res.numpy(), ref)
    assert np.array_equal(u.numpy(), u_after)
    assert np.array_equal(un.numpy(), po._keep)  # Trapezoid: f!(du_n, u_n) runs its BC code on u_n in place
    v0 = RNG.standard_normal(d["u0"].shape)
    v, out = nk.DeviceVector.from_numpy(v0, ctx), u.zero()
    nk.mul_(out, nk.JacobianOperator(F_, res, u, p), v)
    refj, v_after = oracle.jvp(po, u0, v0)
    assert np.array_equal(out.numpy(), refj) and np.array_equal(v.numpy(), v_after)


def test_jvp_with_cached_coefficient_equals_recomputed(nk, ctx):
    """Bratu: JVP reading lambda*exp(u) cached by ak_residual == JVP recomputing exp(u)."""
    for d in (P.bratu1d(513), P.bratu2d(96, 40)):
        F_, u, p, _ = P.device_setup(nk, ctx, d)
        res, coef = u.zero(), u.similar()
        prob = F_.problem(u, p, coef=coef)
        import ctypes as C
        nk._lib.check(ctx.lib.ak_residual(ctx.h, C.byref(prob), C.c_void_p(u.ptr), C.c_void_p(res.ptr), None))
        v = nk.DeviceVector.from_numpy(RNG.standard_normal(d["u0"].shape), ctx)
        o1, o2 = u.zero(), u.zero()
        nk.mul_(o1, nk.JacobianOperator(F_, res, u, p, coef=coef), v)
        nk.mul_(o2, nk.JacobianOperator(F_, res, u, p), v)
        assert np.array_equal(o1.numpy(), o2.numpy())


def test_known_answer_jacobian_2x2(nk, ctx):
    """test/runtests.jl:28-54 through the GPU path."""
    u = nk.DeviceVector.from_numpy(np.array([3.0, 5.0]), ctx)
    res = u.zero()
    J = nk.JacobianOperator(nk.simple_F_, res, u, None)
    assert J.size() == (2, 2) and len(J) == 4 and J.eltype() == np.float64
    out = nk.DeviceVector.from_numpy(np.array([np.nan, np.nan]), ctx)
    nk.mul_(out, J, nk.DeviceVector.from_numpy(np.array([1.0, 0.0]), ctx))
    assert np.array_equal(out.numpy(), np.array([6.0, 7.38905609893065]))
    nk.mul_(out, nk.transpose(J), nk.DeviceVector.from_numpy(np.array([1.0, 0.0]), ctx))
    assert np.array_equal(out.numpy(), np.array([6.0, 10.0]))
    Jd = nk.collect(J)
    assert np.array_equal(Jd, np.array([[6.0, 10.0], [np.exp(2.0), 10.0]]))
    assert np.array_equal(nk.collect(nk.transpose(J)), Jd.T)
    v = RNG.random(2)
    nk.mul_(out, J, nk.DeviceVector.from_numpy(v, ctx))
    assert np.allclose(out.numpy(), Jd @ v, rtol=1e-15)


def test_empty_vectors(nk, ctx):
    """n = 0: reductions return 0, element-wise hooks are no-ops (Krylov.jl calls them with length(x))."""
    x = nk.DeviceVector.from_numpy(np.array([1.0, 2.0]), ctx)
    assert nk.kdot(0, x, x) == 0.0 and nk.knorm(0, x) == 0.0
    nk.kaxpy_(0, 2.0, x, x)
    nk.kscal_(0, 2.0, x)
    assert np.array_equal(x.numpy(), np.array([1.0, 2.0]))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 1000, 4099, 1 << 20])
def test_vector_hooks(nk, ctx, oracle, n):
    """Krylov.k* hooks (examples/halovector.jl:51-147) vs numpy on ragged sizes."""
    x0, y0 = RNG.standard_normal(n), RNG.standard_normal(n)
    x, y = nk.DeviceVector.from_numpy(x0, ctx), nk.DeviceVector.from_numpy(y0, ctx)
    assert nk.kdot(n, x, y) == pytest.approx(float(np.dot(x0, y0)), rel=1e-12, abs=1e-12 * np.sqrt(n))
    assert nk.knorm(n, x) == pytest.approx(float(np.linalg.norm(x0)), rel=1e-13)
    nk.kaxpy_(n, 0.3, x, y)
    y1 = np.array([np.float64(0.3) * a for a in x0]) + y0 if n < 10 else 0.3 * x0 + y0
    assert np.allclose(y.numpy(), y1, rtol=1e-15, atol=1e-16)
    y1 = y.numpy()
    nk.kaxpby_(n, 2.0, x, -0.5, y)
    assert np.allclose(y.numpy(), 2.0 * x0 - 0.5 * y1, rtol=1e-15, atol=1e-16)
    nk.kscal_(n, 3.0, x)
    assert np.array_equal(x.numpy(), 3.0 * x0)
    nk.kcopy_(n, y, x)
    assert np.array_equal(y.numpy(), x.numpy())
    nk.kfill_(y, 1.25)
    assert np.array_equal(y.numpy(), np.full(n, 1.25))
    nk.kdivcopy_(n, y, x, 7.0)
    assert np.array_equal(y.numpy(), (3.0 * x0) / 7.0)
    xa, ya = x.numpy(), y.numpy()
    c, s = 0.6, 0.8
    nk.kref_(n, x, y, c, s)
    assert np.allclose(x.numpy(), c * xa + s * ya, rtol=1e-15, atol=1e-16)
    assert np.allclose(y.numpy(), s * xa - c * ya, rtol=1e-15, atol=1e-16)


def test_vector_hooks_unaligned_views(nk, ctx):
    """Operands that are not 32-byte aligned take the scalar path and give the same numbers."""
    n = 1003
    base = nk.DeviceVector.from_numpy(RNG.standard_normal(2 * n + 8), ctx)
    h = base.numpy()
    x = nk.DeviceVector(ctx, (n,), ptr=base.ptr + 8, owner=base)
    y = nk.DeviceVector(ctx, (n,), ptr=base.ptr + 8 * (n + 3), owner=base)
    x0, y0 = h[1:1 + n], h[n + 3:2 * n + 3]
    assert nk.kdot(n, x, y) == pytest.approx(float(np.dot(x0, y0)), rel=1e-12)
    nk.kaxpy_(n, -1.5, x, y)
    assert np.allclose(y.numpy(), y0 - 1.5 * x0, rtol=1e-15, atol=1e-16)


def test_halovector_layout_bridge(nk, ctx, oracle):
    """examples/halovector.jl:3-45 / heat_2D.jl:76-91: padded OffsetArray <-> compact slab."""
    nx, ny = 12, 9
    padded = RNG.standard_normal((nx + 2, ny + 2))  # Julia index order [i, j]
    hv = nk.HaloVector.from_padded(padded, ctx)
    assert len(hv) == nx * ny  # logical length = interior only (halovector.jl:17-21)
    assert np.array_equal(hv.numpy(), padded[1:-1, 1:-1].T)
    z = hv.padded(A.AK_BC_ZERO)
    assert np.array_equal(z[1:-1, 1:-1], padded[1:-1, 1:-1])
    assert np.all(z[0, :] == 0) and np.all(z[-1, :] == 0) and np.all(z[:, 0] == 0) and np.all(z[:, -1] == 0)
    per = hv.padded(A.AK_BC_PERIODIC)
    ref = np.zeros((ny + 2, nx + 2))
    import ctypes as C
    comp = np.ascontiguousarray(padded[1:-1, 1:-1].T)
    oracle.load().ok_halo_unpack(ref.ctypes.data_as(A.c_double_p), comp.ctypes.data_as(A.c_double_p), nx, ny, A.AK_BC_PERIODIC)
    assert np.array_equal(per, ref.T)


def test_linearity_and_symmetry_large(nk, ctx):
    """Size-independent properties at a size the oracle is not asked to match (2048^2):
    J(u)(a v + b w) == a J v + b J w to rounding, and <w, J v> == <v, J w> (symmetric J)."""
    d = P.bratu2d(2048)
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    res = u.zero()
    J = nk.JacobianOperator(F_, res, u, p)
    n = u.n
    v = nk.DeviceVector.from_numpy(RNG.standard_normal(d["u0"].shape), ctx)
    w = nk.DeviceVector.from_numpy(RNG.standard_normal(d["u0"].shape), ctx)
    Jv, Jw, comb, Jc = u.zero(), u.zero(), u.zero(), u.zero()
    nk.mul_(Jv, J, v)
    nk.mul_(Jw, J, w)
    nk.kcopy_(n, comb, w)
    nk.kaxpby_(n, 2.0, v, -3.0, comb)  # comb = 2v - 3w
    nk.mul_(Jc, J, comb)
    nk.kaxpby_(n, 2.0, Jv, -3.0, Jw)   # Jw <- 2 Jv - 3 Jw
    nk.kaxpy_(n, -1.0, Jw, Jc)
    assert nk.knorm(n, Jc) <= 1e-12 * nk.knorm(n, Jw)
    nk.mul_(Jv, J, v)
    nk.mul_(Jw, J, w)
    a, b = nk.kdot(n, w, Jv), nk.kdot(n, v, Jw)
    assert a == pytest.approx(b, rel=1e-11)


def test_fused_finite_difference_jvp(nk, ctx, oracle):
    """AK_JVP_FD_FUSED (north_star item 2): (F(u + eps v) - F(u)) / eps in one pass that reads u and v and never
    materialises u + eps v.  It approximates the exact tangent to O(eps |F''| + eps_mach |F| / eps) ~ 1e-7 relative
    — which is why the analytic tangent, not this mode, is the parity path."""
    for d in (P.bratu2d(64), P.bratu2d(130, 37), P.bratu2d(33, 31), P.bratu1d(1000), P.bratu1d(1001), P.bratu1d(5)):
        F_, u, p, _ = P.device_setup(nk, ctx, d)
        res = u.zero()
        v0 = RNG.standard_normal(d["u0"].shape)
        v0 /= np.linalg.norm(v0)
        v = nk.DeviceVector.from_numpy(v0, ctx)
        exact, fd = u.zero(), u.zero()
        nk.mul_(exact, nk.JacobianOperator(F_, res, u, p), v)
        nk.mul_(fd, nk.JacobianOperator(F_, res, u, p, jvp_mode="fd"), v)
        e, f = exact.numpy(), fd.numpy()
        assert 1e-12 < np.linalg.norm(f - e) / np.linalg.norm(e) < 1e-5
        # the same finite difference formed from two oracle residuals agrees to rounding of the difference
        po = P.oracle_problem(oracle, d)
        eps = 1.4901161193847656e-08
        f1, _ = oracle.residual(po, d["u0"] + eps * v0)
        f0, _ = oracle.residual(po, d["u0"])
        ref = (f1 - f0) / eps
        assert np.linalg.norm(f - ref) / np.linalg.norm(ref) < 1e-6
    # Newton with the FD operator converges in the same number of steps (histories agree to ~1e-6 only)
    d = P.generic(P.bratu2d(32))
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    h_fd, h_ex = [], []
    _, r_fd = nk.newton_krylov_(F_, u, p, None, history=h_fd, jvp_mode="fd")
    F_, u2, p, _ = P.device_setup(nk, ctx, d)
    _, r_ex = nk.newton_krylov_(F_, u2, p, None, history=h_ex)
    assert r_fd.solved and r_fd.stats.outer_iterations == r_ex.stats.outer_iterations
    assert np.linalg.norm(u.numpy() - u2.numpy()) / np.linalg.norm(u2.numpy()) < 1e-6


# ---- collect(J) by probing, J^T for the DG operator, batched mul! ----------------------------------
COLLECT_CASES = [
    ("bratu1d", lambda: P.bratu1d(50)), ("bratu1d_tiny", lambda: P.bratu1d(3)),
    ("heat1d_zero", lambda: P.heat1d(40)), ("heat1d_periodic", lambda: P.heat1d(41, bc=A.AK_BC_PERIODIC)),
    ("dg", lambda: P.heat1d_dg(13)), ("dg_min", lambda: P.heat1d_dg(3)),
    ("bratu2d", lambda: P.bratu2d(7, 5)), ("bratu2d_thin", lambda: P.bratu2d(9, 2)),
    ("heat2d_zero", lambda: P.heat2d(8, dt_scale=16.0)), ("heat2d_periodic", lambda: P.heat2d(7, dt_scale=16.0, bc=A.AK_BC_PERIODIC)),
]


@pytest.mark.parametrize("name,make", COLLECT_CASES, ids=[c[0] for c in COLLECT_CASES])
def test_sparse_collect_equals_column_by_column_collect(nk, ctx, name, make):
    """`collect(J)` (src/Ariadne.jl:140-162) assembled from a handful of colour-probe JVPs equals the literal
    one-JVP-per-column algorithm entry for entry."""
    d = make()
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    J = nk.JacobianOperator(F_, u.zero(), u, p)
    dense = nk.collect(J)
    sparse = nk.collect(J, sparse=True)
    assert sparse.shape == dense.shape
    assert np.array_equal(sparse.toarray(), dense)
    assert sparse.nnz <= 5 * dense.shape[0] if d["kind"] != A.AK_HEAT1D_DG else sparse.nnz <= 12 * dense.shape[0]


@pytest.mark.parametrize("scheme", ["euler", "midpoint", "trapezoid"])
def test_dg_transpose_product(nk, ctx, oracle, scheme):
    """J^T v for the DG operator (src/Ariadne.jl:93-107): J^T = W J W^-1 by upwind-SBP duality, against the
    transpose of the collected Jacobian and the oracle's dense probe."""
    d = P.heat1d_dg(16, dt=1e-3)
    d["scheme"] = {"euler": A.AK_EULER, "midpoint": A.AK_MIDPOINT, "trapezoid": A.AK_TRAPEZOID}[scheme]
    G_ = {"euler": nk.G_Euler_, "midpoint": nk.G_Midpoint_, "trapezoid": nk.G_Trapezoid_}[scheme]
    u = nk.DeviceVector.from_numpy(d["u0"], ctx)
    un = nk.DeviceVector.from_numpy(d["u0"], ctx)
    F_ = nk.ImplicitResidual(G_, nk.heat_1D_DG_)
    J = nk.JacobianOperator(F_, u.zero(), u, (un, d["dt"], un.zero(), (d["dx"],), 0.0))
    v0 = RNG.standard_normal(d["nx"])
    out = u.zero()
    nk.mul_(out, nk.transpose(J), nk.DeviceVector.from_numpy(v0, ctx))
    Jd = nk.collect(J)
    ref = Jd.T @ v0
    assert np.linalg.norm(out.numpy() - ref) <= 1e-12 * np.linalg.norm(ref)
    po = P.oracle_problem(oracle, d, un=d["u0"])
    ro = oracle.jvp_transpose_dense(po, d["u0"], v0)
    assert np.linalg.norm(out.numpy() - ro) <= 1e-12 * np.linalg.norm(ro)


def test_batched_jvp(nk, ctx):
    """mul!(Out, J, V) for a matrix V (src/Ariadne.jl:69-83, test/runtests.jl:57-66) == column-wise mul!."""
    d = P.bratu2d(20, 12)
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    J = nk.JacobianOperator(F_, u.zero(), u, p)
    V0 = RNG.standard_normal((5, 12 * 20))
    V = nk.DeviceVector.from_numpy(V0, ctx)
    Out = nk.DeviceVector(ctx, (5, 12 * 20))
    nk.mul_batched_(Out, J, V)
    for c in range(5):
        o = u.zero()
        nk.mul_(o, J, nk.DeviceVector.from_numpy(V0[c].reshape(12, 20), ctx))
        assert np.array_equal(Out.numpy()[c], o.numpy().reshape(-1))


BATCH_CASES = [("bratu2d", lambda: P.bratu2d(20, 12)), ("bratu2d_odd", lambda: P.bratu2d(33, 7)),
               ("bratu2d_tall", lambda: P.bratu2d(8, 37)), ("bratu2d_wide", lambda: P.bratu2d(1030, 9)),
               ("bratu1d", lambda: P.bratu1d(1000)), ("bratu1d_odd", lambda: P.bratu1d(37)),
               ("heat2d", lambda: P.heat2d(12, dt_scale=8.0)), ("dg", lambda: P.heat1d_dg(9))]


@pytest.mark.parametrize("cached", [False, True], ids=["exp_from_u", "cached_coef"])
@pytest.mark.parametrize("name,make", BATCH_CASES, ids=[c[0] for c in BATCH_CASES])
def test_multi_rhs_jvp_equals_columnwise_bit_for_bit(nk, ctx, name, make, cached):
    """The multi-RHS tangent kernels (lambda e^u read once for all columns of `mul!(Out, J, V)`, src/Ariadne.jl:69-83)
    give every column exactly what the single-column kernel gives; guard bands around V and Out stay intact."""
    d = make()
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    n, ncols = u.n, 6
    coef = None
    if cached and d["kind"] in (A.AK_BRATU1D, A.AK_BRATU2D):
        coef = u.similar()
        prob = F_.problem(u, p, coef=coef)
        nk._lib.check(ctx.lib.ak_residual(ctx.h, C.byref(prob), C.c_void_p(u.ptr), C.c_void_p(u.zero().ptr), None))
    J = nk.JacobianOperator(F_, u.zero(), u, p, coef=coef)
    V0 = RNG.standard_normal((ncols, n))
    base, (V, Out), mask = _guarded(nk, ctx, [V0, np.zeros((ncols, n))])
    guards_before = base.numpy()[mask]
    nk.mul_batched_(Out, J, V)
    got = Out.numpy()
    assert np.array_equal(np.isnan(base.numpy()[mask]), np.isnan(guards_before)) and not np.isnan(got).any()
    for c in range(ncols):
        o = u.zero()
        nk.mul_(o, J, nk.DeviceVector.from_numpy(V0[c].reshape(u.shape), ctx))
        assert np.array_equal(got[c], o.numpy().reshape(-1)), (name, c)


@pytest.mark.parametrize("scheme", ["euler", "trapezoid"])
def test_periodic_heat1d_transpose_product(nk, ctx, scheme):
    """J^T v (src/Ariadne.jl:93-107) for heat_1D! with periodic_bc! (heat_1D.jl:39-42), whose Jacobian is not
    symmetric: `collect(transpose(J)) == transpose(collect(J))` exactly (the reference's own test, runtests.jl:54)
    and a random product against the dense transpose."""
    d = P.heat1d(41, bc=A.AK_BC_PERIODIC)
    d["scheme"] = {"euler": A.AK_EULER, "trapezoid": A.AK_TRAPEZOID}[scheme]
    G_ = {"euler": nk.G_Euler_, "trapezoid": nk.G_Trapezoid_}[scheme]
    u = nk.DeviceVector.from_numpy(d["u0"], ctx)
    un = nk.DeviceVector.from_numpy(d["u0"], ctx)
    F_ = nk.ImplicitResidual(G_, nk.heat_1D_)
    J = nk.JacobianOperator(F_, u.zero(), u, (un, d["dt"], un.zero(), (d["a"], d["dx"], nk.bc_periodic_), 0.0))
    Jd = nk.collect(J)
    assert not np.array_equal(Jd, Jd.T)
    assert np.array_equal(nk.collect(nk.transpose(J)), Jd.T)
    v0 = RNG.standard_normal(d["nx"])
    out = u.zero()
    nk.mul_(out, nk.transpose(J), nk.DeviceVector.from_numpy(v0, ctx))
    ref = Jd.T @ v0
    assert np.linalg.norm(out.numpy() - ref) <= 1e-13 * np.linalg.norm(ref)


# ---- guard bands: compute-sanitizer is not available on the GPU pool, so out-of-bounds accesses are caught here -----
GUARD = 64


def _guarded(nk, ctx, arrays):
    """Place the arrays back to back in ONE device buffer, separated and surrounded by NaN guard bands; vectors start
    32-byte aligned.  An out-of-bounds read poisons the result with NaN, an out-of-bounds write clears a guard."""
    sizes = [int(np.asarray(a).size) for a in arrays]
    padded = [(s + 3) // 4 * 4 for s in sizes]
    total = GUARD + sum(p + GUARD for p in padded)
    host = np.full(total, np.nan)
    offs, o = [], GUARD
    for a, s, p in zip(arrays, sizes, padded):
        host[o:o + s] = np.ravel(a)
        offs.append(o)
        o += p + GUARD
    base = nk.DeviceVector.from_numpy(host, ctx)
    vecs = [nk.DeviceVector(ctx, np.asarray(a).shape, ptr=base.ptr + 8 * off, owner=base) for a, off in zip(arrays, offs)]
    mask = np.ones(total, dtype=bool)
    for off, s in zip(offs, sizes):
        mask[off:off + s] = False
    return base, vecs, mask


GUARD_CASES = [("bratu1d", lambda: P.bratu1d(1001)), ("bratu2d", lambda: P.bratu2d(36, 27)), ("bratu2d_odd", lambda: P.bratu2d(33, 5)),
               ("heat1d", lambda: P.heat1d(99)), ("heat1d_periodic", lambda: P.heat1d(100, bc=A.AK_BC_PERIODIC)),
               ("heat2d", lambda: P.heat2d(20, dt_scale=8.0)), ("heat2d_periodic", lambda: P.heat2d(12, dt_scale=8.0, bc=A.AK_BC_PERIODIC)),
               ("dg", lambda: P.heat1d_dg(9))]


@pytest.mark.parametrize("name,make", GUARD_CASES, ids=[c[0] for c in GUARD_CASES])
def test_stencil_kernels_stay_inside_their_vectors(nk, ctx, oracle, name, make):
    """Residual, JVP (plain and with the fused first dot of the Arnoldi step) and the vector hooks on vectors
    surrounded by NaN guard bands: results equal the oracle's (no NaN leaked in) and every guard is still NaN."""
    d = make()
    v0 = RNG.standard_normal(d["u0"].shape)
    zeros = np.zeros_like(d["u0"])
    base, (u, un, v, res, out), mask = _guarded(nk, ctx, [d["u0"], d["u0"], v0, zeros, zeros])
    k = d["kind"]
    bc = nk.bc_zero_ if d.get("bc", A.AK_BC_ZERO) == A.AK_BC_ZERO else nk.bc_periodic_
    if k == A.AK_BRATU1D:
        F_, p = nk.bratu_, (d["dx"], d["lam"])
    elif k == A.AK_BRATU2D:
        F_, p = nk.bratu2d_, (d["dx"], d["dy"], d["lam"])
    elif k == A.AK_HEAT1D:
        F_, p = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_), (un, d["dt"], None, (d["a"], d["dx"], bc), 0.0)
    elif k == A.AK_HEAT2D:
        F_, p = nk.ImplicitResidual(nk.G_Euler_, nk.diffusion_), (un, d["dt"], None, (d["a"], d["dx"], d["dy"], bc), 0.0)
    else:
        F_, p = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_DG_), (un, d["dt"], None, (d["dx"],), 0.0)
    po = P.oracle_problem(oracle, d, un=d["u0"] if d.get("scheme") else None)
    F_(res, u, p)
    rr, _ = oracle.residual(po, d["u0"])
    assert np.allclose(res.numpy(), rr, rtol=1e-12, atol=1e-12 * np.max(np.abs(rr)))
    nk.mul_(out, nk.JacobianOperator(F_, res, u, p), v)
    jr, _ = oracle.jvp(po, d["u0"], v0)
    assert np.allclose(out.numpy(), jr, rtol=1e-12, atol=1e-12 * np.max(np.abs(jr)))
    n = u.n
    assert np.isfinite(nk.kdot(n, out, res)) and np.isfinite(nk.knorm(n, out))
    nk.kaxpy_(n, 0.5, out, res)
    nk.kaxpby_(n, 2.0, out, -1.0, res)
    nk.kscal_(n, 0.25, res)
    nk.kcopy_(n, out, res)
    nk.kref_(n, out, res, 0.6, 0.8)
    assert np.all(np.isfinite(res.numpy())) and np.all(np.isfinite(out.numpy()))
    h = base.numpy()
    assert np.all(np.isnan(h[mask])), "a kernel wrote outside its vector"


@pytest.mark.parametrize("fuse", ["none", "mgs", "full", "pair", "block4", "block8", "sweep"])
def test_gmres_reads_only_its_operands(nk, ctx, oracle, fuse):
    """A whole GMRES solve (7 iterations: ragged blocks of the blocked sweep) with u and b inside NaN guard bands."""
    d = P.bratu2d(34, 19) if fuse == "sweep" else P.bratu2d(33, 19)  # (the sweep kernels take even row lengths)
    b0 = RNG.standard_normal(d["u0"].shape)
    base, (u, b, res), mask = _guarded(nk, ctx, [d["u0"], b0, np.zeros_like(b0)])
    J = nk.JacobianOperator(nk.bratu2d_, res, u, (d["dx"], d["dy"], d["lam"]))
    ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res), memory=5)
    nk.krylov_solve_(ws, J, b, itmax=7, restart=True, history=True, fuse=fuse)
    po = P.oracle_problem(oracle, d)
    xr, sr, hr = oracle.krylov_solve(po, d["u0"], b0, memory=5, itmax=7, restart=True, hist_cap=16)
    assert ws.stats.niter == sr["niter"] == 7
    assert np.linalg.norm(ws.x.numpy() - xr) <= 1e-9 * np.linalg.norm(xr)
    assert np.all(np.isnan(base.numpy()[mask]))
    if fuse == "sweep":  # the bulk copies of the sweep kernel (rows of u with their rim columns) stayed inside u
        ctx.profile(True)
        nk.krylov_solve_(ws, J, b, itmax=7, restart=True, fuse=fuse)
        assert ctx.profile_read(13)[0] > 0
        ctx.profile(False)
