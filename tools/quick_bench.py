"""Developer timing script (not the bench contract): per-kernel GB/s at 8192^2 through the C ABI."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import newtonkrylov_jl_b200 as nk
from newtonkrylov_jl_b200 import _abi as A

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = nk.get_context(0)
n = N * N
dx = 1.0 / (N + 1)
x = dx * np.arange(1, N + 1)
u0 = np.sin(np.pi * x)[:, None] * np.sin(np.pi * x)[None, :]
u = nk.DeviceVector.from_numpy(u0, ctx)
res, coef, v, w, w2, out = (u.similar() for _ in range(6))
nk.kcopy_(n, v, u); nk.kcopy_(n, w, u); nk.kcopy_(n, w2, u)
prob = nk.bratu2d_.problem(u, (dx, dx, 3.5), coef=coef)
lib, h = ctx.lib, ctx.h
P = lambda t: C.c_void_p(t.ptr)

def timeit(name, fn, nbytes, reps=20):
    for _ in range(3): fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps): fn()
    ms = ctx.timer_stop() / reps
    print(f"{name:28s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.1f} GB/s  ({nbytes/ms/1e6/6552.6*100:5.1f}% of measured peak)")

nrm = C.c_double()
timeit("residual bratu2d (+coef)", lambda: lib.ak_residual(h, C.byref(prob), P(u), P(res), None), 24 * n)
timeit("jvp bratu2d (cached coef)", lambda: lib.ak_jvp(h, C.byref(prob), P(u), P(v), P(out)), 24 * n)
prob2 = nk.bratu2d_.problem(u, (dx, dx, 3.5))
timeit("jvp bratu2d (exp recompute)", lambda: lib.ak_jvp(h, C.byref(prob2), P(u), P(v), P(out)), 24 * n)
timeit("axpy", lambda: lib.ak_axpy(h, n, 1e-9, P(v), P(w)), 24 * n)
timeit("copy", lambda: lib.ak_copy(h, n, P(w2), P(v)), 16 * n)
timeit("fill", lambda: lib.ak_fill(h, n, P(w2), 1.0), 8 * n)
d = C.c_double()
timeit("dot (incl. host sync)", lambda: lib.ak_dot(h, n, P(v), P(w), C.byref(d)), 16 * n)
timeit("nrm2 (incl. host sync)", lambda: lib.ak_nrm2(h, n, P(v), C.byref(d)), 8 * n)

# GMRES(20) restart cycles on J(u0): protocol A of SURVEY 8d
ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res), memory=20)
J = nk.JacobianOperator(nk.bratu2d_, res, u, (dx, dx, 3.5), coef=coef)
lib.ak_residual(h, C.byref(prob), P(u), P(res), None)
b = res.copy()
for fuse in ("none", "mgs", "full", "pair", "block4", "block8", "sweep"):
    for _ in range(2):
        ctx.sync(); ctx.launch_count(reset=True)
        ctx.timer_start()
        nk.krylov_solve_(ws, J, b, rtol=1e-30, atol=0.0, restart=True, itmax=40, fuse=fuse)
        ms = ctx.timer_stop()
    it = ws.stats.niter
    ref_bytes = 2 * 8 * n * (5 * 20 * 21 / 2 + 6 * 20)
    print(f"gmres(20) fuse={fuse:6s} {it} its in {ms:8.2f} ms -> {it/ms*1e3:7.1f} it/s ; reference-op-list traffic {ref_bytes/ms/1e6:8.1f} GB/s "
          f"({ref_bytes/ms/1e6/6552.6*100:5.1f}% of peak); launches {ctx.launch_count()}")

# ---- other BASELINE configs at full size: kernel GB/s (C2 heat 1-D 2^24, C3 heat 2-D 8192^2, C5 DG 2^22 elements)
import gc
del ws, J, b; gc.collect()
un = nk.DeviceVector.from_numpy(u0, ctx)
F2 = nk.ImplicitResidual(nk.G_Euler_, nk.diffusion_)
p2 = (un, 1e-9, un.zero(), (0.01, dx, dx, nk.bc_zero_), 0.0)
pr2 = F2.problem(u, p2)
timeit("residual heat2d 8192^2", lambda: lib.ak_residual(h, C.byref(pr2), P(u), P(res), None), 24 * n)
timeit("jvp heat2d 8192^2", lambda: lib.ak_jvp(h, C.byref(pr2), P(u), P(v), P(out)), 16 * n)
N1 = 1 << 24
x1 = np.linspace(0.0, 1.0, N1)
u1 = nk.DeviceVector.from_numpy(4 * x1 * (1 - x1), ctx)
un1, r1, v1, o1 = u1.copy(), u1.similar(), u1.copy(), u1.similar()
F1 = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_)
pr1 = F1.problem(u1, (un1, 1e-12, None, (0.2, 1.0 / (N1 - 1), nk.bc_zero_), 0.0))
timeit("residual heat1d 2^24", lambda: lib.ak_residual(h, C.byref(pr1), P(u1), P(r1), None), 24 * N1)
timeit("jvp heat1d 2^24", lambda: lib.ak_jvp(h, C.byref(pr1), P(u1), P(v1), P(o1)), 16 * N1)
prb = nk.bratu_.problem(u1, (1.0 / (N1 + 1), 3.5))
timeit("residual bratu1d 2^24", lambda: lib.ak_residual(h, C.byref(prb), P(u1), P(r1), None), 16 * N1)
timeit("jvp bratu1d 2^24 (exp)", lambda: lib.ak_jvp(h, C.byref(prb), P(u1), P(v1), P(o1)), 24 * N1)
Fd = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_DG_)
prd = Fd.problem(u1, (un1, 1e-12, None, (4.0 / N1,), 0.0))
timeit("residual DG 2^22 elements", lambda: lib.ak_residual(h, C.byref(prd), P(u1), P(r1), None), 24 * N1)
timeit("jvp DG 2^22 elements", lambda: lib.ak_jvp(h, C.byref(prd), P(u1), P(v1), P(o1)), 16 * N1)

# ---- CG (algo = :cg of examples/bratu.jl:59-108) on 1-D Bratu N = 2^24, lambda = 1: bytes per iteration 88n
#      (tangent + <p,Ap> 24n, r -= alpha Ap + <r,r> 24n, x += alpha p ; p = r + beta p 40n)
prc = nk.bratu_.problem(u1, (1.0 / (N1 + 1), 1.0))
wsc = nk.krylov_workspace("cg", nk.KrylovConstructor(r1))
Jc = nk.JacobianOperator(nk.bratu_, r1, u1, (1.0 / (N1 + 1), 1.0))
lib.ak_residual(h, C.byref(prc), P(u1), P(r1), None)
bc = r1.copy()
for _ in range(2):
    ctx.sync(); ctx.launch_count(reset=True); ctx.timer_start()
    nk.krylov_solve_(wsc, Jc, bc, rtol=1e-30, atol=0.0, itmax=100)
    ms = ctx.timer_stop()
it = wsc.stats.niter
print(f"cg bratu1d 2^24 {it} its in {ms:8.2f} ms -> {it/ms*1e3:7.1f} it/s ; 88n bytes/it -> {88*N1*it/ms/1e6:8.1f} GB/s "
      f"({88*N1*it/ms/1e6/6552.6*100:5.1f}% of peak); launches {ctx.launch_count()}")
# multi-RHS tangent (mul!(Out, J, V), 8 columns) at 8192^2 vs 8 single launches
ncol = 8
Vb = nk.DeviceVector(ctx, (ncol, n)); Ob = nk.DeviceVector(ctx, (ncol, n))
nk.kfill_(Vb, 1.0)
timeit("jvp bratu2d x8 multi-RHS (exp)", lambda: lib.ak_jvp_batched(h, C.byref(prob2), P(u), P(Vb), n, P(Ob), n, ncol), (16 * ncol + 8) * n, reps=5)
timeit("jvp bratu2d x8 single launches", lambda: [lib.ak_jvp(h, C.byref(prob2), P(u), C.c_void_p(Vb.ptr + 8 * n * c), C.c_void_p(Ob.ptr + 8 * n * c)) for c in range(ncol)], 24 * ncol * n, reps=5)
