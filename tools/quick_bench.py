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
for fuse in ("none", "mgs", "full"):
    for _ in range(2):
        ctx.sync(); ctx.launch_count(reset=True)
        ctx.timer_start()
        nk.krylov_solve_(ws, J, b, rtol=1e-30, atol=0.0, restart=True, itmax=40, fuse=fuse)
        ms = ctx.timer_stop()
    it = ws.stats.niter
    ref_bytes = 2 * 8 * n * (5 * 20 * 21 / 2 + 6 * 20)
    print(f"gmres(20) fuse={fuse:5s} {it} its in {ms:8.2f} ms -> {it/ms*1e3:7.1f} it/s ; reference-op-list traffic {ref_bytes/ms/1e6:8.1f} GB/s "
          f"({ref_bytes/ms/1e6/6552.6*100:5.1f}% of peak); launches {ctx.launch_count()}")
