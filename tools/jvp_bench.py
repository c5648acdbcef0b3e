"""Developer timing: plain and fused (normalise + JVP) 2-D Bratu stencil at 8192^2."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import newtonkrylov_jl_b200 as nk
N = 8192
ctx = nk.get_context(0); n = N * N; dx = 1.0 / (N + 1)
x = dx * np.arange(1, N + 1)
u = nk.DeviceVector.from_numpy(np.sin(np.pi * x)[:, None] * np.sin(np.pi * x)[None, :], ctx)
res, coef = u.similar(), u.similar()
prob = nk.bratu2d_.problem(u, (dx, dx, 3.5), coef=coef)
lib, h = ctx.lib, ctx.h
P = lambda t: C.c_void_p(t.ptr)
lib.ak_residual(h, C.byref(prob), P(u), P(res), None)
ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res), memory=20)
J = nk.JacobianOperator(nk.bratu2d_, res, u, (dx, dx, 3.5), coef=coef)
b = res.copy()
names = {5: "jvp", 6: "residual", 10: "mgs_pair", 11: "pair_edge", 1: "axpy_norm"}
for fuse in ("pair", "full"):
    for rep in range(2):
        ctx.sync(); ctx.profile(True); ctx.timer_start()
        nk.krylov_solve_(ws, J, b, rtol=1e-30, atol=0.0, restart=True, itmax=40, fuse=fuse)
        ms = ctx.timer_stop()
    out = {names[c]: ctx.profile_read(c) for c in names}
    ctx.profile(False)
    cnt, t = out["jvp"]
    print(f"fuse={fuse}: {40/ms*1e3:.1f} it/s; jvp avg {t/cnt*1e3:.1f} us ({(32 if True else 24)*n/(t/cnt*1e-3)/1e9:.0f} GB/s of 32n)", {k: (v[0], round(v[1], 2)) for k, v in out.items()})
v, out_ = u.copy(), u.similar()
for _ in range(3): lib.ak_jvp(h, C.byref(prob), P(u), P(v), P(out_))
ctx.sync(); ctx.timer_start()
for _ in range(20): lib.ak_jvp(h, C.byref(prob), P(u), P(v), P(out_))
ms = ctx.timer_stop() / 20
print(f"plain jvp {ms*1e3:.1f} us {24*n/ms/1e6:.0f} GB/s")
for _ in range(3): lib.ak_residual(h, C.byref(prob), P(u), P(res), None)
ctx.sync(); ctx.timer_start()
for _ in range(20): lib.ak_residual(h, C.byref(prob), P(u), P(res), None)
ms = ctx.timer_stop() / 20
print(f"residual {ms*1e3:.1f} us {24*n/ms/1e6:.0f} GB/s")
