"""Developer timing script (not the bench contract): GMRES(20) restart cycles at 8192^2 with the one-sweep kernels
(fuse = sweep) next to the eight-step blocked passes, 2-D Bratu (cached lambda e^u) and 2-D heat with
re-orthogonalisation.  `--short`: one cycle of each (for an ncu launch list of k_sweep)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import newtonkrylov_jl_b200 as nk

short = "--short" in sys.argv
N = 8192
ctx = nk.get_context(0)
n = N * N
dx = 1.0 / (N + 1)
x = dx * np.arange(1, N + 1)
u0 = np.sin(np.pi * x)[:, None] * np.sin(np.pi * x)[None, :]
u = nk.DeviceVector.from_numpy(u0, ctx)
res, coef = u.similar(), u.similar()
prob = nk.bratu2d_.problem(u, (dx, dx, 3.5), coef=coef)
lib, h = ctx.lib, ctx.h
P = lambda t: C.c_void_p(t.ptr)
lib.ak_residual(h, C.byref(prob), P(u), P(res), None)
ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res), memory=20)
J = nk.JacobianOperator(nk.bratu2d_, res, u, (dx, dx, 3.5), coef=coef)
b = res.copy()
PEAK = 6552.6


def units_sweep(m, tang, reorth):
    s = 1 + tang
    for k in range(1, m + 1):
        if reorth:
            s += k + 2
        s += k + 2 + (tang if k < m else 0)
    return s


def cycles(J, b, fuse, its, reorth=False, reps=2, label=""):
    for _ in range(reps):
        ctx.sync(); ctx.launch_count(reset=True); ctx.profile(True)
        ctx.timer_start()
        nk.krylov_solve_(ws, J, b, rtol=1e-30, atol=0.0, restart=True, itmax=its, fuse=fuse, reorthogonalization=reorth)
        ms = ctx.timer_stop()
        sw = ctx.profile_read(13)
        ctx.profile(False)
    it = ws.stats.niter
    line = f"{label} fuse={fuse:6s} {it} its in {ms:8.2f} ms -> {it/ms*1e3:7.1f} it/s ; launches {ctx.launch_count()}"
    if sw[0]:
        tang = 2.0 if "bratu" in label else 1.0
        by = 8.0 * n * units_sweep(20, tang, reorth) * (its // 20)
        line += f" ; k_sweep: {sw[0]} launches, {sw[1]:.2f} ms, {by/sw[1]/1e6:7.1f} GB/s ({by/sw[1]/1e6/PEAK*100:5.1f}% of peak)"
    print(line, flush=True)


its = 20 if short else 40
for fuse in (("sweep",) if short else ("block8", "sweep")):
    cycles(J, b, fuse, its, label="bratu2d 8192^2")
if not short:
    un = nk.DeviceVector.from_numpy(u0, ctx)
    F2 = nk.ImplicitResidual(nk.G_Euler_, nk.diffusion_)
    dt = 16.0 * dx**4 / (2 * 0.01 * 2 * dx**2)
    p2 = (un, dt, None, (0.01, dx, dx, nk.bc_zero_), 0.0)
    J2 = nk.JacobianOperator(F2, res, u, p2)
    for fuse in ("block8", "sweep"):
        cycles(J2, b, fuse, its, reorth=True, label="heat2d  8192^2 reorth")
    for fuse in ("block8", "sweep"):
        cycles(J2, b, fuse, its, reorth=False, label="heat2d  8192^2")
