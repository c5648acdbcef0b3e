"""Developer timing script: DG residual / tangent at 2^22 elements (BASELINE config 5 size), GB/s of the measured peak."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import newtonkrylov_jl_b200 as nk

ctx = nk.get_context(0)
lib, h = ctx.lib, ctx.h
P = lambda t: C.c_void_p(t.ptr)
N1 = 1 << 24
x1 = np.linspace(0.0, 1.0, N1)
u1 = nk.DeviceVector.from_numpy(np.sin(np.pi * x1), ctx)
un1, r1, v1, o1 = u1.copy(), u1.similar(), u1.copy(), u1.similar()
Fd = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_DG_)
prd = Fd.problem(u1, (un1, 1e-12, None, (4.0 / N1,), 0.0))


def timeit(name, fn, nbytes, reps=40):
    for _ in range(5):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    ms = ctx.timer_stop() / reps
    print(f"AK_DG_MB={os.environ.get('AK_DG_MB', '-')} {name:28s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.1f} GB/s  ({nbytes/ms/1e6/6552.6*100:5.1f}% of measured peak)")


timeit("residual DG 2^22 elements", lambda: lib.ak_residual(h, C.byref(prd), P(u1), P(r1), None), 24 * N1)
timeit("jvp DG 2^22 elements", lambda: lib.ak_jvp(h, C.byref(prd), P(u1), P(v1), P(o1)), 16 * N1)
