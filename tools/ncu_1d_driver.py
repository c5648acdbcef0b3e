"""Developer driver for ncu captures of the 1-D / DG stencil kernels at BASELINE sizes (N = 2^24 points, 2^22 elements):
a handful of launches of each residual / tangent kernel through the C ABI (nothing else), so that
`ncu -k regex:k_stencil1d|k_dg` sees steady-state launches on inputs larger than L2."""
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
u1 = nk.DeviceVector.from_numpy(4 * x1 * (1 - x1), ctx)
un1, r1, v1, o1 = u1.copy(), u1.similar(), u1.copy(), u1.similar()
F1 = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_)
pr1 = F1.problem(u1, (un1, 0.1, None, (0.2, 1.0 / (N1 - 1), nk.bc_zero_), 0.0))
prb = nk.bratu_.problem(u1, (1.0 / (N1 + 1), 3.5))
Fd = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_DG_)
prd = Fd.problem(u1, (un1, 0.01, None, (4.0 / N1,), 0.0))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for _ in range(reps):
    for pr in (pr1, prb, prd):
        nk._lib.check(lib.ak_residual(h, C.byref(pr), P(u1), P(r1), None))
        nk._lib.check(lib.ak_jvp(h, C.byref(pr), P(u1), P(v1), P(o1)))
ctx.sync()
print("ncu_1d_driver ok")
