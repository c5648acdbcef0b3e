import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import newtonkrylov_jl_b200 as nk
from newtonkrylov_jl_b200 import _abi as A
import oracle as O
from test_gpu_user import bvp_setup, bvp_np, bvp_torch
ctx = nk.get_context(0)
n, h, tv, tvdag, U0 = bvp_setup(101)
Fn, Jn = bvp_np(n, h, tv, tvdag); Ft, Jt = bvp_torch(n, h, tv, tvdag)
po = O.make_user_problem(2*n, Fn, Jn)
F_ = nk.UserResidual(Ft, Jt)
u = nk.DeviceVector.from_numpy(U0, ctx); res = u.zero()
F_(res, u, None)
J = nk.JacobianOperator(F_, res, u, None)
b0 = res.numpy().copy()
rel = lambda a,b: np.linalg.norm(a-b)/np.linalg.norm(b)
for fuse in ("none","mgs","full","block4"):
    ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res))
    nk.krylov_solve_(ws, J, nk.DeviceVector.from_numpy(b0, ctx), rtol=1e-8, itmax=60, history=True, fuse=fuse)
    xr, sr, hr = O.krylov_solve(po, U0, b0, rtol=1e-8, itmax=60, hist_cap=100)
    hh=np.array(ws.stats.residuals)
    print("plain gmres", fuse, ws.stats.niter, sr["niter"], rel(ws.x.numpy(), xr), np.max(np.abs(hh-hr[:len(hh)]))/hr[0])
y = u.zero()
nk.precond_apply_(y, nk.GmresPreconditioner(J, 30), nk.DeviceVector.from_numpy(b0, ctx))
yr = O.precond_apply(po, U0, 1, b0, itmax=30)
print("inner gmres apply", rel(y.numpy(), yr))
for algo in ("gmres", "fgmres"):
    ws = nk.krylov_workspace(algo, nk.KrylovConstructor(res))
    nk.krylov_solve_(ws, J, nk.DeviceVector.from_numpy(b0, ctx), rtol=0.0312, history=True, N=nk.GmresPreconditioner(J, 30))
    xr, sr, hr = O.krylov_solve(po, U0, b0, algo=A.AK_ALGO_GMRES if algo=="gmres" else A.AK_ALGO_FGMRES, rtol=0.0312, hist_cap=100, precond_n=1, precond_itmax=30)
    print(algo, ws.stats.niter, sr["niter"], rel(ws.x.numpy(), xr)); print(np.array(ws.stats.residuals)); print(hr)
