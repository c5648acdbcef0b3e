"""CPU model of the one-sweep GMRES recurrences (fuse = sweep, csrc/sweep.cu + k_gmres_sweep_scalar): un-normalised basis,
raw tangent W = J S_{k-1}, all projections of the next W and the Gram row of S_k measured by the sweep that forms S_k,
modified-Gram-Schmidt multipliers by forward substitution with the cached Gram matrix.  The NumPy model
(tests/proto_onepass.py) must reproduce the oracle's gmres! histories (Krylov.jl semantics, src/Ariadne.jl:338) to
rounding: this pins the ALGORITHM the CUDA kernels implement without a GPU; tests/test_gpu_sweep.py pins the kernels."""
import numpy as np
import pytest

import problems as P
import proto_onepass as M
from newtonkrylov_jl_b200 import _abi as A

CASES = [
    ("bratu2d_restart", lambda: P.bratu2d(48, 40), dict(memory=20, itmax=60, restart=True, rtol=1e-9)),
    ("heat2d_stiff", lambda: P.heat2d(32, dt_scale=64.0, ic="poly"), dict(memory=20, itmax=60, restart=True, rtol=1e-10)),
    ("heat2d_reorth", lambda: P.heat2d(32, dt_scale=64.0, ic="poly"),
     dict(memory=10, itmax=35, restart=True, rtol=1e-10, reorthogonalization=True)),
    ("dg", lambda: P.heat1d_dg(64, dt=1e-4), dict(memory=20, itmax=60, restart=True, rtol=1e-10)),
    ("bratu1d_short_cycles", lambda: P.generic(P.bratu1d(200)), dict(memory=6, itmax=30, restart=True, rtol=1e-10)),
]


@pytest.mark.parametrize("name,make,kw", CASES, ids=[c[0] for c in CASES])
def test_onepass_recurrences_reproduce_the_oracle(oracle, name, make, kw):
    d = make()
    u0 = d["u0"].astype(np.float64)
    un = u0.copy() if d.get("scheme", 0) != A.AK_STEADY else None
    po = P.oracle_problem(oracle, d, un=un)
    b = np.random.default_rng(1).standard_normal(u0.shape)
    xr, sr, hr = oracle.krylov_solve(po, u0, b, hist_cap=256, atol=0.0, **kw)

    def jmul(v):
        return oracle.jvp(po, u0, v.reshape(u0.shape).copy())[0].reshape(-1)

    x, hist, it, solved = M.gmres_onepass(jmul, b.reshape(-1), mem=kw["memory"], itmax=kw["itmax"], rtol=kw["rtol"], atol=0.0,
                                          reorth=bool(kw.get("reorthogonalization", False)))
    hr = np.asarray(hr)[: sr["niter"] + 1]
    assert it == sr["niter"] and solved == sr["solved"]
    assert np.max(np.abs(np.array(hist) - hr)) <= 1e-12 * hr[0]
    assert np.linalg.norm(x - xr.reshape(-1)) <= 1e-11 * np.linalg.norm(xr)
