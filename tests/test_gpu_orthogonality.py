"""Loss-of-orthogonality stress test of the Arnoldi basis.

The blocked sweeps (fuse = pair / block4 / block8, the library default) project on the un-updated `w` and correct with
cached Gram entries: algebraically modified Gram-Schmidt, numerically not the reference's operation order.  What matters
for GMRES is how orthogonal the basis stays, so this test measures ||I - V^T V||_F on a long NON-restarted basis
(150 iterations, no convergence: rtol = 1e-30, atol = 0) for every fusion level and requires the blocked levels to lose
no more than twice what the reference op list (fuse = none) loses.  The numbers go to the parity ledger.
"""
import numpy as np
import pytest

import ledger
import problems as P

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(21)
ITERS = 150

CASES = [
    ("dg_64_dt1e-3", lambda: P.heat1d_dg(64, dt=1e-3)),            # far-from-normal operator
    ("bratu2d_64", lambda: P.generic(P.bratu2d(64))),
]


def basis_loss(nk, ctx, d, b0, fuse, reorth=False):
    F_, u, p, _ = P.device_setup(nk, ctx, d)
    res = u.zero()
    J = nk.JacobianOperator(F_, res, u, p)
    b = nk.DeviceVector.from_numpy(b0, ctx)
    ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res), memory=20)
    nk.krylov_solve_(ws, J, b, rtol=1e-30, atol=0.0, itmax=ITERS, fuse=fuse, history=True,
                     reorthogonalization=reorth)
    k = ws.stats.niter
    V = np.empty((k, u.n))
    for i in range(k):
        v, scale = ws.basis(i)
        V[i] = v.numpy().reshape(-1) / scale
    G = V @ V.T
    return k, float(np.linalg.norm(np.eye(k) - G)), ws.stats.residuals[-1] / ws.stats.residuals[0]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_blocked_sweeps_keep_the_basis_as_orthogonal_as_the_reference_op_list(nk, ctx, name, make):
    d = make()
    b0 = RNG.standard_normal(d["u0"].shape)
    loss = {}
    for fuse in ("none", "mgs", "full", "pair", "block4", "block8", "sweep"):
        k, loss[fuse], red = basis_loss(nk, ctx, d, b0, fuse)
        assert k == ITERS, (fuse, k)
        ledger.record("loss_of_orthogonality", f"{name}/{fuse}", iterations=k, frobenius_I_minus_VtV=loss[fuse],
                      residual_reduction=red)
    for fuse in ("pair", "block4", "block8", "sweep"):
        assert loss[fuse] <= 2.0 * loss["none"] + 1e-13, (fuse, loss)
    # the second sweep restores orthogonality to rounding level, also in the blocked form
    for fuse in ("none", "block8", "sweep"):
        k, l2, _ = basis_loss(nk, ctx, d, b0, fuse, reorth=True)
        ledger.record("loss_of_orthogonality", f"{name}/{fuse}+reorth", iterations=k, frobenius_I_minus_VtV=l2)
        assert l2 <= max(1e-12, 1e-2 * loss["none"]), (fuse, l2, loss["none"])
