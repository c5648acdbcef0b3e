"""Rank-emulated JFNK on slabs (CPU, test infrastructure): every rank holds its rows of the global
grid, ghost rows travel with torch.distributed send/recv in the order the library posts its NCCL
messages (nk.dist.halo_message_order), every inner product is an all-reduce.  Arithmetic comes
from the oracle applied to a slab padded with its ghost rows, so the result must equal the
single-domain oracle up to summation order."""
import numpy as np
import torch
import torch.distributed as dist

from newtonkrylov_jl_b200 import _abi as A
from newtonkrylov_jl_b200 import dist as nkdist


def exchange(v, periodic):
    rank, world = dist.get_rank(), dist.get_world_size()
    nx = v.shape[1]
    bufs = {"halo_lo": torch.zeros(nx, dtype=torch.float64), "halo_hi": torch.zeros(nx, dtype=torch.float64)}
    reqs = []
    for op, peer, what in nkdist.halo_message_order(rank, world, periodic):
        if op == "send":
            row = v[-1] if what == "last_row" else v[0]
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(row)), peer))
        else:
            reqs.append(dist.irecv(bufs[what], peer))
    for r in reqs:
        r.wait()
    down, up = nkdist.halo_neighbors(rank, world, periodic)
    lo = bufs["halo_lo"].numpy() if down >= 0 else np.zeros(nx)
    hi = bufs["halo_hi"].numpy() if up >= 0 else np.zeros(nx)
    return lo, hi


def allsum(x):
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])


class SlabBratu2D:
    """F and J v of 2-D Bratu on a slab, computed by the oracle on the slab extended by its ghost rows."""

    def __init__(self, O, nx, ny, gny, dx, dy, lam):
        self.O, self.nx, self.ny, self.gny, self.dx, self.dy, self.lam = O, nx, ny, gny, dx, dy, lam

    def _ext(self, v):
        lo, hi = exchange(v, False)
        return np.vstack([lo[None, :], v, hi[None, :]])

    def residual(self, u):
        e = self._ext(u)
        po = self.O.make_problem(A.AK_BRATU2D, self.nx, self.ny + 2, dx=self.dx, dy=self.dy, lam=self.lam)
        r, _ = self.O.residual(po, e)
        return r[1:-1]

    def jvp(self, u, v):
        eu = np.vstack([np.zeros((1, self.nx)), u, np.zeros((1, self.nx))])
        ev = self._ext(v)
        po = self.O.make_problem(A.AK_BRATU2D, self.nx, self.ny + 2, dx=self.dx, dy=self.dy, lam=self.lam)
        o, _ = self.O.jvp(po, eu, ev)
        return o[1:-1]


def gmres(prob, u, b, rtol, atol=np.sqrt(np.finfo(float).eps), itmax=None):
    """Krylov.jl gmres! (no restart) with distributed dots — same steps as oracle/nk_oracle.c:ok_gmres."""
    dot = lambda a, c: allsum(float(np.vdot(a, c)))
    n_global = allsum(float(b.size))
    itmax = int(2 * n_global) if itmax is None else itmax
    x = np.zeros_like(b)
    beta = np.sqrt(dot(b, b))
    rnorm = beta
    eps = atol + rtol * rnorm
    if beta == 0:
        return x, 0
    V, R, c, s, z = [b / rnorm], [], [], [], [beta]
    k = 0
    btol = np.finfo(float).eps ** 0.75
    while True:
        k += 1
        w = prob.jvp(u, V[k - 1])
        col = []
        for i in range(k):
            h = dot(V[i], w)
            col.append(h)
            w = w - h * V[i]
        hbis = np.sqrt(dot(w, w))
        for i in range(k - 1):
            t = c[i] * col[i] + s[i] * col[i + 1]
            col[i + 1] = s[i] * col[i] - c[i] * col[i + 1]
            col[i] = t
        ck, sk, rho = prob.O.sym_givens(col[k - 1], hbis)
        c.append(ck); s.append(sk); col[k - 1] = rho
        R.append(col)
        zeta = sk * z[k - 1]
        z[k - 1] = ck * z[k - 1]
        rnorm = abs(zeta)
        solved = rnorm <= eps or rnorm + 1.0 <= 1.0
        if solved or hbis <= btol or k >= itmax:
            break
        V.append(w / hbis)
        z.append(zeta)
    y = list(z[:k])
    for i in range(k - 1, -1, -1):
        for j in range(k - 1, i, -1):
            y[i] -= R[j][i] * y[j]
        y[i] /= R[i][i]
    for i in range(k):
        x = x + y[i] * V[i]
    return x, k


def newton(prob, u0, tol_rel=1e-6, tol_abs=1e-12, max_niter=50, eta_max=0.999, gamma=0.9):
    u = u0.copy()
    res = prob.residual(u)
    n_res = np.sqrt(allsum(float(np.vdot(res, res))))
    tol = tol_rel * n_res + tol_abs
    eta = eta_max
    hist = [dict(n_res=n_res, inner=0)]
    outer = 0
    while n_res > tol and outer <= max_niter:
        d, k = gmres(prob, u, res.copy(), eta)
        u = u - d
        prior = n_res
        res = prob.residual(u)
        n_res = np.sqrt(allsum(float(np.vdot(res, res))))
        eta = prob.O.forcing_ew(eta_max, gamma, eta, tol, n_res, prior)
        outer += 1
        hist.append(dict(n_res=n_res, inner=k))
    return u, hist
