"""2-D Bratu solved to tolerance with a caller-supplied right preconditioner.

    newton_krylov!(bratu2d!, u0, (dx, dy, lambda), res; algo = :gmres, N = (J) -> P)        src/Ariadne.jl:288-372

Unpreconditioned GMRES needs O(N) iterations per Newton step on an N x N grid, which is why the
reference's own scripts precondition (`N = J -> ilu(collect(J))`, examples/bratu.jl:121-139).  Here `P` is user code
plugged into the library's `N` hook (AK_PRECOND_USER): a fast Poisson solve  y = (Laplacian + mean(lambda e^u))^-1 x
by sine transforms (torch.fft on the library's stream).  The Jacobian is Laplacian + diag(lambda e^u), so P J = I + a
small compact perturbation and every GMRES solve takes a handful of iterations at any N.

    python examples/python/bratu2d_fast_poisson.py [N=2048] [lambda=3.5]
"""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import newtonkrylov_jl_b200 as nk  # noqa: E402


def dst1(x, dim):
    """Unnormalised DST-I along `dim` through an FFT of the odd extension: X_k = sum_j x_j sin(pi j k / (n + 1))."""
    import torch

    n = x.shape[dim]
    z = torch.zeros_like(x.narrow(dim, 0, 1))
    ext = torch.cat([z, x, z, -torch.flip(x, dims=(dim,))], dim=dim)
    return -torch.fft.rfft(ext, dim=dim).imag.narrow(dim, 1, n) / 2.0


class FastPoisson:
    """y = (L + c I)^-1 x for the five-point Dirichlet Laplacian L on an nx x ny grid, c = mean(lambda e^u)."""

    def __init__(self, nx, ny, dx, dy, device):
        import torch

        kx = torch.arange(1, nx + 1, device=device, dtype=torch.float64)
        ky = torch.arange(1, ny + 1, device=device, dtype=torch.float64)
        mx = -4.0 * torch.sin(math.pi * kx / (2 * (nx + 1))) ** 2 / dx**2
        my = -4.0 * torch.sin(math.pi * ky / (2 * (ny + 1))) ** 2 / dy**2
        self.eig = my[:, None] + mx[None, :]
        self.nx, self.ny, self.shift = nx, ny, 0.0
        self.scale = 4.0 / ((nx + 1) * (ny + 1))

    def set_shift(self, c):
        self.shift = float(c)

    def __call__(self, y, x):
        X = dst1(dst1(x.view(self.ny, self.nx), 1), 0)
        X = X / (self.eig + self.shift)
        y.view(self.ny, self.nx).copy_(dst1(dst1(X, 1), 0) * self.scale)


def solve(N=2048, lam=3.5, verbose=True, ctx=None):
    import torch

    ctx = ctx or nk.get_context(0)
    dx = 1.0 / (N + 1)
    x = dx * np.arange(1, N + 1)
    u = nk.DeviceVector.from_numpy(np.sin(np.pi * x)[:, None] * np.sin(np.pi * x)[None, :], ctx)
    fp = FastPoisson(N, N, dx, dx, torch.device("cuda", ctx.device))
    P = nk.UserPreconditioner(fp, device=ctx.device)

    def N_hook(J):  # called once per Newton step, like N(J) at src/Ariadne.jl:324-326
        ut = nk.as_torch(J.u.ptr, (J.u.n,), ctx.device)
        with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream)):
            fp.set_shift(lam * torch.exp(ut).mean().item())
        return P

    hist = []
    ctx.sync()
    t0 = time.perf_counter()
    _, r = nk.newton_krylov_(nk.bratu2d_, u, (dx, dx, lam), None, N=N_hook, history=hist)
    ctx.sync()
    dt = time.perf_counter() - t0
    if verbose:
        print(f"2-D Bratu {N} x {N}, lambda = {lam}: solved = {r.solved} in {r.stats.outer_iterations} Newton steps, "
              f"{r.stats.inner_iterations} GMRES iterations, {dt:.3f} s")
        for k, h in enumerate(hist):
            print(f"  step {k}: ||F|| = {h['n_res']:.6e}  gmres its = {h['inner']}")
    return u, r, hist, dt


if __name__ == "__main__":
    solve(int(sys.argv[1]) if len(sys.argv) > 1 else 2048, float(sys.argv[2]) if len(sys.argv) > 2 else 3.5)
