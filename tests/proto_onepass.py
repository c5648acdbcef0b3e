"""CPU prototype of the one-sweep-per-iteration GMRES (fuse = sweep) used to fix the scalar recurrences before they
were written in CUDA: un-normalised basis S_j = rho_j v_j, raw tangent W_k = J S_{k-1}, all projections of W_{k+1} and the
Gram row of S_k measured by the sweep that forms S_k, modified-Gram-Schmidt coefficients recovered by forward
substitution with the cached Gram matrix.  Test infrastructure: compares with the oracle's gmres (run: python
tests/proto_onepass.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import problems as P  # noqa: E402


def sym_givens(a, b):
    return O.sym_givens(a, b)


def gmres_onepass(Jmul, b, mem=20, itmax=40, rtol=1e-8, atol=1e-8, restart=True, reorth=False):
    n = b.size
    x = np.zeros(n)
    hist = []
    it = 0
    npass = 0
    eps = None
    while True:
        r0 = b - Jmul(x) if npass > 0 else b.copy()
        beta = np.sqrt(r0 @ r0)
        if npass == 0:
            eps = atol + rtol * beta
            hist.append(beta)
            rNorm = beta
        npass += 1
        if rNorm <= eps:
            return x, hist, it, True
        S = [r0]
        rho = [rNorm]   # (Krylov.jl uses rNorm of the previous cycle's recurrence? no: beta of the new residual)
        rho[0] = beta
        Gam = np.zeros((mem + 1, mem + 1))
        R = np.zeros((mem, mem))
        c = np.zeros(mem); s = np.zeros(mem); z = np.zeros(mem + 1)
        z[0] = beta
        # init sweep: W_1 = J S_0, T_0 = <S_0, W_1>
        W = Jmul(S[0])
        T = np.array([S[0] @ W])
        h = np.array([T[0] / rho[0] * (1.0 / rho[0])])
        cc = h / np.array(rho)
        K = 0
        solved = False
        inner_limit = min(mem, itmax - it)
        for k in range(1, inner_limit + 1):
            zt = W * (1.0 / rho[k - 1])
            for j in range(k):
                zt = zt - cc[j] * S[j]
            hcol = h.copy()
            if reorth:
                g2 = np.array([S[j] @ zt for j in range(k)])
                p2 = g2 / np.array(rho[:k])
                h2 = p2.copy()
                for a in range(k):
                    for j in range(a + 1, k):
                        h2[j] -= h2[a] * Gam[j, a]
                c2 = h2 / np.array(rho[:k])
                for j in range(k):
                    zt = zt - c2[j] * S[j]
                hcol = hcol + h2
            nrm = zt @ zt
            g = np.array([S[j] @ zt for j in range(k)])
            stencil = k < inner_limit
            if stencil:
                W = Jmul(zt)
                T = np.array([S[j] @ W for j in range(k)] + [zt @ W])
            S.append(zt)
            Hbis = np.sqrt(nrm)
            rho.append(Hbis)
            for a in range(k):
                Gam[k, a] = g[a] / rho[k] / rho[a]
            # Givens
            col = np.concatenate([hcol, [0.0]])
            for i in range(k - 1):
                Rt = c[i] * col[i] + s[i] * col[i + 1]
                col[i + 1] = s[i] * col[i] - c[i] * col[i + 1]
                col[i] = Rt
            ck, sk, rr = sym_givens(col[k - 1], Hbis)
            c[k - 1] = ck; s[k - 1] = sk
            col[k - 1] = rr
            R[:k, k - 1] = col[:k]
            zeta = sk * z[k - 1]
            z[k - 1] = ck * z[k - 1]
            rNorm = abs(zeta)
            hist.append(rNorm)
            K = k
            solved = rNorm <= eps
            if solved or k == inner_limit:
                break
            z[k] = zeta
            # next coefficients
            p = T / np.array(rho[:k + 1]) * (1.0 / rho[k])
            h = p.copy()
            for a in range(k + 1):
                for j in range(a + 1, k + 1):
                    h[j] -= h[a] * Gam[j, a]
            cc = h / np.array(rho[:k + 1])
        y = np.linalg.solve(np.triu(R[:K, :K]), z[:K])
        for i in range(K):
            x = x + (y[i] / rho[i]) * S[i]
        it += K
        if solved or it >= itmax or not restart:
            return x, hist, it, solved


def main():
    O.build(); O.load()
    from newtonkrylov_jl_b200 import _abi as A
    worst = 0.0
    for name, d, kw in [
        ("bratu2d_48", P.bratu2d(48, 40), dict(memory=20, itmax=60, restart=True, rtol=1e-9)),
        ("heat2d_32_s64", P.heat2d(32, dt_scale=64.0, ic="poly"), dict(memory=20, itmax=60, restart=True, rtol=1e-10)),
        ("heat2d_32_reorth", P.heat2d(32, dt_scale=64.0, ic="poly"), dict(memory=10, itmax=35, restart=True, rtol=1e-10, reorthogonalization=True)),
        ("dg_64", P.heat1d_dg(64, dt=1e-4), dict(memory=20, itmax=60, restart=True, rtol=1e-10)),
    ]:
        u0 = d["u0"].astype(np.float64)
        un = u0.copy() if d.get("scheme", 0) != A.AK_STEADY else None
        po = P.oracle_problem(O, d, un=un)
        rng = np.random.default_rng(1)
        b = rng.standard_normal(u0.shape)
        xr, sr, hr = O.krylov_solve(po, u0, b, hist_cap=256, atol=0.0, **kw)
        Jmul = lambda v: O.jvp(po, u0, v.reshape(u0.shape).copy())[0].reshape(-1)
        x, hist, it, solved = gmres_onepass(Jmul, b.reshape(-1), mem=kw["memory"], itmax=kw["itmax"], rtol=kw["rtol"], atol=0.0,
                                            reorth=bool(kw.get("reorthogonalization", False)))
        hr = np.asarray(hr)[:sr["niter"] + 1]
        hd = np.max(np.abs(np.array(hist) - hr)) / hr[0]
        xd = np.linalg.norm(x - xr.reshape(-1)) / np.linalg.norm(xr)
        worst = max(worst, hd)
        print(f"{name}: niter {it} vs {sr['niter']}, solved {solved} vs {sr['solved']}, hist dev {hd:.2e}, x dev {xd:.2e}")
    return worst


if __name__ == "__main__":
    main()
