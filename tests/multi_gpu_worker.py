"""Run under torchrun with one rank per GPU: slab-decomposed kernels and solves through the C ABI
(NCCL all-reduce + halo exchange issued by the library) against the single-domain oracle."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import torch
import torch.distributed as dist

import newtonkrylov_jl_b200 as nk
import oracle as O
import problems as P
from newtonkrylov_jl_b200 import _abi as A


def rel(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300))


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = nk.dist.init_distributed(local)
    assert (ctx.rank, ctx.nranks) == (rank, world)
    rng = np.random.default_rng(0)
    run_cases(ctx, rank, world, rng, "nccl")
    # same cases again with the ranks' peer memory mapped: the pair-wise sweep now reduces through NVLink
    # mailboxes and pushes ghost rows with peer stores (ragged widths keep the NCCL ghost-row exchange)
    ctx.enable_p2p(64)
    assert ctx.p2p
    run_cases(ctx, rank, world, rng, "p2p")
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()


def run_1d_cases(ctx, rank, world, rng, tag):
    """1-D problems on segments (BASELINE config 5 runs the DG problem on 1 and 8 GPUs)."""
    import ctypes as C
    for name, d in [("bratu1d", P.generic(P.bratu1d(203))), ("heat1d", P.heat1d(98)),
                    ("dg", P.heat1d_dg(max(26, 3 * world), dt=1e-4))]:
        gn = d["nx"]
        if d["kind"] == A.AK_HEAT1D_DG:  # whole elements per rank
            e0, ne = nk.dist.slab_partition(gn // 4, world, rank)
            g0, n = 4 * e0, 4 * ne
        else:
            g0, n = nk.dist.slab_partition(gn, world, rank)
        sl = slice(g0, g0 + n)
        u = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
        v0 = rng.standard_normal(gn)
        v = nk.DeviceVector.from_numpy(v0[sl], ctx)
        if d["kind"] == A.AK_BRATU1D:
            F_, p, po = nk.bratu_, (d["dx"], d["lam"]), P.oracle_problem(O, d)
        else:
            un = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
            f_ = nk.heat_1D_ if d["kind"] == A.AK_HEAT1D else nk.heat_1D_DG_
            pin = (d["a"], d["dx"], nk.bc_zero_) if d["kind"] == A.AK_HEAT1D else (d["dx"],)
            F_ = nk.ImplicitResidual(nk.G_Euler_, f_)
            p = (un, d["dt"], un.zero(), pin, 0.0)
            po = P.oracle_problem(O, d, un=d["u0"])
        res, out = u.zero(), u.zero()
        F_(res, u, p)
        ref, u_after = O.residual(po, d["u0"])
        exact = d["kind"] != A.AK_BRATU1D
        assert (np.array_equal(res.numpy(), ref[sl]) if exact else rel(res.numpy(), ref[sl]) < 1e-14), name
        assert np.array_equal(u.numpy(), u_after[sl]), name
        nk.mul_(out, nk.JacobianOperator(F_, res, u, p), v)
        refj, v_after = O.jvp(po, d["u0"], v0)
        assert (np.array_equal(out.numpy(), refj[sl]) if exact else rel(out.numpy(), refj[sl]) < 1e-14), name
        assert np.array_equal(v.numpy(), v_after[sl]), name
        # one linear solve with every fusion level
        b0 = rng.standard_normal(gn)
        if d["kind"] == A.AK_HEAT1D:
            b0[0] = b0[-1] = 0.0
        # (a fixed number of iterations: 1-D Laplacians converge only when the Krylov space is exhausted, which
        #  makes the final count a knife edge)
        xr, sr, hr = O.krylov_solve(po, d["u0"], b0, rtol=1e-8, itmax=40, hist_cap=100000)
        for fuse in ("none", "mgs", "full", "pair", "block4", "block8", "sweep"):
            u = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
            ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res))
            nk.krylov_solve_(ws, nk.JacobianOperator(F_, res, u, p), nk.DeviceVector.from_numpy(b0[sl], ctx),
                             rtol=1e-8, itmax=40, history=True, fuse=fuse)
            assert (ws.stats.niter, ws.stats.solved) == (sr["niter"], sr["solved"]), (name, fuse, ws.stats.niter, sr)
            assert np.max(np.abs(np.array(ws.stats.residuals) - hr)) <= 1e-9 * hr[0], (name, fuse)
            assert rel(ws.x.numpy(), xr[sl]) < 1e-7, (name, fuse)
        if rank == 0:
            print(f"[multi-gpu x{world} {tag}] {name}: ok", flush=True)


def run_scheme_cases(ctx, rank, world, rng, tag):
    """G_Midpoint! / G_Trapezoid! (examples/implicit.jl:17-37) on slabs / segments: residual, tangent, two time steps."""
    import ctypes as C
    cases = [("heat1d", P.heat1d(98), nk.heat_1D_), ("dg", P.heat1d_dg(max(26, 3 * world), dt=1e-4), nk.heat_1D_DG_),
             ("heat2d", P.heat2d(36, dt_scale=48.0, ic="poly"), nk.diffusion_)]
    for name, d0, f_ in cases:
        for G_, code in ((nk.G_Midpoint_, A.AK_MIDPOINT), (nk.G_Trapezoid_, A.AK_TRAPEZOID)):
            d = dict(d0)
            d["scheme"] = code
            if d["kind"] == A.AK_HEAT2D:
                gy0, ny = nk.dist.slab_partition(d["ny"], world, rank)
                sl = slice(gy0, gy0 + ny)
                pin = (d["a"], d["dx"], d["dy"], nk.bc_zero_, d["ny"], gy0)
            elif d["kind"] == A.AK_HEAT1D_DG:
                e0, ne = nk.dist.slab_partition(d["nx"] // 4, world, rank)
                sl = slice(4 * e0, 4 * (e0 + ne))
                pin = (d["dx"],)
            else:
                g0, n = nk.dist.slab_partition(d["nx"], world, rank)
                sl = slice(g0, g0 + n)
                pin = (d["a"], d["dx"], nk.bc_zero_)
            # a state that differs from u_n, so that both halves of the schemes are exercised
            ustate = d["u0"] * (1.0 + 0.1 * np.cos(np.arange(d["u0"].size).reshape(d["u0"].shape) * 0.37))
            u = nk.DeviceVector.from_numpy(ustate[sl], ctx)
            un = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
            F_ = nk.ImplicitResidual(G_, f_)
            p = (un, d["dt"], un.zero(), pin, 0.0)
            po = P.oracle_problem(O, d, un=d["u0"].copy())
            res, out = u.zero(), u.zero()
            F_(res, u, p)
            ref, _ = O.residual(po, ustate)
            assert rel(res.numpy(), ref[sl]) < 1e-13, (name, code)
            v0 = rng.standard_normal(d["u0"].shape)
            if d["kind"] == A.AK_HEAT1D:
                v0[0] = v0[-1] = 0.0
            nk.mul_(out, nk.JacobianOperator(F_, res, u, p), nk.DeviceVector.from_numpy(v0[sl], ctx))
            refj, _ = O.jvp(po, ustate, v0)
            assert rel(out.numpy(), refj[sl]) < 1e-13, (name, code)
            # two time steps through ak_implicit_solve
            prob = F_.problem(u, p)
            o = nk.host._newton_opts(1e-6, 6e-6, 50, nk.EisenstatWalker(), "gmres", 20, 0, {})
            nsteps = 2
            newt, inner, solved = np.zeros(nsteps, np.int32), np.zeros(nsteps, np.int64), np.zeros(nsteps, np.int32)
            un2 = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
            nk._lib.check(ctx.lib.ak_implicit_solve(ctx.h, C.byref(prob), C.c_void_p(un2.ptr), nsteps, C.byref(o),
                                                    newt.ctypes.data_as(A.c_int32_p), inner.ctypes.data_as(A.c_int64_p),
                                                    solved.ctypes.data_as(A.c_int32_p)))
            po2 = P.oracle_problem(O, d, un=d["u0"].copy())
            ur, nr, ir, sr_ = O.implicit_solve(po2, d["u0"], nsteps, o)
            assert list(newt) == list(nr) and list(solved) == list(sr_), (name, code, newt, nr)
            assert all(abs(int(a) - int(b)) <= 1 for a, b in zip(inner, ir)), (name, code, inner, ir)
            assert rel(un2.numpy(), ur[sl]) < 1e-7, (name, code)
        if rank == 0:
            print(f"[multi-gpu x{world} {tag}] {name} midpoint/trapezoid: ok", flush=True)


def run_linear_extras(ctx, rank, world, rng, tag):
    """Collectives outside the plain blocked sweep, on slabs: CG (two all-reduces per iteration: NCCL, or the mailbox
    all-reduce kernel with peer memory), and the blocked sweeps with re-orthogonalisation + restart (memory = 5)."""
    d = P.bratu2d(48, 40, lam=1.0)
    gy0, ny = nk.dist.slab_partition(d["ny"], world, rank)
    sl = slice(gy0, gy0 + ny)
    po = P.oracle_problem(O, d)
    b0 = rng.standard_normal(d["u0"].shape)
    p = (d["dx"], d["dy"], d["lam"], d["ny"], gy0)

    def solve(algo, memory=20, **kw):
        u = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
        res = u.zero()
        ws = nk.krylov_workspace(algo, nk.KrylovConstructor(res), memory=memory)
        nk.krylov_solve_(ws, nk.JacobianOperator(nk.bratu2d_, res, u, p), nk.DeviceVector.from_numpy(b0[sl], ctx),
                         history=True, **kw)
        return ws.x.numpy(), ws.stats

    xr, sr, hr = O.krylov_solve(po, d["u0"], b0, algo=A.AK_ALGO_CG, rtol=1e-8, hist_cap=100000)
    x, st = solve("cg", rtol=1e-8)
    assert (st.niter, st.solved) == (sr["niter"], sr["solved"]), ("cg", st.niter, sr)
    assert np.max(np.abs(np.array(st.residuals) - hr)) <= 1e-9 * hr[0] and rel(x, xr[sl]) < 1e-8, "cg"
    for fuse in ("mgs", "pair", "block8"):
        kw = dict(rtol=1e-9, restart=True, reorthogonalization=True, itmax=33)
        xr, sr, hr = O.krylov_solve(po, d["u0"], b0, memory=5, hist_cap=1000, **kw)
        x, st = solve("gmres", memory=5, fuse=fuse, **kw)
        assert (st.niter, st.solved, st.npass) == (sr["niter"], sr["solved"], sr["npass"]), (fuse, st.niter, sr)
        assert np.max(np.abs(np.array(st.residuals) - hr)) <= 1e-9 * hr[0] and rel(x, xr[sl]) < 1e-8, fuse
    # Krylov.jl's default itmax = 2n counts the unknowns of the WHOLE system: with local sizes that differ between the
    # ranks (here 6 * world + 1 points) every rank must stop at the same iteration (a rank-local 2n deadlocked 8 ranks)
    d1 = P.generic(P.bratu1d(6 * world + 1))
    g0, n1 = nk.dist.slab_partition(d1["nx"], world, rank)
    s1 = slice(g0, g0 + n1)
    po1 = P.oracle_problem(O, d1)
    b1 = rng.standard_normal(d1["nx"])
    kw = dict(rtol=1e-15, atol=0.0, restart=True)  # GMRES(3) stagnates: the solve runs into the default itmax
    xr, sr, hr = O.krylov_solve(po1, d1["u0"], b1, memory=3, hist_cap=1000, **kw)
    for fuse in ("mgs", "block8"):
        u = nk.DeviceVector.from_numpy(d1["u0"][s1], ctx)
        res = u.zero()
        ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res), memory=3)
        nk.krylov_solve_(ws, nk.JacobianOperator(nk.bratu_, res, u, (d1["dx"], d1["lam"])),
                         nk.DeviceVector.from_numpy(b1[s1], ctx), fuse=fuse, **kw)
        assert ws.stats.niter == sr["niter"] and (sr["solved"] or sr["niter"] == 2 * d1["nx"]), (fuse, ws.stats.niter, sr)
    if rank == 0:
        print(f"[multi-gpu x{world} {tag}] cg + reorthogonalised blocked gmres + default itmax: ok", flush=True)


def run_cases(ctx, rank, world, rng, tag):
    run_1d_cases(ctx, rank, world, rng, tag)
    run_linear_extras(ctx, rank, world, rng, tag)
    if tag == "nccl":
        run_scheme_cases(ctx, rank, world, rng, tag)
    for name, d, bc in [("bratu2d", P.generic(P.bratu2d(48, 40)), nk.bc_zero_),
                        ("bratu2d_ragged", P.generic(P.bratu2d(37, 29)), nk.bc_zero_),
                        ("heat2d", P.heat2d(36, dt_scale=48.0, ic="poly"), nk.bc_zero_),
                        ("heat2d_periodic", P.heat2d(32, dt_scale=16.0, bc=A.AK_BC_PERIODIC, ic="poly"), nk.bc_periodic_)]:
        nx, gny = d["nx"], d["ny"]
        gy0, ny = nk.dist.slab_partition(gny, world, rank)
        sl = slice(gy0, gy0 + ny)
        u = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
        v0 = rng.standard_normal(d["u0"].shape)
        v = nk.DeviceVector.from_numpy(v0[sl], ctx)
        if d["kind"] == A.AK_BRATU2D:
            F_, p = nk.bratu2d_, (d["dx"], d["dy"], d["lam"], gny, gy0)
            po = P.oracle_problem(O, d)
        else:
            un = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
            F_ = nk.ImplicitResidual(nk.G_Euler_, nk.diffusion_)
            p = (un, d["dt"], un.zero(), (d["a"], d["dx"], d["dy"], bc, gny, gy0), 0.0)
            po = P.oracle_problem(O, d, un=d["u0"])
        # kernel level: residual, JVP, global reductions
        res, out = u.zero(), u.zero()
        F_(res, u, p)
        ref, _ = O.residual(po, d["u0"])
        assert rel(res.numpy(), ref[sl]) < 1e-14, name
        nk.mul_(out, nk.JacobianOperator(F_, res, u, p), v)
        refj, _ = O.jvp(po, d["u0"], v0)
        assert rel(out.numpy(), refj[sl]) < 1e-14, name
        assert abs(nk.kdot(u.n, u, v) - float(np.vdot(d["u0"], v0))) <= 1e-12 * np.linalg.norm(d["u0"]) * np.linalg.norm(v0)
        assert abs(nk.knorm(v.n, v) - np.linalg.norm(v0)) <= 1e-13 * np.linalg.norm(v0)
        # Jacobi as right and as left preconditioner (hooks N / M): local kernel, global reductions
        b0 = rng.standard_normal(d["u0"].shape)
        for side, kw, okw in (("N", "N", dict(precond_n=A.AK_PRECOND_JACOBI)), ("M", "M", dict(precond_m=A.AK_PRECOND_JACOBI))):
            xr, sr, hr = O.krylov_solve(po, d["u0"], b0, rtol=1e-8, itmax=60, hist_cap=100, **okw)
            uu = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
            J = nk.JacobianOperator(F_, res, uu, p)
            ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res))
            nk.krylov_solve_(ws, J, nk.DeviceVector.from_numpy(b0[sl], ctx), rtol=1e-8, itmax=60, history=True,
                             **{kw: nk.JacobiPreconditioner(J)})
            assert (ws.stats.niter, ws.stats.solved) == (sr["niter"], sr["solved"]), (name, side, ws.stats.niter, sr)
            assert np.max(np.abs(np.array(ws.stats.residuals) - hr)) <= 1e-9 * hr[0], (name, side)
            assert rel(ws.x.numpy(), xr[sl]) < 1e-7, (name, side)
        # solver level
        if d["kind"] == A.AK_BRATU2D:
            for fuse in ("none", "mgs", "full", "pair", "block4", "block8", "sweep"):
                u = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
                hist = []
                ctx.profile(True)
                _, r = nk.newton_krylov_native_(F_, u, p, None, history=hist, krylov_kwargs=dict(fuse=fuse))
                nsweep = ctx.profile_read(13)[0]
                ctx.profile(False)
                # the one-sweep kernels really ran on the slabs (peer memory, even row length), and only there
                assert (nsweep > 0) == (fuse == "sweep" and tag == "p2p" and nx % 2 == 0), (name, fuse, tag, nsweep)
                ur, sr, hr = O.newton(po, d["u0"])
                assert r.solved and r.stats.outer_iterations == sr["outer_iterations"], (name, fuse, r, sr)
                assert [h["inner"] for h in hist] == [h["inner"] for h in hr], (name, fuse)
                for a, b in zip(hist, hr):
                    assert abs(a["n_res"] - b["n_res"]) <= 1e-8 * b["n_res"] + 1e-13 * hr[0]["n_res"], (name, fuse, a, b)
                assert rel(u.numpy(), ur[sl]) < 1e-7, (name, fuse)
        else:
            prob = F_.problem(u, p)
            import ctypes as C
            # peer memory: the pair-wise sweep and the one-sweep kernels (ghost rows of every basis vector pushed by
            # the neighbours; with and without the second Gram-Schmidt sweep)
            variants = [dict(fuse="pair"), dict(fuse="sweep"), dict(fuse="sweep", reorthogonalization=True)] if tag == "p2p" else [{}]
            for kk in variants:
                o = nk.host._newton_opts(1e-6, 6e-6, 50, nk.EisenstatWalker(), "gmres", 20, 0, dict(kk))
                nsteps = 2
                newt, inner, solved = np.zeros(nsteps, np.int32), np.zeros(nsteps, np.int64), np.zeros(nsteps, np.int32)
                un = nk.DeviceVector.from_numpy(d["u0"][sl], ctx)
                nk._lib.check(ctx.lib.ak_implicit_solve(ctx.h, C.byref(prob), C.c_void_p(un.ptr), nsteps, C.byref(o),
                                                        newt.ctypes.data_as(A.c_int32_p), inner.ctypes.data_as(A.c_int64_p),
                                                        solved.ctypes.data_as(A.c_int32_p)))
                ur, nr, ir, sr_ = O.implicit_solve(po, d["u0"], nsteps, o)
                assert list(newt) == list(nr) and list(solved) == list(sr_), (name, kk, newt, nr)
                assert all(abs(int(a) - int(b)) <= 1 for a, b in zip(inner, ir)), (name, kk, inner, ir)
                assert rel(un.numpy(), ur[sl]) < 1e-7, (name, kk)
        if rank == 0:
            print(f"[multi-gpu x{world} {tag}] {name}: ok", flush=True)


if __name__ == "__main__":
    main()
