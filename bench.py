#!/usr/bin/env python
"""bench.py — GMRES iterations/s of the JFNK inner loop (default: 2-D Bratu 8192^2 fp64 per GPU).

Contract (driver):  python bench.py --gpus N --steps K --warmup W   (N > 1 under torchrun)
prints ONE JSON line from rank 0.

Workloads (`--config`, BASELINE.json configs):
  c4 (default, the headline metric)  2-D Bratu, lambda = 3.5, u0 = sin(pi x) sin(pi y) on the global unit square;
        8192 x 8192 unknowns PER GPU (weak scaling, slab decomposition along y).
  c3    2-D implicit-Euler heat 8192^2 (examples/heat_2D.jl), non-eigenfunction IC, dt = 16 x the example's rule,
        `reorthogonalization = true` (heat_2D.jl:131).
  c2    1-D implicit-Euler heat, N = 2^24 points, dt = 0.1 (examples/heat_1D.jl), one GPU.
  c5    DG 1-D heat, 2^22 elements x 4 LGL nodes PER GPU, periodic, dt = 0.01 (examples/heat_1D_DG.jl).
One STEP = one Newton step of `newton_krylov!` (src/Ariadne.jl:321-368) with
`krylov_kwargs = (; restart = true, itmax = 40, rtol = 1e-30, atol = 0)` and memory = 20: copy(res) -> GMRES(20) x 2
restart cycles (40 iterations, each = 1 JVP + k modified-Gram-Schmidt steps + Givens update; with the default
--fuse sweep an iteration is ONE pass over the basis, csrc/sweep.cu) -> u .-= d -> F!(res, u) + norm(res).  The tolerance is set so that every step does exactly 40 iterations (fixed work per step); the
time-dependent configs (c2, c3, c5) restart every step from u = u_n (one extra copy + residual, inside the timed
region), c4 keeps iterating on the same Newton sequence.

value  = GMRES iterations/s per slab, summed over the slabs (= GPUs): 40*K*N / t, inputs resident in HBM, timed with
         CUDA events on the library's stream, max over ranks.
e2e    = the same metric through the host-buffer entry point ak_newton_solve_host (what a Julia caller holding an
         Array{Float64} calls): every step copies u host->device from pinned memory, allocates the Krylov workspace
         like the reference does per newton_krylov! call, runs the same Newton step and copies u back.
roofline = the dominant kernel (--fuse sweep: k_sweep, 8n(k + 4) bytes at basis size k, bytes and time averaged over the
         launches of a step; --fuse block8: the full blocked pass, 144n bytes per launch) timed live with CUDA events
         inside the timed region (library profiler), plus the step-level figure: algorithmic bytes of the whole step /
         step time.  `traffic` is a COMMITTED constant (profiles/kernel_traffic.json, from the ncu captures of the same
         kernel), not measured here.
cpu_baseline = the CPU oracle (oracle/nk_oracle.c, a port of the reference algorithm) on all host cores, one
         GMRES(20) restart cycle of the same solve; at every N (rank 0 runs it after the GPU legs).
other_configs (default c4 line only) = c2, c3, c5 at N = 1 and c5 at N > 1: it/s and per-kernel GB/s.

--impl reference times the reference's CPU path (the oracle port — Julia is not installed and the reference's
dependencies are not vendored, so oracle/_ref does not exist) on the same workload, with the OpenMP team set to the
host's cores explicitly (torchrun exports OMP_NUM_THREADS=1).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAMBDA = 3.5
MEMORY = 20
ITMAX = 40
METRIC = "gmres_iters_per_sec"
FUSE_CODE = {"none": 0, "mgs": 1, "full": 2, "pair": 3, "block4": 4, "block8": 5, "sweep": 6}
SWEEP_KMAX = 24  # kSwKMax of csrc/sweep.h
LGL = np.array([-1.0, -1.0 / np.sqrt(5.0), 1.0 / np.sqrt(5.0), 1.0])

CONFIGS = {
    "c4": dict(kind="bratu2d", nx=8192, ny=8192, reorth=False, jvp_bytes=24, res_bytes=24,
               unit="GMRES it/s per 8192^2 slab, summed over GPUs",
               desc="2D Bratu {nx}x{ny} fp64 per GPU (lambda=3.5, u0=sin(pi x)sin(pi y))"),
    "c3": dict(kind="heat2d", nx=8192, ny=8192, reorth=True, jvp_bytes=16, res_bytes=24, dt_scale=16.0, a=0.01,
               unit="GMRES it/s per 8192^2 slab, summed over GPUs",
               desc="2D implicit-Euler heat {nx}x{ny} fp64 per GPU (a=0.01, IC 16x(1-x)y(1-y), dt=16*dx^2dy^2/(2a(dx^2+dy^2)), "
                    "bc_zero!, reorthogonalization=true)"),
    "c2": dict(kind="heat1d", nx=1 << 24, ny=1, reorth=False, jvp_bytes=16, res_bytes=24, dt=0.1, a=0.2,
               unit="GMRES it/s on N=2^24 points",
               desc="1D implicit-Euler heat N={nx} points fp64 (a=0.2, dt=0.1, IC 4x(1-x), bc!)"),
    "c5": dict(kind="dg", nx=4 << 22, ny=1, reorth=False, jvp_bytes=16, res_bytes=24, dt=0.01,
               unit="GMRES it/s per 2^22-element segment, summed over GPUs",
               desc="DG 1D heat, {ne} elements x 4 LGL nodes fp64 per GPU (periodic, IC sin(pi x), implicit Euler, dt=0.01)"),
}


def workload_config(name, cfg, n_gpus, fuse):
    nx, ny = cfg["nx"], cfg["ny"]
    return {
        "workload": cfg["desc"].format(nx=nx, ny=ny, ne=nx // 4) +
                    f", one Newton step/step: GMRES(restart, memory={MEMORY}, itmax={ITMAX}) + update + residual",
        "baseline_config": name,
        "grid_per_gpu": [nx, ny],
        "global_grid": [nx, ny * n_gpus] if ny > 1 else [nx * n_gpus, 1],
        "gmres_iters_per_step": ITMAX,
        "decomposition": ("slab along y / element segments; Arnoldi sums and ghost rows through NVLink peer memory fused "
                          "into the Gram-Schmidt kernels (NCCL only at cycle boundaries)") if n_gpus > 1 else "single GPU",
        "l2": "inputs larger than L2 (each vector is >= 128 MiB, L2 is 126 MB): no flush needed",
        "jvp": "analytic tangent stencil (exact, matches the reference's Enzyme forward mode)",
        "fuse": fuse,
    }


# ---------------------------------------------------------------------------------------------
# synthetic inputs (deterministic; the same arrays go to the GPU path and to the CPU oracle)
# ---------------------------------------------------------------------------------------------
def local_inputs(cfg, rank, world):
    """(u0 local ndarray, problem scalars) of this rank's slab / segment of the global weak-scaled grid."""
    k = cfg["kind"]
    nx, ny = cfg["nx"], cfg["ny"]
    if k in ("bratu2d", "heat2d"):
        gny, gy0 = ny * world, ny * rank
        dx, dy = 1.0 / (nx + 1), 1.0 / (gny + 1)
        x = dx * np.arange(1, nx + 1)
        y = dy * np.arange(gy0 + 1, gy0 + ny + 1)
        if k == "bratu2d":
            u0 = np.sin(np.pi * y)[:, None] * np.sin(np.pi * x)[None, :]
            return u0, dict(dx=dx, dy=dy, gny=gny, gy0=gy0)
        X, Y = x[None, :], y[:, None]
        u0 = 16.0 * X * (1 - X) * Y * (1 - Y)
        dt = cfg["dt_scale"] * dx**2 * dy**2 / (2.0 * cfg["a"] * (dx**2 + dy**2))
        return u0, dict(dx=dx, dy=dy, gny=gny, gy0=gy0, dt=dt)
    if k == "heat1d":
        if world != 1:
            raise SystemExit("bench.py: config c2 (1-D heat with its two boundary points) is a one-GPU config")
        dx = 1.0 / (nx - 1)
        xs = dx * np.arange(nx)
        return 4.0 * xs * (1.0 - xs), dict(dx=dx, dt=cfg["dt"])
    ne = nx // 4
    gne = ne * world
    h = 1.0 / gne
    e = np.arange(rank * ne, (rank + 1) * ne)
    x = (e[:, None] * h + (LGL[None, :] + 1.0) * h / 2.0).reshape(-1)
    return np.sin(np.pi * x), dict(dx=h, dt=cfg["dt"])


class Workload:
    """Device state of one config on this rank + the Newton step the bench times."""

    def __init__(self, nk, ctx, name, cfg, rank, world, fuse):
        self.nk, self.ctx, self.name, self.cfg, self.fuse = nk, ctx, name, cfg, fuse
        self.u0, self.sc = local_inputs(cfg, rank, world)
        sc, k = self.sc, cfg["kind"]
        self.n = self.u0.size
        self.u = nk.DeviceVector.from_numpy(self.u0, ctx)
        self.res, self.rhs = self.u.similar(), self.u.similar()
        self.coef = self.u.similar() if k == "bratu2d" else None
        self.timedep = k != "bratu2d"
        if k == "bratu2d":
            self.F_, self.p = nk.bratu2d_, (sc["dx"], sc["dy"], LAMBDA, sc["gny"], sc["gy0"])
        else:
            self.un = nk.DeviceVector.from_numpy(self.u0, ctx)
            du = None  # scratch of the reference's f!(du, u, p, t); the fused kernels do not need it
            if k == "heat2d":
                self.F_ = nk.ImplicitResidual(nk.G_Euler_, nk.diffusion_)
                pin = (cfg["a"], sc["dx"], sc["dy"], nk.bc_zero_, sc["gny"], sc["gy0"])
            elif k == "heat1d":
                self.F_ = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_)
                pin = (cfg["a"], sc["dx"], nk.bc_zero_)
            else:
                self.F_ = nk.ImplicitResidual(nk.G_Euler_, nk.heat_1D_DG_)
                pin = (sc["dx"],)
            self.p = (self.un, sc["dt"], du, pin, 0.0)
        self.prob = self.F_.problem(self.u, self.p, coef=self.coef)
        self.ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(self.res), memory=MEMORY)
        self.J = nk.JacobianOperator(self.F_, self.res, self.u, self.p, coef=self.coef)
        self.kw = dict(restart=True, itmax=ITMAX, rtol=1e-30, atol=0.0, fuse=fuse,
                       reorthogonalization=bool(cfg["reorth"]))
        self.its_done = 0
        self._nrm = C.c_double()

    def residual(self):
        nk, c = self.nk, self.ctx
        nk._lib.check(c.lib.ak_residual(c.h, C.byref(self.prob), C.c_void_p(self.u.ptr), C.c_void_p(self.res.ptr),
                                        C.byref(self._nrm)))
        return self._nrm.value

    def reset_state(self):
        self.u.set(self.u0)
        return self.residual()

    def newton_step(self):
        nk, n = self.nk, self.n
        if self.timedep:                                    # u = copy(u_n); F!(res, u, p)  (implicit.jl:58, Ariadne.jl:302)
            nk.kcopy_(n, self.u, self.un)
            self.residual()
        nk.kcopy_(n, self.rhs, self.res)                    # copy(res)            src/Ariadne.jl:338
        nk.krylov_solve_(self.ws, self.J, self.rhs, **self.kw)  # krylov_solve!    :338
        self.its_done += self.ws.stats.niter
        nk.kaxpy_(n, -1.0, self.ws.x, self.u)               # u .-= s .* d         :344
        return self.residual()                              # F!(res,u,p); norm    :349-350

    def step_bytes(self):
        """Algorithmic bytes (unique reads + writes, 8-byte reals) one step moves at this fusion level."""
        n, cfg = self.n, self.cfg
        if self.sweep_active():  # + the opening tangent J S_0 of every cycle
            return 8.0 * n * (self.sweep_units() + (ITMAX // MEMORY) * cfg["jvp_bytes"] / 8.0 + self.boundary_units())
        blk = {"pair": 2, "block4": 4, "block8": 8, "sweep": 8}.get(self.fuse, 0)
        sweeps = 2 if cfg["reorth"] else 1
        units = 0.0
        ncyc = ITMAX // MEMORY
        for _ in range(ncyc):
            for k in range(1, MEMORY + 1):
                units += cfg["jvp_bytes"] / 8.0
                if blk:
                    P = -(-k // blk)
                    blen = [min(blk, k - blk * j) for j in range(P)]
                    seq = [blen[r % P] for r in range(sweeps * P)]
                    # first pass: block 0 (+ w, unless the 2-D tangent kernel does the projections while w is in registers)
                    units += seq[0] + (0 if cfg["kind"] in ("bratu2d", "heat2d") else 1)
                    for r in range(1, len(seq)):
                        units += 2 + seq[r - 1] + seq[r]                  # subtract one block, project on the next
                    units += 2 + seq[-1]                                  # final pass
                elif self.fuse == "none":
                    units += sweeps * k * (2 + 3) + 1 + 2                # dot 16n, axpy 24n per step; nrm2; divcopy
                else:
                    units += sweeps * k * 4 - (1 if self.fuse == "full" else 0) + 1 + 2
        return 8.0 * n * (units + self.boundary_units())

    def boundary_units(self):
        """8n-byte units of a step outside the Arnoldi iterations: solution updates, restart residual, Newton bookkeeping."""
        cfg = self.cfg
        ncyc = ITMAX // MEMORY
        units = ncyc * (MEMORY + 2)                                       # x (+)= V y
        units += (ncyc - 1) * (cfg["jvp_bytes"] / 8.0 + 1)                # restart residual b - J x (+ b)
        units += 2 + 1                                                    # w = copy(b), ||w||^2
        units += 2 + 3 + cfg["res_bytes"] / 8.0                           # copy(res), u -= d, F(u)
        if self.timedep:
            units += 2 + cfg["res_bytes"] / 8.0                           # u = copy(u_n), F(u)
        return units

    def sweep_active(self):
        return self.fuse == "sweep" and MEMORY <= SWEEP_KMAX  # every config of the bench has a sweep kernel

    def sweep_units(self):
        """8n-byte units of the sweep kernels of one step (csrc/sweep.cu): iteration k reads S_0..S_{k-1} and W, writes S_k
        and, unless it is the last of the cycle, reads lambda e^u and writes the next W; re-orthogonalisation adds a sweep
        without the tangent (k + 2).  (The opening tangent of a cycle is the plain tangent kernel: boundary_units.)"""
        cfg = self.cfg
        tang = cfg["jvp_bytes"] / 8.0 - 1.0   # units the tangent adds to a sweep: (lambda e^u) + W out
        units = 0.0
        for _ in range(ITMAX // MEMORY):
            for k in range(1, MEMORY + 1):
                if cfg["reorth"]:
                    units += k + 2
                units += k + 2 + (tang if k < MEMORY else 0)
        return units

    def sweep_launches(self):
        return (ITMAX // MEMORY) * MEMORY * (2 if self.cfg["reorth"] else 1)


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        try:
            if self.proc:
                self.proc.terminate()
        except Exception:
            pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# CPU legs (the oracle port of the reference algorithm on the host cores)
# ---------------------------------------------------------------------------------------------
def oracle_setup(name, cfg):
    """Single-domain oracle problem of one GPU's share of the workload: (problem, u0, r0 = F(u0))."""
    import oracle as O
    from newtonkrylov_jl_b200 import _abi as A

    O.build()
    try:  # an affinity mask set for the GPU legs must not shrink the CPU team
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except Exception:  # noqa: BLE001
        pass
    cores = O.use_all_cores()
    u0, sc = local_inputs(cfg, 0, 1)
    k = cfg["kind"]
    if k == "bratu2d":
        po = O.make_problem(A.AK_BRATU2D, cfg["nx"], cfg["ny"], dx=sc["dx"], dy=sc["dy"], lam=LAMBDA)
    elif k == "heat2d":
        po = O.make_problem(A.AK_HEAT2D, cfg["nx"], cfg["ny"], scheme=A.AK_EULER, dx=sc["dx"], dy=sc["dy"], a=cfg["a"],
                            dt=sc["dt"], un=u0)
    elif k == "heat1d":
        po = O.make_problem(A.AK_HEAT1D, cfg["nx"], 1, scheme=A.AK_EULER, dx=sc["dx"], a=cfg["a"], dt=sc["dt"], un=u0)
    else:
        po = O.make_problem(A.AK_HEAT1D_DG, cfg["nx"], 1, bc=A.AK_BC_PERIODIC, scheme=A.AK_EULER, dx=sc["dx"],
                            dt=sc["dt"], un=u0)
    r0, _ = O.residual(po, u0)
    return O, po, u0, r0, cores


def oracle_cycle(O, po, u0, r0, cfg, its):
    t = time.perf_counter()
    _, st, _ = O.krylov_solve(po, u0, r0, memory=MEMORY, restart=True, itmax=its, rtol=1e-30, atol=0.0,
                              reorthogonalization=int(bool(cfg["reorth"])))
    return time.perf_counter() - t, st["niter"]


def cpu_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on ALL host cores; rank 0 only."""
    if rank != 0:
        return
    name, cfg = args.config, CONFIGS[args.config]
    O, po, u0, r0, cores = oracle_setup(name, cfg)
    for _ in range(max(args.warmup, 1) - 1 if args.steps > 1 else 0):
        oracle_cycle(O, po, u0, r0, cfg, MEMORY)
    t_tot, it_tot = 0.0, 0
    for _ in range(args.steps):
        t, it = oracle_cycle(O, po, u0, r0, cfg, MEMORY)
        t_tot += t
        it_tot += it
    val = it_tot / t_tot
    sample = (f"{args.steps} x one GMRES({MEMORY}) restart cycle ({MEMORY} iterations) of the first Newton step of the same "
              f"{cfg['nx']}x{cfg['ny']} solve; oracle/nk_oracle.c (C + OpenMP port of the reference algorithm; Julia not "
              f"installed), OpenMP team set explicitly to {cores} threads")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": cfg["unit"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(name, cfg, 1, "n/a (CPU)"),
        "cpu_baseline": {"value": val, "unit": cfg["unit"], "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def cpu_baseline_leg(name, cfg):
    """Bounded sample for the `cpu_baseline` object of our own line: one full GMRES(20) restart cycle."""
    O, po, u0, r0, cores = oracle_setup(name, cfg)
    dtc, its = oracle_cycle(O, po, u0, r0, cfg, MEMORY)
    return {"value": its / dtc, "unit": cfg["unit"], "cores": cores, "kind": "port",
            "sample": f"one GMRES({MEMORY}) restart cycle ({its} iterations, {dtc:.2f} s) of the first Newton step of the "
                      f"same {cfg['nx']}x{cfg['ny']} solve; oracle/nk_oracle.c with OpenMP on all host cores",
            "measured_seconds": dtc, "measured_iterations": its}


# ---------------------------------------------------------------------------------------------
PROF_NAMES = ["mgs_axpy_dot", "mgs_axpy_norm", "mgs_axpy", "dot", "sumsq", "jvp", "residual", "elementwise",
              "basis_combine", "scalar", "mgs_pair", "mgs_pair_edge", "mgs_block_final", "sweep"]
NPROF = len(PROF_NAMES)


def timed_run(W, steps, warmup, barrier, sampler=None):
    """Device-resident leg: `warmup` untimed + `steps` timed Newton steps; CUDA events on the library's stream."""
    ctx = W.ctx
    W.reset_state()
    for _ in range(warmup):
        W.newton_step()
    W.reset_state()
    W.its_done = 0
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ctx.profile(True)
    ctx.launch_count(reset=True)
    ctx.timer_start()
    n_res = 0.0
    for _ in range(steps):
        n_res = W.newton_step()
    ms = ctx.timer_stop()
    barrier()
    launches = ctx.launch_count()
    prof = {c: ctx.profile_read(c) for c in range(NPROF)}
    ctx.profile(False)
    if sampler:
        sampler.stop()
    return ms, launches, prof, n_res


def kernel_table(W, prof, ms, peak):
    """Per-kernel-class launches / time / share, and GB/s for the classes whose bytes per launch are fixed."""
    n = W.n
    fixed = {"jvp": W.cfg["jvp_bytes"] * n, "residual": W.cfg["res_bytes"] * n}
    out = {}
    for c in range(NPROF):
        cnt, kms = prof[c]
        if not cnt:
            continue
        e = {"launches": cnt, "ms": round(kms, 3), "share_of_step_time": round(kms / ms, 4)}
        b = fixed.get(PROF_NAMES[c])  # (restart residuals b - J x read one more vector: counted at the plain JVP's bytes)
        if PROF_NAMES[c] == "jvp" and W.cfg["kind"] in ("bratu2d", "heat2d") and W.fuse in ("pair", "block4", "block8", "sweep"):
            b = None  # the 2-D tangent launches also carry the first projection pass: bytes per launch vary with k
        if b:
            e["GBs"] = round(b / (kms / cnt * 1e-3) / 1e9, 1)
            e["frac_of_peak"] = round(e["GBs"] / peak, 4)
        out[PROF_NAMES[c]] = e
    return out


def c1_small_regime(nk, ctx):
    """BASELINE config 1 (examples/bratu.jl: 1-D Bratu, N = 10 000, the CPU-runnable case): CG iterations/s of one linear
    solve of the first Newton step, GPU (one persistent block: the whole of cg! without a launch inside) against the CPU
    oracle on the host cores.  80 KB per vector: a launch-latency regime, not a bandwidth one."""
    import oracle as O
    from newtonkrylov_jl_b200 import _abi as A

    N, lam, its = 10_000, LAMBDA, 2000
    dx = 1.0 / (N + 1)
    x = np.linspace(dx, 1.0 - dx, N)
    u0 = np.sin(np.pi * x)
    u = nk.DeviceVector.from_numpy(u0, ctx)
    res, coef = u.similar(), u.similar()
    prob = nk.bratu_.problem(u, (dx, lam), coef=coef)
    nk._lib.check(ctx.lib.ak_residual(ctx.h, C.byref(prob), C.c_void_p(u.ptr), C.c_void_p(res.ptr), None))
    J = nk.JacobianOperator(nk.bratu_, res, u, (dx, lam), coef=coef)
    ws = nk.krylov_workspace("cg", nk.KrylovConstructor(res))
    b = res.copy()
    best = None
    for _ in range(4):
        ctx.sync()
        t0 = time.perf_counter()
        nk.krylov_solve_(ws, J, b, rtol=1e-30, atol=0.0, itmax=its)
        ctx.sync()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    gpu_its = ws.stats.niter
    O.build()
    cores = O.use_all_cores()
    po = O.make_problem(A.AK_BRATU1D, N, 1, dx=dx, lam=lam)
    r0, _ = O.residual(po, u0)
    t0 = time.perf_counter()
    _, stc, _ = O.krylov_solve(po, u0, r0, algo=A.AK_ALGO_CG, rtol=1e-30, atol=0.0, itmax=its)
    dtc = time.perf_counter() - t0
    return {"workload": f"1D Bratu N={N} (examples/bratu.jl:40-46), algo=:cg, {its} iterations of one linear solve",
            "gpu_cg_iters_per_sec": gpu_its / best, "gpu_iterations": gpu_its, "gpu_seconds": best,
            "gpu_path": "k_cg_small_bratu1d: one persistent 1024-thread block, p and r in shared memory, no launch inside the solve",
            "cpu_cg_iters_per_sec": stc["niter"] / dtc, "cpu_cores": cores, "cpu_kind": "port (oracle/nk_oracle.c)",
            "wall_clock": "host perf_counter around the blocking solve call (one launch + one sync)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS), help="BASELINE.json config (default: the headline c4)")
    ap.add_argument("--nx", type=int, default=None, help="override the grid width (testing)")
    ap.add_argument("--ny", type=int, default=None, help="override rows per GPU (testing)")
    ap.add_argument("--fuse", default="sweep", choices=sorted(FUSE_CODE))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL all-reduce / send-recv instead of peer memory")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    cfgs = {k: dict(v) for k, v in CONFIGS.items()}
    if args.nx:
        cfgs[args.config]["nx"] = args.nx
    if args.ny and cfgs[args.config]["ny"] > 1:
        cfgs[args.config]["ny"] = args.ny
    CONFIGS.update(cfgs)
    name, cfg = args.config, CONFIGS[args.config]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        cpu_reference_arm(args, rank, world)
        return

    import torch
    import newtonkrylov_jl_b200 as nk
    from newtonkrylov_jl_b200 import _abi as A

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # one process per GPU: run (and allocate the pinned staging buffer of the e2e leg) on the CPUs next to this
        # GPU, so that eight simultaneous 512 MiB host<->device copies do not all go through one socket's memory
        try:
            import pynvml

            pynvml.nvmlInit()
            try:  # NVML numbers physical devices; go through the UUID in case CUDA_VISIBLE_DEVICES remaps them
                uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:  # noqa: BLE001
                handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            pynvml.nvmlDeviceSetCpuAffinity(handle)
        except Exception as e:  # noqa: BLE001 - affinity is an optimisation only
            print(f"[bench] rank {rank}: CPU affinity not set ({e})", file=sys.stderr)
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # for waits that must not spin on a host core (the CPU-baseline leg)
    ctx = nk.get_context(local_rank)
    if world > 1:
        ids = [nk.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_comm(world, rank, ids[0])
        if not args.no_p2p:
            # NVLink peer memory: fused reductions + ghost-row push; if any rank cannot map its
            # peers (no IPC in this container) every rank falls back to the NCCL path together
            try:
                ctx.enable_p2p(max(c["nx"] for c in CONFIGS.values() if c["ny"] > 1))
                ok = 1
            except Exception as e:  # noqa: BLE001
                print(f"[bench] rank {rank}: peer memory unavailable ({e}); using the NCCL path", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag[0]) == 0 and ctx.p2p:
                ctx.use_p2p(False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def reduce_max_sum(ms, launches):
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax, tsum = t.clone(), t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return float(tmax[0]), int(tsum[1])
        return ms, launches

    peak, peak_src = measured_peak_gbs()
    lib, h = ctx.lib, ctx.h

    # ---- device-resident leg (value) ---------------------------------------------------------
    W = Workload(nk, ctx, name, cfg, rank, world, args.fuse)
    n = W.n
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches, prof, n_res = timed_run(W, args.steps, args.warmup, barrier, sampler)
    ms_max, launches_total = reduce_max_sum(ms, launches)
    iters = W.its_done                       # identical on every rank (global iterations)
    value = iters * world / (ms_max * 1e-3)  # slab-iterations per second

    # ---- roofline of the dominant kernel (rank 0) ------------------------------------------------
    share = kernel_table(W, prof, ms, peak)
    # dominant kernel of the timed region and its algorithmic bytes per launch
    if args.fuse == "pair":
        dom, dom_bytes = 10, 48 * n
        dom_name = "k_mgs_block<2,2> (w -= h_a v_a + h_b v_b ; <y_a,w>, <y_b,w>, <y_b,y_a>): two Gram-Schmidt steps per pass"
    elif args.fuse == "block4":
        dom, dom_bytes = 10, 80 * n
        dom_name = ("k_mgs_block<4,4> (w -= sum_b h_b v_b ; 4 projections <y_b,w> + 6 Gram entries <y_b,y_a>): "
                    "four Gram-Schmidt steps per pass")
    elif args.fuse == "block8":
        dom, dom_bytes = 10, 144 * n
        dom_name = "k_mgs_block<8,8> (w -= sum_b c_b S_b ; 8 projections <S'_b,w>): eight Gram-Schmidt steps per pass"
    elif args.fuse == "sweep" and W.sweep_active():
        # every launch of the template handles a different basis size k: bytes per launch = the mean over a step
        dom, dom_bytes = 13, 8.0 * n * W.sweep_units() / W.sweep_launches()
        dom_name = ("k_sweep<KB, stencil> (z = W/rho - sum_j c_j S_j ; ||z||^2, <S_j,z> ; y = J z ; <S_j,y>, <z,y>): one pass over the "
                    "basis per GMRES iteration, 8n(k+4) bytes at basis size k; achieved = mean bytes per launch / mean "
                    "launch time over the %d launches of a step (k = 1..%d)" % (W.sweep_launches(), MEMORY))
    elif args.fuse == "sweep":
        dom, dom_bytes = 10, 144 * n
        dom_name = "k_mgs_block<8,8> (fuse = sweep applies to the 2-D problems; this config runs the eight-step blocked passes)"
    elif args.fuse == "none":
        dom, dom_bytes = 2, 24 * n
        dom_name = "k_mgs_step<AXPY> (w -= h_i v_i)"
    else:
        dom, dom_bytes = 0, 32 * n
        dom_name = "k_mgs_step<AXPY,DOT> (w -= h_i v_i ; h_{i+1} = <v_{i+1}, w>)"
    cnt, kms = prof[dom]
    traffic, traffic_src = None, None
    try:  # committed constant: dram bytes per launch of this kernel from the ncu --set full capture under profiles/
        kt = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json"))).get(args.fuse)
        if kt and kt["n"] == n and not (args.fuse == "sweep" and not W.sweep_active()):
            traffic = kt["dram_bytes_per_launch"]
            traffic_src = "committed constant from profiles/kernel_traffic.json (ncu --set full capture), not measured in this run"
    except Exception:
        pass
    step_bytes = W.step_bytes()
    step_gbs = step_bytes * args.steps / (ms * 1e-3) / 1e9
    gs_ms = sum(prof[c][1] for c in (0, 1, 2, 10, 11, 12))
    family = {
        "one_sweep_iterations (k_sweep: Gram-Schmidt update + norms + tangent + projections)": round(prof[13][1] / ms, 4),
        "gram_schmidt_passes (k_mgs_block / k_mgs_step, all instantiations)": round(gs_ms / ms, 4),
        "jvp (k_stencil*/k_dg tangent; 2-D: with the first projection pass folded in)": round(prof[5][1] / ms, 4),
        "cycle_boundary (basis_combine + element-wise + norms)": round((prof[8][1] + prof[7][1] + prof[4][1] + prof[3][1]) / ms, 4),
        "residual": round(prof[6][1] / ms, 4),
        "scalar (Givens / back-substitution, one block)": round(prof[9][1] / ms, 4),
    }
    roofline = None
    if cnt:
        achieved = dom_bytes / (kms / cnt * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom_name,
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                    "avg_launch_ms": kms / cnt, "launches_timed": cnt, "traffic": traffic, "traffic_source": traffic_src,
                    "dominant_kernel_share_of_step": round(kms / ms, 4),
                    "bytes_moved_per_step": step_bytes, "step_GBs": step_gbs, "frac_step": step_gbs / peak,
                    "frac_step_note": "algorithmic bytes of ALL kernels of a step (DESIGN.md byte model of this fusion "
                                      "level) / step time / measured copy peak",
                    "kernel_family_share_of_step": family, "kernel_share_of_step": share}
    # per-iteration view against the reference op list  B(k) = 8n(5k+6)  (10k with reorthogonalisation)
    kfac = 10 if cfg["reorth"] else 5
    ref_bytes = sum(8.0 * n * (kfac * k + 6) for k in range(1, MEMORY + 1)) * (ITMAX // MEMORY) * args.steps
    per_iter = {"reference_op_list_bytes": ref_bytes, "achieved_GBs_vs_reference_op_list": ref_bytes / (ms * 1e-3) / 1e9,
                "frac_of_peak": ref_bytes / (ms * 1e-3) / 1e9 / peak,
                "note": "fused kernels move fewer bytes than the reference op list, so this may exceed 1"}

    # ---- end-to-end leg: host buffers through ak_newton_solve_host -------------------------------
    e2e = None
    if not args.no_e2e:
        hp, hun = C.c_void_p(), C.c_void_p()
        nk._lib.check(lib.ak_host_alloc(n, C.byref(hp)))
        ubuf = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=(n,))
        if W.timedep:
            nk._lib.check(lib.ak_host_alloc(n, C.byref(hun)))
            np.ctypeslib.as_array(C.cast(hun, C.POINTER(C.c_double)), shape=(n,))[:] = W.u0.reshape(-1)
        o = A.default_newton_opts(max_niter=0)  # `outer <= max_niter` admits exactly one Newton step
        o.krylov = A.default_krylov_opts(restart=1, itmax=ITMAX, rtol=1e-30, atol=0.0, fuse=FUSE_CODE[args.fuse],
                                         reorthogonalization=int(bool(cfg["reorth"])))
        o.krylov_rtol_override = 1
        st = A.ak_newton_stats()
        prob_h = W.F_.problem(W.u, W.p)  # coef / u_n device copies are made by the entry point
        e_steps = min(args.steps, 5)
        e_its = 0

        def e2e_step():
            nonlocal e_its
            nk._lib.check(lib.ak_newton_solve_host(h, C.byref(prob_h), hp, hun if W.timedep else None, C.byref(o),
                                                   C.byref(st), None, None, 0))
            e_its += int(st.inner_iterations)

        ubuf[:] = W.u0.reshape(-1)
        e2e_step()  # warm-up (first-touch of the allocator)
        ubuf[:] = W.u0.reshape(-1)
        e_its = 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            if W.timedep:
                ubuf[:] = W.u0.reshape(-1)  # every time step starts from u = u_n (host copy, inside the timed region)
            e2e_step()
        barrier()
        dt_e = time.perf_counter() - t0
        te = torch.tensor([dt_e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = 8 * n * world * (2 if W.timedep else 1)
        e2e = {"value": e_its * world / float(te[0]), "unit": cfg["unit"], "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 8 * n * world, "steps": e_steps, "ms_per_step": 1e3 * float(te[0]) / e_steps,
               "api": "ak_newton_solve_host (pinned host u in/out, workspace allocated per call like the reference)",
               # time the host<->device copies (and the per-call workspace set-up) add to a step, and the copy rate it implies
               "copy_ms_per_step": 1e3 * float(te[0]) / e_steps - ms_max / args.steps,
               "host_copy_GBs_per_rank": (h2d + 8 * n * world) / world / 1e9 /
                                         max(float(te[0]) / e_steps - ms_max / args.steps * 1e-3, 1e-9),
               "host_copy_GBs_all_ranks": (h2d + 8 * n * world) / 1e9 /
                                          max(float(te[0]) / e_steps - ms_max / args.steps * 1e-3, 1e-9),
               "final_n_res": st.n_res}
        lib.ak_host_free(hp)
        if W.timedep:
            lib.ak_host_free(hun)

    # ---- the other BASELINE configs (default c4 line): it/s and per-kernel GB/s ---------------------
    others = None
    if name == "c4" and not args.no_other_configs and not args.nx and not args.ny:
        del W
        others = {}
        for oname in (("c2", "c3", "c5") if world == 1 else ("c5",)):
            ocfg = CONFIGS[oname]
            Wo = Workload(nk, ctx, oname, ocfg, rank, world, args.fuse)
            oms, ol, oprof, onres = timed_run(Wo, min(args.steps, 5), 3, barrier)
            oms_max, _ = reduce_max_sum(oms, ol)
            osteps = min(args.steps, 5)
            ob = Wo.step_bytes()
            others[oname] = {
                "workload": workload_config(oname, ocfg, world, args.fuse)["workload"], "n_gpus": world,
                "value": Wo.its_done * world / (oms_max * 1e-3), "unit": ocfg["unit"], "steps": osteps,
                "ms_per_step": oms_max / osteps, "final_n_res": onres,
                "bytes_moved_per_step": ob, "frac_step": ob * osteps / (oms * 1e-3) / 1e9 / peak,
                "kernels": kernel_table(Wo, oprof, oms, peak),
            }
            del Wo
        if world == 1:
            others["c1"] = c1_small_regime(nk, ctx)

    # ---- CPU baseline (rank 0): one restart cycle on all host cores, at every N -------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(name, cfg)
    if world > 1:
        dist.barrier(group=cpu_group)  # the other ranks block in a socket wait meanwhile: they do not take cores from the oracle

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(name, cfg, world, args.fuse),
            "gmres_iters_per_sec_global": iters / (ms_max * 1e-3), "gmres_iterations_timed": iters,
            "final_n_res": n_res, "fuse": args.fuse, "peer_memory_path": bool(ctx.p2p),
            "clocks": sampler.summary() if sampler else None,
            "e2e": e2e, "gpu_launches": launches_total, "roofline": roofline, "per_iteration": per_iter,
            "cpu_baseline": cpu, "other_configs": others,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
