#!/usr/bin/env python
"""bench.py — GMRES iterations/s of the JFNK inner loop on 2-D Bratu 8192^2 fp64 per GPU.

Contract (driver):  python bench.py --gpus N --steps K --warmup W   (N > 1 under torchrun)
prints ONE JSON line from rank 0.

Workload (BASELINE.json metric, SURVEY.md §8d config C4, protocol A): 2-D Bratu, lambda = 3.5,
u0 = sin(pi x) sin(pi y) on the global unit square; 8192 x 8192 unknowns PER GPU (weak scaling,
slab decomposition along y, one halo row per neighbour per stencil application, NCCL all-reduce
for every Arnoldi inner product).  One STEP = one Newton step of `newton_krylov!`
(src/Ariadne.jl:321-368) with `krylov_kwargs = (; restart = true, itmax = 40, rtol = 1e-30, atol = 0)`
and memory = 20: copy(res) -> GMRES(20) x 2 restart cycles (40 iterations, each = 1 JVP + k fused
modified-Gram-Schmidt steps + Givens update) -> u .-= d -> F!(res, u) + norm(res).
The tolerance is set so that every step does exactly 40 iterations (fixed work per step).

value  = GMRES iterations/s per 8192^2 slab, summed over the slabs (= GPUs): 40*K*N / t, inputs
         resident in HBM, timed with CUDA events on the library's stream, max over ranks.
e2e    = the same metric through the host-buffer entry point ak_newton_solve_host (what a Julia
         caller holding an Array{Float64} calls): every step copies u host->device from pinned
         memory, allocates the Krylov workspace like the reference does per newton_krylov! call,
         runs the same Newton step and copies u back.
roofline = the dominant kernel (full pass of the blocked modified-Gram-Schmidt sweep: 144n bytes per launch = 18n per
         Gram-Schmidt step with --fuse block8; 80n / 48n with block4 / pair; with --fuse mgs/full the axpy_i + dot_{i+1}
         kernel, 32n) timed live with CUDA events
         inside the timed region (library profiler); traffic from the committed ncu --set full capture.
cpu_baseline = the CPU oracle (oracle/nk_oracle.c, a port of the reference algorithm) on the host
         cores, one GMRES(20) restart cycle of the same solve.

--impl reference times the reference's CPU path (the oracle port — Julia is not installed and the
reference's dependencies are not vendored, so oracle/_ref does not exist) on the same workload.
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

NX = NY_PER_GPU = 8192
LAMBDA = 3.5
MEMORY = 20
ITMAX = 40
METRIC = "gmres_iters_per_sec"
UNIT = "GMRES it/s per 8192^2 slab, summed over GPUs"


def workload_config(n_gpus, nx=NX, ny=NY_PER_GPU):
    return {
        "workload": f"2D Bratu {nx}x{ny} fp64 per GPU (lambda=3.5, u0=sin(pi x)sin(pi y)), "
                    f"one Newton step/step: GMRES(restart, memory={MEMORY}, itmax={ITMAX}) + update + residual",
        "grid_per_gpu": [nx, ny],
        "global_grid": [nx, ny * n_gpus],
        "gmres_iters_per_step": ITMAX,
        "decomposition": "slab along y; Arnoldi sums and ghost rows through NVLink peer memory fused into the "
                         "Gram-Schmidt kernels (NCCL only at cycle boundaries)" if n_gpus > 1 else "single GPU",
        "l2": "inputs larger than L2 (each vector is 512 MiB, L2 is 126 MB): no flush needed",
        "jvp": "analytic tangent stencil (exact, matches the reference's Enzyme forward mode)",
    }


def initial_guess(nx, ny_local, gy0, gny):
    dx, dy = 1.0 / (nx + 1), 1.0 / (gny + 1)
    x = dx * np.arange(1, nx + 1)
    y = dy * np.arange(gy0 + 1, gy0 + ny_local + 1)
    return np.sin(np.pi * y)[:, None] * np.sin(np.pi * x)[None, :], dx, dy


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
def cpu_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on the host cores."""
    if rank != 0:
        return
    import oracle as O
    from newtonkrylov_jl_b200 import _abi as A

    O.build()
    nx, ny = args.nx, args.ny
    sample_its = MEMORY
    u0, dx, dy = initial_guess(nx, ny, 0, ny)
    po = O.make_problem(A.AK_BRATU2D, nx, ny, dx=dx, dy=dy, lam=LAMBDA)
    res, _ = O.residual(po, u0)

    def step(its):
        t = time.perf_counter()
        _, st, _ = O.krylov_solve(po, u0, res, memory=MEMORY, restart=True, itmax=its, rtol=1e-30, atol=0.0)
        return time.perf_counter() - t, st["niter"]

    # bound the sample: if one restart cycle takes > 20 s on this host, time half a cycle
    t1, it1 = step(sample_its)
    if t1 > 20.0:
        sample_its = MEMORY // 2
    for _ in range(max(args.warmup - 1, 0)):
        step(sample_its)
    t_tot, it_tot = 0.0, 0
    for _ in range(args.steps):
        t, it = step(sample_its)
        t_tot += t
        it_tot += it
    val = it_tot / t_tot
    sample = (f"{args.steps} x the first {sample_its} GMRES iterations (one restart cycle, memory={MEMORY}) of the "
              f"same {nx}x{ny} solve; oracle/nk_oracle.c (C + OpenMP port of the reference algorithm; Julia not installed)")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(1, nx, ny),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": O.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=NX)
    ap.add_argument("--ny", type=int, default=NY_PER_GPU, help="rows per GPU")
    ap.add_argument("--fuse", default="block8", choices=["none", "mgs", "full", "pair", "block4", "block8"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL all-reduce / send-recv instead of peer memory")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

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
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = nk.get_context(local_rank)
    if world > 1:
        ids = [nk.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_comm(world, rank, ids[0])
        if not args.no_p2p:
            # NVLink peer memory: fused reductions + ghost-row push; if any rank cannot map its
            # peers (no IPC in this container) every rank falls back to the NCCL path together
            try:
                ctx.enable_p2p(args.nx)
                ok = 1
            except Exception as e:  # noqa: BLE001
                print(f"[bench] rank {rank}: peer memory unavailable ({e}); using the NCCL path", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag[0]) == 0 and ctx.p2p:
                ctx.use_p2p(False)

    nx, ny = args.nx, args.ny
    n = nx * ny
    gny, gy0 = ny * world, ny * rank
    u0, dx, dy = initial_guess(nx, ny, gy0, gny)
    lib, h = ctx.lib, ctx.h

    u = nk.DeviceVector.from_numpy(u0, ctx)
    res, coef, rhs = u.similar(), u.similar(), u.similar()
    prob = nk.bratu2d_.problem(u, (dx, dy, LAMBDA, gny, gy0), coef=coef)
    ws = nk.krylov_workspace("gmres", nk.KrylovConstructor(res), memory=MEMORY)
    J = nk.JacobianOperator(nk.bratu2d_, res, u, (dx, dy, LAMBDA, gny, gy0), coef=coef)
    nrm = C.c_double()
    P = lambda v: C.c_void_p(v.ptr)
    kw = dict(restart=True, itmax=ITMAX, rtol=1e-30, atol=0.0, fuse=args.fuse)

    def residual():
        nk._lib.check(lib.ak_residual(h, C.byref(prob), P(u), P(res), C.byref(nrm)))
        return nrm.value

    its_done = [0]

    def newton_step():
        nk.kcopy_(n, rhs, res)                       # copy(res)            src/Ariadne.jl:338
        nk.krylov_solve_(ws, J, rhs, **kw)           # krylov_solve!        :338
        its_done[0] += ws.stats.niter
        nk.kaxpy_(n, -1.0, ws.x, u)                  # u .-= s .* d         :344
        return residual()                            # F!(res,u,p); norm    :349-350

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def reset_state():
        u.set(u0)
        residual()

    # ---- device-resident leg (value) ---------------------------------------------------------
    reset_state()
    for _ in range(args.warmup):
        newton_step()
    reset_state()
    its_done[0] = 0
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ctx.profile(True)
    ctx.launch_count(reset=True)
    ctx.timer_start()
    for _ in range(args.steps):
        n_res = newton_step()
    ms = ctx.timer_stop()
    barrier()
    launches = ctx.launch_count()
    prof = {c: ctx.profile_read(c) for c in range(12)}
    ctx.profile(False)
    if sampler:
        sampler.stop()
    t = torch.tensor([ms, float(launches), float(its_done[0])], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_max, launches_total = float(tmax[0]), int(tsum[1])
    else:
        ms_max, launches_total = ms, launches
    iters = its_done[0]                      # identical on every rank (global iterations)
    value = iters * world / (ms_max * 1e-3)  # slab-iterations per second

    # ---- roofline of the dominant kernel (rank 0) ------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    names = ["mgs_axpy_dot", "mgs_axpy_norm", "mgs_axpy", "dot", "sumsq", "jvp", "residual", "elementwise",
             "basis_combine", "scalar", "mgs_pair", "mgs_pair_edge"]
    share = {names[c]: {"launches": prof[c][0], "ms": round(prof[c][1], 3), "share_of_step_time": round(prof[c][1] / ms, 4)}
             for c in range(12) if prof[c][0]}
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
    elif args.fuse == "none":
        dom, dom_bytes = 2, 24 * n
        dom_name = "k_mgs_step<AXPY> (w -= h_i v_i)"
    else:
        dom, dom_bytes = 0, 32 * n
        dom_name = "k_mgs_step<AXPY,DOT> (w -= h_i v_i ; h_{i+1} = <v_{i+1}, w>)"
    cnt, kms = prof[dom]
    traffic = None  # dram bytes per launch of the dominant kernel from the committed ncu --set full capture
    try:
        kt = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json"))).get(args.fuse)
        if kt and kt["n"] == n:
            traffic = kt["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = None
    if cnt:
        achieved = dom_bytes / (kms / cnt * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom_name,
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                    "avg_launch_ms": kms / cnt, "launches_timed": cnt, "traffic": traffic,
                    "kernel_share_of_step": share}
    # per-iteration view against the reference op list  B(k) = 8n(5k+6)
    ref_bytes = sum(8.0 * n * (5 * k + 6) for k in range(1, MEMORY + 1)) * (ITMAX // MEMORY) * args.steps
    per_iter = {"reference_op_list_bytes": ref_bytes, "achieved_GBs_vs_reference_op_list": ref_bytes / (ms * 1e-3) / 1e9,
                "frac_of_peak": ref_bytes / (ms * 1e-3) / 1e9 / peak,
                "note": "fused kernels move fewer bytes than the reference op list, so this may exceed 1"}

    # ---- end-to-end leg: host buffers through ak_newton_solve_host -------------------------------
    e2e = None
    if not args.no_e2e:
        hp = C.c_void_p()
        nk._lib.check(lib.ak_host_alloc(n, C.byref(hp)))
        ubuf = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=(n,))
        o = A.default_newton_opts(max_niter=0)  # `outer <= max_niter` admits exactly one Newton step
        o.krylov = A.default_krylov_opts(restart=1, itmax=ITMAX, rtol=1e-30, atol=0.0,
                                         fuse={"none": 0, "mgs": 1, "full": 2, "pair": 3, "block4": 4, "block8": 5}[args.fuse])
        o.krylov_rtol_override = 1
        st = A.ak_newton_stats()
        prob_h = nk.bratu2d_.problem(u, (dx, dy, LAMBDA, gny, gy0))  # coef is allocated by the entry point
        e_steps = min(args.steps, 5)
        e_its = 0

        def e2e_step():
            nonlocal e_its
            nk._lib.check(lib.ak_newton_solve_host(h, C.byref(prob_h), hp, None, C.byref(o), C.byref(st), None, None, 0))
            e_its += int(st.inner_iterations)

        ubuf[:] = u0.reshape(-1)
        e2e_step()  # warm-up (first-touch of the allocator)
        ubuf[:] = u0.reshape(-1)
        e_its = 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        barrier()
        dt_e = time.perf_counter() - t0
        te = torch.tensor([dt_e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": e_its * world / float(te[0]), "unit": UNIT, "h2d_bytes_per_step": 8 * n * world,
               "d2h_bytes_per_step": 8 * n * world, "steps": e_steps, "ms_per_step": 1e3 * float(te[0]) / e_steps,
               "api": "ak_newton_solve_host (pinned host u in/out, workspace allocated per call like the reference)",
               "final_n_res": st.n_res}
        lib.ak_host_free(hp)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample on the host cores -----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O

        O.build()
        po = O.make_problem(A.AK_BRATU2D, nx, ny, dx=dx, dy=dy, lam=LAMBDA)
        r0, _ = O.residual(po, u0)
        sample_its = MEMORY // 2
        t0 = time.perf_counter()
        _, stc, _ = O.krylov_solve(po, u0, r0, memory=MEMORY, restart=True, itmax=sample_its, rtol=1e-30, atol=0.0)
        dtc = time.perf_counter() - t0
        # the first `sample_its` iterations are cheaper than the average of a cycle: scale by the op-list bytes
        b_sample = sum(5 * k + 6 for k in range(1, sample_its + 1))
        b_cycle = sum(5 * k + 6 for k in range(1, MEMORY + 1))
        cyc_time = dtc * b_cycle / b_sample
        cpu = {"value": MEMORY / cyc_time, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
               "sample": f"first {sample_its} GMRES iterations of one restart cycle of the same {nx}x{ny} solve "
                         f"({dtc:.2f} s measured), scaled to a full {MEMORY}-iteration cycle by the op-list bytes "
                         f"8n(5k+6); oracle/nk_oracle.c with OpenMP on all host cores",
               "measured_seconds": dtc, "measured_iterations": stc["niter"]}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world, nx, ny),
            "gmres_iters_per_sec_global": iters / (ms_max * 1e-3), "gmres_iterations_timed": iters,
            "final_n_res": n_res, "fuse": args.fuse, "peer_memory_path": bool(ctx.p2p),
            "clocks": sampler.summary() if sampler else None,
            "e2e": e2e, "gpu_launches": launches_total, "roofline": roofline, "per_iteration": per_iter,
            "cpu_baseline": cpu,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
