"""bench.py contract pieces that run without a GPU: the reference arm (`--impl reference`, the CPU oracle on the host
cores) prints ONE JSON line with the keys the driver reads; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e)


def test_reference_arm_prints_one_json_line():
    r = run(["--impl", "reference", "--nx", "192", "--ny", "160", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gmres_iters_per_sec" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["config"]["grid_per_gpu"] == [192, 160] and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    r = run(["--impl", "reference", "--nx", "64", "--ny", "64", "--steps", "1", "--warmup", "1", "--gpus", "2"],
            env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    r = run(["--steps", "1", "--warmup", "1", "--nx", "64", "--ny", "64"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_reference_arm_uses_all_host_cores_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1; the CPU column must still be timed on the host's cores and say how many,
    on the same sample (one full GMRES(20) restart cycle) at every N (SCALE vs_reference was void)."""
    env = {"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "4", "OMP_NUM_THREADS": "1"}
    r = run(["--impl", "reference", "--nx", "128", "--ny", "96", "--steps", "1", "--warmup", "1", "--gpus", "4"], env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count()
    assert d["cpu_baseline"]["cores"] == ncores and d["n_gpus"] == 4
    assert "20 iterations" in d["cpu_baseline"]["sample"]
    assert d["config"]["grid_per_gpu"] == [128, 96]  # the per-GPU workload, whatever N is


def test_reference_arm_covers_every_baseline_config():
    for cfg, key in (("c2", "1D implicit-Euler heat"), ("c3", "reorthogonalization=true"), ("c5", "DG 1D heat")):
        r = run(["--impl", "reference", "--config", cfg, "--nx", "1024", "--ny", "48", "--steps", "1", "--warmup", "1"])
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads(r.stdout.strip().splitlines()[-1])
        assert key in d["config"]["workload"] and d["config"]["baseline_config"] == cfg and d["value"] > 0
