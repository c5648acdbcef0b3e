"""bench.py's byte model of a step (the `roofline.bytes_moved_per_step` / `frac_step` figures) against the counts
DESIGN.md section 4 states for the one-sweep kernels: iteration k moves 8n(k + 4) bytes (k + 3 without a coefficient
vector, k + 2 for the last iteration of a cycle and for the first sweep of a re-orthogonalised iteration)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class _W:  # the byte model only needs the config, the fusion level and n
    pass


def workload(b, name, fuse):
    w = b.Workload.__new__(b.Workload)
    w.cfg, w.fuse = b.CONFIGS[name], fuse
    w.n = w.cfg["nx"] * w.cfg["ny"]
    w.timedep = w.cfg["kind"] != "bratu2d"
    return w


def test_sweep_units_match_the_stated_bytes_per_iteration():
    b = load_bench()
    m, ncyc = b.MEMORY, b.ITMAX // b.MEMORY
    # C4 (Bratu: coefficient vector): sum_{k=1}^{19} (k + 4) + (20 + 2) = 288 per cycle
    w = workload(b, "c4", "sweep")
    assert w.sweep_active()
    assert w.sweep_units() == ncyc * (sum(k + 4 for k in range(1, m)) + m + 2) == 576
    assert w.sweep_launches() == ncyc * m
    # C3 (heat, two sweeps per iteration): (k + 2) + (k + 3), last iteration (k + 2) + (k + 2)
    w3 = workload(b, "c3", "sweep")
    assert w3.sweep_units() == ncyc * (sum(2 * k + 5 for k in range(1, m)) + 2 * m + 4)
    assert w3.sweep_launches() == 2 * ncyc * m
    # C2 / C5 (no coefficient vector): sum (k + 3) + (20 + 2)
    for name in ("c2", "c5"):
        wi = workload(b, name, "sweep")
        assert wi.sweep_units() == ncyc * (sum(k + 3 for k in range(1, m)) + m + 2)


def test_step_bytes_of_the_fusion_levels_are_ordered():
    """Every further fusion level moves fewer bytes per step; the sweep moves about half of block8's."""
    b = load_bench()
    levels = ["none", "mgs", "full", "pair", "block4", "block8", "sweep"]
    by = [workload(b, "c4", f).step_bytes() for f in levels]
    assert all(x > y for x, y in zip(by, by[1:])), by
    n = 8192 * 8192
    assert by[-1] == 8.0 * n * 641 and by[-2] == 8.0 * n * 1163
    # the reference op list of gmres!: 8n(5k + 6) per iteration + the step's bookkeeping
    ref = sum(5 * k + 6 for k in range(1, b.MEMORY + 1)) * (b.ITMAX // b.MEMORY)
    assert abs(by[0] / (8.0 * n) - ref) <= 80
