"""Parity ledger: every GPU-vs-oracle comparison records what it OBSERVED (not only that it passed).

The `-m gpu` tests call `record(section, key, **fields)`; the entries are merged into
`profiles/parity_ledger.json` (committed after a GPU session) and, when the run happens under gpurun,
into `gpurun_out/parity_ledger.json` (the copy that travels back from the GPU box).

Fields used by the Newton-level cases (tests/test_gpu_solvers.py::assert_newton_parity):
    max_rel_nres_dev      max_k |n_res_gpu[k] - n_res_oracle[k]| / n_res_oracle[k]
    final_u_rel_dev       ||u_gpu - u_oracle|| / ||u_oracle||
    inner_count_diffs     GMRES/CG iterations per Newton step, GPU minus oracle
    oracle_ulp_sensitivity  the same deviation between two ORACLE runs whose u0 differ by one ulp in one entry
    met_1e-10_1e-8        True when the north_star bar (1e-10 on ||F||, 1e-8 on u, equal counts) held WITHOUT the
                          sensitivity-based relaxation
"""
import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATHS = [os.path.join(ROOT, "profiles", "parity_ledger.json"), os.path.join(ROOT, "gpurun_out", "parity_ledger.json")]
_session = {}


def _jsonable(v):
    try:
        import numpy as np

        if isinstance(v, (np.floating, np.integer)):
            return v.item()
        if isinstance(v, np.bool_):
            return bool(v)
        if isinstance(v, np.ndarray):
            return [_jsonable(x) for x in v.tolist()]
    except Exception:  # noqa: BLE001
        pass
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    if isinstance(v, dict):
        return {str(k): _jsonable(x) for k, x in v.items()}
    return v


def record(section, key, **fields):
    """Store one observation and rewrite the ledger files (cheap: a few hundred small entries)."""
    _session.setdefault(section, {})[key] = _jsonable(fields)
    for path in _PATHS:
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            data = {}
            if os.path.exists(path):
                try:
                    data = json.load(open(path))
                except Exception:  # noqa: BLE001
                    data = {}
            for sec, entries in _session.items():
                data.setdefault(sec, {}).update(entries)
            data["_meta"] = {
                "what": "observed GPU-vs-oracle deviations per parity case (tests/ledger.py)",
                "bar": "north_star: same Newton count, per-iteration ||F|| within 1e-10 relative, final u within 1e-8 relative",
                "updated": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                "gpu": _gpu_name(),
            }
            with open(path, "w") as f:
                json.dump(data, f, indent=1, sort_keys=True)
        except OSError:
            pass


_gpu = None


def _gpu_name():
    global _gpu
    if _gpu is None:
        try:
            import torch

            _gpu = torch.cuda.get_device_name(0) if torch.cuda.is_available() else "none"
        except Exception:  # noqa: BLE001
            _gpu = "unknown"
    return _gpu
