"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/multi_gpu_worker.py
under torchrun, one rank per GPU.  The same logic is covered on CPU by tests/test_dist_cpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_decomposition_matches_oracle(nk, world):
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(HERE, "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    # every family of cases reported in: 1-D segments, Midpoint/Trapezoid, 2-D slabs, with NCCL and with peer memory
    for needle in ("nccl] dg: ok", "heat2d midpoint/trapezoid: ok", "nccl] heat2d_periodic: ok", "p2p] bratu2d: ok",
                   "p2p] dg: ok", "p2p] cg + reorthogonalised blocked gmres + default itmax: ok", "nccl] cg + reorthogonalised blocked gmres + default itmax: ok"):
        assert needle in r.stdout, (needle, r.stdout[-3000:])
