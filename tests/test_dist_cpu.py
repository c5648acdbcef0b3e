"""CPU tests of the N > 1 path (gloo, world_size 2 and 3): slab partition, neighbour / message
ordering logic shared with the library, and a rank-emulated Newton-GMRES on slabs that must
reproduce the single-domain oracle."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def test_slab_partition_covers_every_row():
    from newtonkrylov_jl_b200 import dist as D

    for gny in (1, 7, 64, 8192 * 8, 100):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                gy0, ny = D.slab_partition(gny, world, r)
                rows += list(range(gy0, gy0 + ny))
            assert rows == list(range(gny))
            sizes = [D.slab_partition(gny, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_halo_neighbours_and_message_order():
    from newtonkrylov_jl_b200 import dist as D

    assert D.halo_neighbors(0, 1, False) == (-1, -1) and D.halo_neighbors(0, 1, True) == (-1, -1)
    assert D.halo_neighbors(0, 4, False) == (-1, 1) and D.halo_neighbors(3, 4, False) == (2, -1)
    assert D.halo_neighbors(0, 4, True) == (3, 1) and D.halo_neighbors(3, 4, True) == (2, 0)
    # every send has a matching receive posted in the same per-peer order on the other side
    for world in (2, 3, 4):
        for periodic in (False, True):
            sends, recvs = {}, {}
            for r in range(world):
                for op, peer, what in D.halo_message_order(r, world, periodic):
                    (sends if op == "send" else recvs).setdefault((r, peer) if op == "send" else (peer, r), []).append(what)
            assert set(sends) == set(recvs)
            for k in sends:
                want = ["halo_lo" if w == "last_row" else "halo_hi" for w in sends[k]]
                assert recvs[k] == want, (world, periodic, k, sends[k], recvs[k])


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    import dist_emulation as E
    import problems as P
    from newtonkrylov_jl_b200 import dist as D

    nx, gny = 24, 21
    d = P.generic(P.bratu2d(nx, gny))
    gy0, ny = D.slab_partition(gny, world, rank)
    prob = E.SlabBratu2D(O, nx, ny, gny, d["dx"], d["dy"], d["lam"])
    u, hist = E.newton(prob, d["u0"][gy0:gy0 + ny].copy())
    q.put((rank, gy0, ny, u, hist))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rank_emulated_newton_gmres_matches_single_domain_oracle(oracle, world):
    import problems as P

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = P.generic(P.bratu2d(24, 21))
    ur, sr, hr = oracle.newton(P.oracle_problem(oracle, d), d["u0"])
    u = np.zeros_like(ur)
    for rank, gy0, ny, us, hist in out:
        u[gy0:gy0 + ny] = us
        assert [h["inner"] for h in hist] == [h["inner"] for h in hr]
        for a, b in zip(hist, hr):
            assert abs(a["n_res"] - b["n_res"]) <= 1e-9 * b["n_res"] + 1e-13 * hr[0]["n_res"]
    assert np.linalg.norm(u - ur) / np.linalg.norm(ur) < 1e-8
