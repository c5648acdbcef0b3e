"""Slab domain decomposition across GPUs (one process per GPU), modelled on the ghost-cell layout
of examples/halovector.jl: every rank owns `ny` consecutive rows of the global nx x gny grid, the
logical vector is the interior only, and the two ghost rows are filled from the neighbouring
ranks inside the stencil launches (the distributed form of bc!(u), examples/heat_2D.jl:51).

`torch.distributed` is plumbing only: it carries the 128-byte NCCL id from rank 0 to the other
ranks (and test data); every data-path collective is issued by libariadne_b200.so itself.
"""
import os

from . import host


def slab_partition(gny, world, rank):
    """Rows [gy0, gy0 + ny) owned by `rank`: the first gny % world ranks get one extra row."""
    base, extra = divmod(int(gny), int(world))
    ny = base + (1 if rank < extra else 0)
    gy0 = rank * base + min(rank, extra)
    return gy0, ny


def halo_neighbors(rank, world, periodic):
    """(down, up): owner of global row gy0-1 and of row gy0+ny; -1 = physical Dirichlet boundary.
    Mirrors exchange_halo_rows() in csrc/context.cu."""
    down = rank - 1 if rank > 0 else (world - 1 if periodic else -1)
    up = rank + 1 if rank < world - 1 else (0 if periodic else -1)
    if world == 1:
        return (-1, -1)
    return down, up


def halo_message_order(rank, world, periodic):
    """Posting order of the point-to-point messages of one halo exchange, as (op, peer, what).
    NCCL matches sends and receives per peer in posting order; with world == 2 and periodic wrap both
    neighbours are the same rank, hence "last row up" before "first row down" and "lo" before "hi"."""
    down, up = halo_neighbors(rank, world, periodic)
    ops = []
    if up >= 0:
        ops.append(("send", up, "last_row"))
    if down >= 0:
        ops.append(("send", down, "first_row"))
    if down >= 0:
        ops.append(("recv", down, "halo_lo"))
    if up >= 0:
        ops.append(("recv", up, "halo_hi"))
    return ops


def init_distributed(device=None, p2p_halo=None):
    """Create the context of this rank and its NCCL communicator.  Expects torch.distributed to be
    initialised (any backend) and RANK / LOCAL_RANK / WORLD_SIZE in the environment.
    `p2p_halo` (row length of the widest 2-D grid) additionally maps the ranks' peer memory so that the
    pair-wise GMRES sweep reduces and exchanges ghost rows with NVLink stores instead of NCCL calls."""
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = host.get_context(device)
    if world > 1 and ctx.nranks == 1:
        ids = [host.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_comm(world, rank, ids[0])
    if world > 1 and p2p_halo is not None and not ctx.p2p:
        ctx.enable_p2p(p2p_halo)
    return ctx
