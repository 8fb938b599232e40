"""Multi-process plumbing for the phi split: one process per GPU, torch.distributed for the
rendezvous only.  The data path has no collective: each rank pushes its boundary rows straight into
its neighbours' ghost buffers through CUDA-IPC peer mappings over NVLink (crd_grid_halo_*), and the
integrator's norms cross ranks as one 8-byte host-level allreduce.

Replaces the reference's MPI_Cart_create / MPI_Cart_shift neighbour discovery
(src/FHNmodel_torus.cpp:724-732,799-811) for a 1-D periodic ring: prev = rank-1, next = rank+1.
"""
from .api import CRD_MAX, CRD_MIN, CRD_SUM, decomp_phi


def ring_neighbours(rank, world):
    """(prev, next) on the periodic phi ring; prev owns the rows below js, next those above je."""
    return (rank - 1) % world, (rank + 1) % world


def slab_extents(ny, world):
    """[(js, je)] for every rank — SetupDecomp's formula with dims = {1, world}."""
    return [decomp_phi(ny, world, r) for r in range(world)]


def exchange_handles(handle, group=None):
    """all-gather the 64-byte halo handles of all ranks (any torch.distributed backend)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, bytes(handle), group=group)
    return out


def ring_connect(grid, rank, world, handles):
    """Open the two neighbours' ghost blocks.  handles: list of all ranks' halo handles."""
    if world == 1:
        return
    prev, nxt = ring_neighbours(rank, world)
    grid.halo_connect_ipc(handles[prev], handles[nxt])


def comm_connect(ctx, rank, world, group=None):
    """Wire the device-side allreduce: all-gather the contexts' mailbox handles, connect to every rank.  After this the
    integrator's norms never leave the GPUs (no host-level allreduce hook is called)."""
    if world == 1:
        return
    ctx.comm_connect_ipc(rank, world, exchange_handles(ctx.comm_handle(), group))


def make_allreduce(group=None):
    """Host-level allreduce hook for Context.set_comm built on a (gloo) process group."""
    import torch
    import torch.distributed as dist
    ops = {CRD_SUM: dist.ReduceOp.SUM, CRD_MAX: dist.ReduceOp.MAX, CRD_MIN: dist.ReduceOp.MIN}

    def allreduce(vals, op):
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t, op=ops[op], group=group)
        return t.tolist()
    return allreduce
