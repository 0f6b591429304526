"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed carries the NCCL unique id, the
library runs its own communicator on the context's stream.

Partition: the state, environments and Krylov vectors are replicated and every rank executes the
same sweep in lock step; each H_eff application is split along the last bond of theta -- rank r contracts
columns [lo_r, hi_r) of theta (1/nranks of the L.theta, MPO and .R work) and one NCCL all-reduce sums the partial
theta'.  `shard_bounds` is the single source of truth for the split (the C library uses the same rule)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def shard_bounds(dim, rank, nranks):
    """Contiguous slab [lo, hi) of a bond of dimension `dim` owned by `rank` (ceil-division blocks)."""
    per = (dim + nranks - 1) // nranks
    lo = min(dim, per * rank)
    return lo, min(dim, lo + per)


def init_comm(ctx, dist, rank, world):
    """Create the library-side NCCL communicator; `dist` is an initialised torch.distributed (any backend)."""
    import torch
    lib = ctx._lib
    buf = C.create_string_buffer(128)
    if rank == 0:
        L.check(lib.nsb_comm_unique_id(buf))
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    ctx.check(lib.nsb_comm_init(ctx.handle, raw, rank, world))


class ShardedMatvec:
    def __init__(self, net, active):
        self.net, self.active = net, active

    def matvec(self, reps=1):
        self.net.matvec_device(reps)


def setup_peer_windows(ctx, dist, rank, world, nbytes):
    """Create this rank's staging window and map every peer's (cudaIpc handles travel through torch.distributed)."""
    import torch
    lib = ctx._lib
    buf = C.create_string_buffer(64)
    ctx.check(lib.nsb_peer_window_create(ctx.handle, int(nbytes), buf))
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allh, mine)
    for r in range(world):
        ctx.check(lib.nsb_peer_window_open(ctx.handle, r, bytes(allh[r].cpu().tolist())))
    dist.barrier()


def setup_sharded_matvec(net, dist, rank, world, fused=False, init=True):
    """Enable the sharded H_eff application on `net` (after nsb_extract).  Returns a handle whose `.matvec()`
    runs one sharded application; `.active` is False when the current position cannot be sharded.  With
    `fused=True` the last GEMM reduces through peer-memory stores in its epilogue (see include/nsb200.h)."""
    if init:
        init_comm(net.ctx, dist, rank, world)
    if fused:
        _, dims = net.local_info()
        setup_peer_windows(net.ctx, dist, rank, world, int(np.prod(dims)) * net.dtype.itemsize)
    net.ctx.set_option("shard_fused", 1 if fused else 0)
    active = C.c_int32()
    net.ctx.check(net._lib.nsb_net_set_shard(net.handle, 1, C.byref(active)))
    return ShardedMatvec(net, bool(active.value))


def reference_sharded_matvec(parts, allreduce):
    """Host-side statement of the partition used by the CPU (gloo) tests: `parts` is this rank's partial theta'
    (NumPy), `allreduce` sums an array over ranks in place."""
    out = np.ascontiguousarray(parts)
    allreduce(out)
    return out
