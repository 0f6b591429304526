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


def setup_sharded_matvec(net, dist, rank, world):
    """Enable the sharded H_eff application on `net` (after nsb_extract).  Returns a handle whose `.matvec()`
    runs one sharded application; `.active` is False when the current position cannot be sharded."""
    init_comm(net.ctx, dist, rank, world)
    active = C.c_int32()
    net.ctx.check(net._lib.nsb_net_set_shard(net.handle, 1, C.byref(active)))
    return ShardedMatvec(net, bool(active.value))


def reference_sharded_matvec(parts, allreduce):
    """Host-side statement of the partition used by the CPU (gloo) tests: `parts` is this rank's partial theta'
    (NumPy), `allreduce` sums an array over ranks in place."""
    out = np.ascontiguousarray(parts)
    allreduce(out)
    return out
