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

    def enable(self, on=True):
        """Switch the sharded application on / off at the current position (off: every rank applies H_eff on its own)."""
        a = C.c_int32()
        self.net.ctx.check(self.net._lib.nsb_net_set_shard(self.net.handle, 1 if on else 0, C.byref(a)))
        self.active = bool(a.value)
        return self.active


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
    if active.value and init:
        net.matvec_device(1)          # one discarded application: NCCL sets its channels up for these buffer sizes
        net.ctx.synchronize()
    return ShardedMatvec(net, bool(active.value))


def reference_sharded_matvec(parts, allreduce):
    """Host-side statement of the partition used by the CPU (gloo) tests: `parts` is this rank's partial theta'
    (NumPy), `allreduce` sums an array over ranks in place."""
    out = np.ascontiguousarray(parts)
    allreduce(out)
    return out


class MultiDeviceNetwork:
    """State + operator replicated on several GPUs of ONE process behind an `nsb_multi` handle (include/nsb200.h): every hook
    call fans out to one host thread per device inside the library and the replicas run the sharded region step in lock
    step.  This is the entry point a single-threaded host (the Julia shim, julia/NetworkSolversB200.jl) uses to drive 8 GPUs;
    the torchrun path (one process per GPU) and this one execute the same device code."""

    def __init__(self, operator, state, devices=(0,), dtype=None, shard=True):
        from .device import DeviceNetwork
        lib = L.load()
        self._lib = lib
        g = state.graph
        self.graph, self.verts = g, g.vertices
        self.vid = {v: i for i, v in enumerate(self.verts)}
        if dtype is None:
            dtype = np.result_type(operator.dtype(), state.dtype())
        self.dtype = np.dtype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
        dt = L.NSB_C128 if self.dtype.kind == "c" else L.NSB_F64
        dev = np.array(list(devices), dtype=np.int32)
        h = C.c_void_p()
        L.check(lib.nsb_multi_create(dev.ctypes.data_as(C.POINTER(C.c_int32)), len(dev), C.byref(h)))
        self.handle = h
        self.ndev = len(dev)
        edges = np.array([[self.vid[u], self.vid[v]] for u, v in g.edges], dtype=np.int32).reshape(-1)
        sdims = np.array([state.tensors[v].shape[state.legs[v].index(("site", v))] for v in self.verts], dtype=np.int64)
        self._check(lib.nsb_multi_network_create(h, len(self.verts), edges.ctypes.data_as(C.POINTER(C.c_int32)), len(g.edges),
                                                 sdims.ctypes.data_as(C.POINTER(C.c_int64)), dt))
        self._enc = DeviceNetwork._encode.__get__(self)
        self._dec = DeviceNetwork._decode.__get__(self)
        for v in self.verts:
            self._upload(v, operator.tensors[v], operator.legs[v], True)
            self._upload(v, state.tensors[v], state.legs[v], False)
        arr = np.array([self.vid[v] for v in state.ortho_region], dtype=np.int32)
        self._check(lib.nsb_multi_set_ortho_region(h, arr.ctypes.data_as(C.POINTER(C.c_int32)), len(arr)))
        self.shard = shard and self.ndev > 1
        n0 = C.c_void_p()
        self._check(lib.nsb_multi_net(h, 0, C.byref(n0)))
        self._net0 = n0

    def _check(self, code):
        if code != L.NSB_OK:
            raise L.NsbError(code, self._lib.nsb_multi_last_error(self.handle).decode(errors="replace"))

    def _upload(self, v, arr, legs, is_operator):
        a = np.asfortranarray(arr, dtype=self.dtype)
        enc = self._enc(legs)
        dims = np.array(a.shape, dtype=np.int64)
        fn = self._lib.nsb_multi_mpo_upload if is_operator else self._lib.nsb_multi_site_upload
        self._check(fn(self.handle, self.vid[v], a.ndim, enc.ctypes.data_as(C.POINTER(C.c_int32)),
                       dims.ctypes.data_as(C.POINTER(C.c_int64)), a.ctypes.data))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self._lib.nsb_multi_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def extract(self, region, trunc=None, expand=None):
        reg = np.array([self.vid[v] for v in region], dtype=np.int32)
        tr = L.Trunc(*(trunc or (0.0, 1, L.INT64_MAX)))
        info = L.ExtractInfo()
        ex = None
        if expand is not None:
            ex = L.Expand(expand["algorithm"], expand.get("north_pass", 1), expand.get("expansion_factor", 1.5),
                          expand.get("max_expand", L.INT64_MAX))
        self._check(self._lib.nsb_multi_extract(self.handle, reg.ctypes.data_as(C.POINTER(C.c_int32)), len(reg), C.byref(tr),
                                                C.byref(ex) if ex is not None else None, C.byref(info)))
        if self.shard:
            a = C.c_int32()
            self._check(self._lib.nsb_multi_set_shard(self.handle, 1, C.byref(a)))
            self.shard_active = bool(a.value)
        return info

    def update_eigsolve(self, krylovdim=3, maxiter=1, tol=1e-14, which="SR", eager=False):
        kp = L.Krylov(krylovdim, maxiter, tol, 0 if which in ("SR", ":SR") else 1, 1 if eager else 0, 4, 0)
        val, info = C.c_double(), L.SolveInfo()
        self._check(self._lib.nsb_multi_update_eigsolve(self.handle, C.byref(kp), C.byref(val), C.byref(info)))
        return val.value, info

    def update_exp(self, t, solver="rk", order=4, krylovdim=30, maxiter=100, tol=1e-12, eager=True, nsites=2, next_vertex=None):
        t = complex(t)
        kp = L.Krylov(krylovdim, maxiter, tol, 0, 1 if eager else 0, order, 0)
        info = L.SolveInfo()
        nv = -1 if next_vertex is None else self.vid[next_vertex]
        self._check(self._lib.nsb_multi_update_exp(self.handle, t.real, t.imag, L.NSB_SOLVER_RK if solver == "rk" else L.NSB_SOLVER_KRYLOV,
                                                   C.byref(kp), nsites, nv, C.byref(info)))
        return info

    def insert(self, trunc=None, normalize=False, set_ortho=True):
        tr = L.Trunc(*(trunc or (0.0, 1, L.INT64_MAX)))
        info = L.InsertInfo()
        self._check(self._lib.nsb_multi_insert(self.handle, C.byref(tr), 1 if normalize else 0, 1 if set_ortho else 0, C.byref(info)))
        return info

    def local_download(self):
        rank = C.c_int32()
        legs = (C.c_int32 * 32)()
        dims = (C.c_int64 * 16)()
        L.check(self._lib.nsb_local_info(self._net0, C.byref(rank), legs, dims))
        shape = [dims[i] for i in range(rank.value)]
        out = np.empty(shape, dtype=self.dtype, order="F")
        self._check(self._lib.nsb_multi_local_download(self.handle, out.ctypes.data))
        return out, self._dec(legs, rank.value)

    def maxlinkdim(self):
        d = C.c_int64()
        L.check(self._lib.nsb_maxlinkdim(self._net0, C.byref(d)))
        return d.value
