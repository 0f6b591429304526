"""Device-side objects: `Context` (one GPU, one stream) and `DeviceNetwork` (state + operator +
projected-operator environments resident in HBM behind an opaque libnsb200 handle)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .models import HostTTN, canonical_legs

_default_ctx = None


class Context:
    def __init__(self, device=0):
        lib = L.load()
        self._lib = lib
        h = C.c_void_p()
        L.check(lib.nsb_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = device

    def close(self):
        if self.handle is not None and self.handle.value:
            self._lib.nsb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, code):
        L.check(code, self.handle)

    def set_option(self, key, value):
        self.check(self._lib.nsb_ctx_set_option(self.handle, key.encode(), int(value)))

    def synchronize(self):
        self.check(self._lib.nsb_ctx_synchronize(self.handle))

    def counters(self):
        c = L.Counters()
        self.check(self._lib.nsb_ctx_counters(self.handle, C.byref(c)))
        return c.as_dict()

    def reset_counters(self):
        self.check(self._lib.nsb_ctx_counters_reset(self.handle))

    def tic(self):
        self.check(self._lib.nsb_event_tic(self.handle))

    def toc(self):
        ms = C.c_double()
        self.check(self._lib.nsb_event_toc(self.handle, C.byref(ms)))
        return ms.value

    def enable_timers(self, on=True):
        self.check(self._lib.nsb_timers_enable(self.handle, 1 if on else 0))

    def timers(self):
        arr = (C.c_double * L.NSB_NUM_TIMERS)()
        self.check(self._lib.nsb_timers_get(self.handle, arr))
        return dict(zip(L.TIMER_NAMES, list(arr)))

    def reset_timers(self):
        self.check(self._lib.nsb_timers_reset(self.handle))

    def mem_info(self):
        f, t, u = C.c_int64(), C.c_int64(), C.c_int64()
        self.check(self._lib.nsb_mem_info(self.handle, C.byref(f), C.byref(t), C.byref(u)))
        return dict(free=f.value, total=t.value, pool_used=u.value)

    def profiler(self, on=True):
        """cudaProfilerStart / Stop (for `ncu --profile-from-start off`)."""
        self.check(self._lib.nsb_profiler(self.handle, 1 if on else 0))

    def gemm_profile(self, on=True):
        """Start (and clear) / stop the per-launch CUDA-event timing of the GEMM kernels."""
        self.check(self._lib.nsb_gemm_profile_enable(self.handle, 1 if on else 0))

    def gemm_profile_read(self):
        """[(ms, flops, (M, N, K, batch)), ...] of the GEMM launches since gemm_profile(True)."""
        n = C.c_int64()
        self.check(self._lib.nsb_gemm_profile_read(self.handle, 0, None, None, None, C.byref(n)))
        cap = max(n.value, 1)
        ms, fl, mnk = (C.c_double * cap)(), (C.c_double * cap)(), (C.c_int64 * (4 * cap))()
        self.check(self._lib.nsb_gemm_profile_read(self.handle, cap, ms, fl, mnk, C.byref(n)))
        return [(ms[i], fl[i], tuple(mnk[4 * i: 4 * i + 4])) for i in range(n.value)]

    # ---- dense helpers (kernels under the hooks, exposed for tests / benchmarks) ----
    @staticmethod
    def _dt(a):
        return L.NSB_C128 if np.iscomplexobj(a) else L.NSB_F64

    def gemm(self, A, B, opa="N", opb="N", impl=0):
        """C = op(A) op(B) through the device GEMM; A, B are 2-D numpy arrays (any layout)."""
        ops = {"N": 0, "T": 1, "C": 2, "J": 3}
        cplx = np.iscomplexobj(A) or np.iscomplexobj(B)
        dt = np.complex128 if cplx else np.float64
        Af, Bf = np.asfortranarray(A, dtype=dt), np.asfortranarray(B, dtype=dt)
        m, k = (Af.shape if opa in "NJ" else Af.shape[::-1])
        k2, n = (Bf.shape if opb in "NJ" else Bf.shape[::-1])
        assert k == k2, (Af.shape, Bf.shape, opa, opb)
        Cf = np.empty((m, n), dtype=dt, order="F")
        self.check(self._lib.nsb_gemm_host(self.handle, L.NSB_C128 if cplx else L.NSB_F64, ops[opa], ops[opb], m, n, k,
                                            Af.ctypes.data, Af.shape[0], Bf.ctypes.data, Bf.shape[0],
                                            Cf.ctypes.data, m, impl))
        return Cf

    def gemm_bench(self, m, n, k, opa="N", opb="N", dtype=np.float64, impl=0, reps=5):
        ops = {"N": 0, "T": 1, "C": 2, "J": 3}
        ms = C.c_double()
        dt = L.NSB_C128 if np.dtype(dtype).kind == "c" else L.NSB_F64
        self.check(self._lib.nsb_gemm_bench(self.handle, dt, ops[opa], ops[opb], m, n, k, impl, reps, C.byref(ms)))
        return ms.value

    def dmma_peak_tflops(self):
        v = C.c_double()
        self.check(self._lib.nsb_dmma_peak(self.handle, C.byref(v)))
        return v.value

    def factorize(self, M, cutoff=0.0, mindim=1, maxdim=None):
        """Truncated left-orthogonal factorisation M = U C.  Returns U, C, spectrum (sigma^2), info."""
        cplx = np.iscomplexobj(M)
        dt = np.complex128 if cplx else np.float64
        Mf = np.asfortranarray(M, dtype=dt)
        rows, cols = Mf.shape
        k = min(rows, cols)
        U = np.empty((rows, k), dtype=dt, order="F")
        Cm = np.empty((k * cols,), dtype=dt)
        spec = np.empty(k)
        tr = L.Trunc(cutoff, mindim, L.INT64_MAX if maxdim is None else int(maxdim))
        info = L.InsertInfo()
        self.check(self._lib.nsb_factorize_host(self.handle, self._dt(Mf), rows, cols, Mf.ctypes.data, C.byref(tr),
                                                 U.ctypes.data, Cm.ctypes.data, spec.ctypes.data_as(C.POINTER(C.c_double)),
                                                 C.byref(info)))
        nk = info.newdim
        Uo = np.asfortranarray(U.reshape(-1, order="F")[: rows * nk].reshape((rows, nk), order="F"))
        Co = Cm[: nk * cols].reshape((nk, cols), order="F")
        return Uo, Co, spec, dict(newdim=nk, truncerr=info.truncerr, decomp=info.decomp, sweeps=info.jacobi_sweeps)

    def eigh(self, A, vectors=True):
        """Hermitian eigen-decomposition on the device (tridiagonalisation + divide & conquer): w ascending, U."""
        cplx = np.iscomplexobj(A)
        dt = np.complex128 if cplx else np.float64
        Af = np.asfortranarray(A, dtype=dt)
        n = Af.shape[0]
        w = np.empty(n)
        U = np.empty((n, n), dtype=dt, order="F") if vectors else None
        self.check(self._lib.nsb_eigh_host(self.handle, self._dt(Af), n, Af.ctypes.data, w.ctypes.data_as(C.POINTER(C.c_double)),
                                            U.ctypes.data if vectors else None))
        return (w, U) if vectors else w

    def sbr_chase(self, ab, b):
        """EXPERIMENTAL: band -> tridiagonal by bulge chasing on the device (nsb_sbr_chase_host).  ab: (2 b + 1) x n lower band
        storage (Fortran order, bulge rows zero); returns (ab_out, V2, tau2)."""
        ab = np.asfortranarray(ab, dtype=np.float64).copy(order="F")
        ld, n = ab.shape
        nst = max(1, -(-(n - 1) // b))
        V2 = np.zeros((n, n), order="F")
        tau2 = np.zeros((nst, n), order="F")
        self.check(self._lib.nsb_sbr_chase_host(self.handle, n, int(b), ab.ctypes.data, ld, V2.ctypes.data, tau2.ctypes.data, nst))
        return ab, V2, tau2

    def qr(self, M):
        cplx = np.iscomplexobj(M)
        dt = np.complex128 if cplx else np.float64
        Mf = np.asfortranarray(M, dtype=dt)
        rows, cols = Mf.shape
        k = min(rows, cols)
        Q = np.empty((rows, k), dtype=dt, order="F")
        R = np.empty((k, cols), dtype=dt, order="F")
        self.check(self._lib.nsb_qr_host(self.handle, self._dt(Mf), rows, cols, Mf.ctypes.data, Q.ctypes.data, R.ctypes.data))
        return Q, R

    def qr_bench(self, rows, cols, dtype=np.float64, reps=3):
        ms = C.c_double()
        dt = L.NSB_C128 if np.dtype(dtype).kind == "c" else L.NSB_F64
        self.check(self._lib.nsb_qr_bench(self.handle, dt, rows, cols, reps, C.byref(ms)))
        return ms.value

    def range_finder(self, A, max_rank, oversample=2, north_pass=2, orthogonal_threshold=1e-12, cutoff=0.0, seed=1, probes=None):
        """Orthonormal basis of the range of the matrix A from random probes (range_finder.jl:54-64 with
        linear_map = A).  probes: (n, sketch) array whose column k is the k-th random_vector(), or None (device Philox)."""
        cplx = np.iscomplexobj(A)
        dt = np.complex128 if cplx else np.float64
        Af = np.asfortranarray(A, dtype=dt)
        m, n = Af.shape
        cap = min(max_rank + oversample, m, n)
        Q = np.empty((m, max(cap, 1)), dtype=dt, order="F")
        rank = C.c_int64()
        pr = None
        if probes is not None:
            pr = np.asfortranarray(probes, dtype=dt)
            assert pr.shape[0] == n and pr.shape[1] >= cap, (pr.shape, n, cap)
        self.check(self._lib.nsb_range_finder_host(self.handle, self._dt(Af), m, n, Af.ctypes.data,
                                                    pr.ctypes.data if pr is not None else None, max_rank, oversample,
                                                    north_pass, orthogonal_threshold, cutoff, seed, Q.ctypes.data, C.byref(rank)))
        return Q[:, : rank.value]


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class DeviceNetwork:
    """State + operator of one sweep problem on the device.  Built from host tensors by
    `EigsolveProblem` / `ApplyExpProblem` construction (the counterpart of `permute_indices` +
    `itn.ProjTTN(H)` in src/eigsolve.jl:69-74, src/applyexp.jl:84-89)."""

    def __init__(self, operator: HostTTN, state: HostTTN, dtype=None, ctx: Context = None):
        self.ctx = ctx or default_context()
        self._lib = self.ctx._lib
        g = state.graph
        assert g.is_tree(), "the network must be a tree"
        self.graph = g
        self.verts = g.vertices
        self.vid = {v: i for i, v in enumerate(self.verts)}
        if dtype is None:
            dtype = np.result_type(operator.dtype(), state.dtype())
        self.dtype = np.dtype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
        self._dt = L.NSB_C128 if self.dtype.kind == "c" else L.NSB_F64
        edges = np.array([[self.vid[u], self.vid[v]] for u, v in g.edges], dtype=np.int32).reshape(-1)
        sdims = np.array([state.tensors[v].shape[state.legs[v].index(("site", v))] for v in self.verts], dtype=np.int64)
        h = C.c_void_p()
        self.ctx.check(self._lib.nsb_network_create(self.ctx.handle, len(self.verts),
                                                     edges.ctypes.data_as(C.POINTER(C.c_int32)), len(g.edges),
                                                     sdims.ctypes.data_as(C.POINTER(C.c_int64)), self._dt, C.byref(h)))
        self.handle = h
        for v in self.verts:
            self._upload(v, operator.tensors[v], operator.legs[v], True)
            self._upload(v, state.tensors[v], state.legs[v], False)
        self.set_ortho_region(state.ortho_region)
        self.qn_enabled = False
        if getattr(state, "qn", None) is not None:
            self._upload_qn(state.qn)

    def _upload_qn(self, qn):
        tot = np.ascontiguousarray(qn["total"], dtype=np.int32)
        i32p = C.POINTER(C.c_int32)
        self.ctx.check(self._lib.nsb_qn_enable(self.handle, len(tot), tot.ctypes.data_as(i32p)))
        for v in self.verts:
            a = np.ascontiguousarray(qn["site"][v], dtype=np.int32)
            self.ctx.check(self._lib.nsb_qn_set_site(self.handle, self.vid[v], a.ctypes.data_as(i32p)))
        for (u, v), arr in qn["link"].items():
            a = np.ascontiguousarray(arr, dtype=np.int32)
            self.ctx.check(self._lib.nsb_qn_set_link(self.handle, self.vid[u], self.vid[v], a.ctypes.data_as(i32p)))
        self.qn_enabled = True
        self._qn_static = dict(total=np.array(qn["total"]), site={v: np.array(a) for v, a in qn["site"].items()})

    def link_charges(self, u, v):
        """Charges (linkdim, nq) of the subtree on u's side of the edge {u, v}."""
        nq = len(self._qn_static["total"])
        out = np.empty((self.linkdim(u, v), nq), dtype=np.int32)
        self.ctx.check(self._lib.nsb_qn_get_link(self.handle, self.vid[u], self.vid[v], out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out.astype(np.int64)

    @classmethod
    def synthetic(cls, operator: HostTTN, sites, chi, seed=1234, dtype=np.float64, ctx=None, ortho_region=None, canonical=False):
        """Network whose state tensors are filled on the device (Philox N(0,1), scaled so that environments stay
        O(1)) with uniform bond dimension chi capped by d^k at the ends: the synthetic state of SURVEY 8(d)
        config 2, without moving 25 GiB over PCIe."""
        from .models import product_state, bond_dims
        g = sites.graph
        placeholder = product_state(sites, {v: 0 for v in g.vertices})
        net = cls(operator, placeholder, dtype=dtype, ctx=ctx)
        dims = bond_dims(g, sites.dim, chi)
        for i, v in enumerate(g.vertices):
            legs = canonical_legs(g, v)
            d = {l: (sites.dim if l[0] == "site" else dims[(l[1], l[2])]) for l in legs}
            numel = int(np.prod([d[l] for l in legs]))
            scale = 1.0 / np.sqrt(max(numel / d[legs[-1]], 1.0)) if len(legs) > 1 else 1.0
            net.fill_random(v, d, seed + 7 * i, scale)
        if canonical:
            # random tensors carry no gauge: declare every vertex part of the orthogonality region, so that the first
            # extract walks the whole network (one device QR per edge) exactly as for itn.random_mps / random_tensornetwork
            net.set_ortho_region(list(g.vertices))
        elif ortho_region is not None:
            net.set_ortho_region(ortho_region)     # gauge flag only (tensors stay as filled): skips the walk
        return net

    @classmethod
    def synthetic_qn(cls, operator: HostTTN, sites, link_charges, total, site_charges, seed=1234, dtype=np.float64, ctx=None):
        """Random QN-conserving state filled on the device: `link_charges[(u, v)]` is the (dim, nq) table of charges of the
        subtree on u's side for every edge (it fixes the sector dimensions), every site tensor is filled with Philox N(0,1)
        numbers and projected onto its symmetry-allowed blocks (nsb_qn_project).  The orthogonality region is all vertices,
        so the first extract orthonormalises the state sector by sector."""
        from .models import product_state
        g = sites.graph
        placeholder = product_state(sites, {v: 0 for v in g.vertices}, conserve_qns=False)
        net = cls(operator, placeholder, dtype=dtype, ctx=ctx)
        i32p = C.POINTER(C.c_int32)
        tot = np.ascontiguousarray(total, dtype=np.int32)
        net.ctx.check(net._lib.nsb_qn_enable(net.handle, len(tot), tot.ctypes.data_as(i32p)))
        sc = np.ascontiguousarray(site_charges, dtype=np.int32)
        for v in net.verts:
            net.ctx.check(net._lib.nsb_qn_set_site(net.handle, net.vid[v], sc.ctypes.data_as(i32p)))
        dims = {}
        for (u, v), arr in link_charges.items():
            dims[(u, v)] = dims[(v, u)] = len(arr)
        for i, v in enumerate(g.vertices):
            legs = canonical_legs(g, v)
            d = {l: (sites.dim if l[0] == "site" else dims[(l[1], l[2])]) for l in legs}
            numel = int(np.prod([d[l] for l in legs]))
            scale = 1.0 / np.sqrt(max(numel / d[legs[-1]], 1.0)) if len(legs) > 1 else 1.0
            net.fill_random(v, d, seed + 7 * i, scale)
        for (u, v), arr in link_charges.items():
            a = np.ascontiguousarray(arr, dtype=np.int32)
            net.ctx.check(net._lib.nsb_qn_set_link(net.handle, net.vid[u], net.vid[v], a.ctypes.data_as(i32p)))
        for v in net.verts:
            net.ctx.check(net._lib.nsb_qn_project(net.handle, net.vid[v]))
        net.qn_enabled = True
        net._qn_static = dict(total=np.array(total), site={v: np.array(site_charges) for v in net.verts})
        net.set_ortho_region(list(g.vertices))
        return net

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self._lib.nsb_network_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- leg encoding -------------------------------------------------------------------------
    def _encode(self, legs):
        out = []
        for l in legs:
            if l[0] == "site":
                out += [self.vid[l[1]], L.NSB_SITE]
            elif l[0] == "site_out":
                out += [self.vid[l[1]], L.NSB_SITE_OUT]
            else:
                out += [self.vid[l[1]], self.vid[l[2]]]
        return np.array(out, dtype=np.int32)

    def _decode(self, legs_arr, rank, owner=None):
        out = []
        for i in range(rank):
            a, b = int(legs_arr[2 * i]), int(legs_arr[2 * i + 1])
            if b == L.NSB_SITE:
                out.append(("site", self.verts[a]))
            elif b == L.NSB_SITE_OUT:
                out.append(("site_out", self.verts[a]))
            elif b >= 0:
                u, v = self.verts[a], self.verts[b]
                if owner is not None and u != owner:
                    u, v = v, u
                out.append(("link", u, v))
            else:
                out.append(("aux", a))
        return out

    def _upload(self, v, arr, legs, is_operator):
        a = np.asfortranarray(arr, dtype=self.dtype)
        enc = self._encode(legs)
        dims = np.array(a.shape, dtype=np.int64)
        fn = self._lib.nsb_mpo_upload if is_operator else self._lib.nsb_site_upload
        self.ctx.check(fn(self.handle, self.vid[v], a.ndim, enc.ctypes.data_as(C.POINTER(C.c_int32)),
                          dims.ctypes.data_as(C.POINTER(C.c_int64)), a.ctypes.data))

    # ---- fitting (src/fitting.jl) -------------------------------------------------------------
    def set_fit_target(self, target: HostTTN):
        """Upload the target network |x> (own link dimensions) and switch the network to fitting mode."""
        for v in self.verts:
            a = np.asfortranarray(target.tensors[v], dtype=self.dtype)
            enc = self._encode(target.legs[v])
            dims = np.array(a.shape, dtype=np.int64)
            self.ctx.check(self._lib.nsb_fit_target_upload(self.handle, self.vid[v], a.ndim, enc.ctypes.data_as(C.POINTER(C.c_int32)),
                                                            dims.ctypes.data_as(C.POINTER(C.c_int64)), a.ctypes.data))

    def update_fit(self):
        ov = C.c_double()
        self.ctx.check(self._lib.nsb_update_fit(self.handle, C.byref(ov)))
        return ov.value

    def fill_random(self, v, dims_by_leg, seed, scale):
        legs = canonical_legs(self.graph, v)
        enc = self._encode(legs)
        dims = np.array([dims_by_leg[l] for l in legs], dtype=np.int64)
        self.ctx.check(self._lib.nsb_site_fill_random(self.handle, self.vid[v], len(legs),
                                                       enc.ctypes.data_as(C.POINTER(C.c_int32)),
                                                       dims.ctypes.data_as(C.POINTER(C.c_int64)), seed, scale))

    # ---- state access -------------------------------------------------------------------------
    def site(self, v):
        rank = C.c_int32()
        legs = (C.c_int32 * 32)()
        dims = (C.c_int64 * 16)()
        self.ctx.check(self._lib.nsb_site_info(self.handle, self.vid[v], C.byref(rank), legs, dims))
        shape = [dims[i] for i in range(rank.value)]
        out = np.empty(shape, dtype=self.dtype, order="F")
        self.ctx.check(self._lib.nsb_site_download(self.handle, self.vid[v], out.ctypes.data))
        return out, self._decode(legs, rank.value, owner=v)

    def to_host(self) -> HostTTN:
        tensors, legs = {}, {}
        for v in self.verts:
            tensors[v], legs[v] = self.site(v)
        qn = None
        if self.qn_enabled:
            qn = dict(self._qn_static, link={(u, v): self.link_charges(u, v) for u, v in self.graph.edges})
        return HostTTN(self.graph, tensors, legs, ortho_region=self.ortho_region(), qn=qn)

    def set_ortho_region(self, verts):
        arr = np.array([self.vid[v] for v in verts], dtype=np.int32)
        self.ctx.check(self._lib.nsb_set_ortho_region(self.handle, arr.ctypes.data_as(C.POINTER(C.c_int32)), len(arr)))

    def ortho_region(self):
        arr = (C.c_int32 * len(self.verts))()
        n = C.c_int32()
        self.ctx.check(self._lib.nsb_get_ortho_region(self.handle, arr, C.byref(n)))
        return [self.verts[arr[i]] for i in range(n.value)]

    def linkdim(self, u, v):
        d = C.c_int64()
        self.ctx.check(self._lib.nsb_linkdim(self.handle, self.vid[u], self.vid[v], C.byref(d)))
        return d.value

    def linkdims(self):
        return {(u, v): self.linkdim(u, v) for u, v in self.graph.edges}

    def maxlinkdim(self):
        d = C.c_int64()
        self.ctx.check(self._lib.nsb_maxlinkdim(self.handle, C.byref(d)))
        return d.value

    def norm(self):
        d = C.c_double()
        self.ctx.check(self._lib.nsb_norm(self.handle, C.byref(d)))
        return d.value

    def env_drop_all(self):
        """Forget the cached environments (a freshly constructed ProjTTN)."""
        self.ctx.check(self._lib.nsb_env_drop_all(self.handle))

    def env_count(self):
        n = C.c_int32()
        self.ctx.check(self._lib.nsb_env_count(self.handle, C.byref(n)))
        return n.value

    # ---- hooks --------------------------------------------------------------------------------
    def extract(self, region, trunc=None, expand=None):
        reg = np.array([self.vid[v] for v in region], dtype=np.int32)
        tr = L.Trunc(*(trunc or (0.0, 1, L.INT64_MAX)))
        info = L.ExtractInfo()
        ex = None
        if expand is not None:
            ex = L.Expand(expand["algorithm"], expand.get("north_pass", 1), expand.get("expansion_factor", 1.5),
                          expand.get("max_expand", L.INT64_MAX))
        self.ctx.check(self._lib.nsb_extract(self.handle, reg.ctypes.data_as(C.POINTER(C.c_int32)), len(reg), C.byref(tr),
                                              C.byref(ex) if ex is not None else None, C.byref(info)))
        return info

    def update_eigsolve(self, krylovdim=3, maxiter=1, tol=1e-14, which="SR", eager=False):
        kp = L.Krylov(krylovdim, maxiter, tol, 0 if which in ("SR", ":SR") else 1, 1 if eager else 0, 4, 0)
        val = C.c_double()
        info = L.SolveInfo()
        self.ctx.check(self._lib.nsb_update_eigsolve(self.handle, C.byref(kp), C.byref(val), C.byref(info)))
        return val.value, info

    def update_exp(self, t, solver="rk", order=4, krylovdim=30, maxiter=100, tol=1e-12, eager=True, nsites=2,
                   next_vertex=None):
        t = complex(t)
        kp = L.Krylov(krylovdim, maxiter, tol, 0, 1 if eager else 0, order, 0)
        info = L.SolveInfo()
        nv = -1 if next_vertex is None else self.vid[next_vertex]
        self.ctx.check(self._lib.nsb_update_exp(self.handle, t.real, t.imag,
                                                 L.NSB_SOLVER_RK if solver == "rk" else L.NSB_SOLVER_KRYLOV,
                                                 C.byref(kp), nsites, nv, C.byref(info)))
        return info

    def insert(self, trunc=None, normalize=False, set_ortho=True):
        tr = L.Trunc(*(trunc or (0.0, 1, L.INT64_MAX)))
        info = L.InsertInfo()
        self.ctx.check(self._lib.nsb_insert(self.handle, C.byref(tr), 1 if normalize else 0, 1 if set_ortho else 0,
                                             C.byref(info)))
        return info

    # ---- local tensor -------------------------------------------------------------------------
    def local_info(self):
        rank = C.c_int32()
        legs = (C.c_int32 * 32)()
        dims = (C.c_int64 * 16)()
        self.ctx.check(self._lib.nsb_local_info(self.handle, C.byref(rank), legs, dims))
        return self._decode(legs, rank.value), [dims[i] for i in range(rank.value)]

    def local_download(self):
        legs, dims = self.local_info()
        out = np.empty(dims, dtype=self.dtype, order="F")
        self.ctx.check(self._lib.nsb_local_download(self.handle, out.ctypes.data))
        return out, legs

    def local_upload(self, arr):
        a = np.asfortranarray(arr, dtype=self.dtype)
        self.ctx.check(self._lib.nsb_local_upload(self.handle, a.ctypes.data))

    def matvec_host(self, x):
        a = np.asfortranarray(x, dtype=self.dtype)
        out = np.empty(a.shape, dtype=self.dtype, order="F")
        self.ctx.check(self._lib.nsb_matvec_host(self.handle, a.ctypes.data, out.ctypes.data))
        return out

    def env_bytes(self):
        """(resident, replicated): bytes of environment tensors on this GPU, and their size if every one were held in full."""
        r, f = C.c_int64(), C.c_int64()
        self.ctx.check(self._lib.nsb_env_bytes(self.handle, C.byref(r), C.byref(f)))
        return r.value, f.value

    def shard_range(self):
        """(lo, hi, last_dim): this rank's slab of the local tensor's last mode at the current position (the whole mode when the
        position is not slab-sharded)."""
        lo, hi, d = C.c_int64(), C.c_int64(), C.c_int64()
        self.ctx.check(self._lib.nsb_shard_range(self.handle, C.byref(lo), C.byref(hi), C.byref(d)))
        return lo.value, hi.value, d.value

    def matvec_host_slab(self, x_slab):
        """H_eff application with the vector distributed over the ranks: this rank's slab in, the matching slab out."""
        a = np.asfortranarray(x_slab, dtype=self.dtype)
        out = np.empty(a.shape, dtype=self.dtype, order="F")
        self.ctx.check(self._lib.nsb_matvec_host_slab(self.handle, a.ctypes.data, out.ctypes.data))
        return out

    def matvec_device(self, reps=1, download=False):
        out = None
        if download:
            _, dims = self.local_info()
            out = np.empty(dims, dtype=self.dtype, order="F")
        self.ctx.check(self._lib.nsb_matvec_device(self.handle, reps, out.ctypes.data if out is not None else None))
        return out

    def range_finder(self, max_rank, oversample=2, north_pass=2, orthogonal_threshold=1e-12, cutoff=0.0, seed=1, probes=None):
        """range_finder(linear_map, random_vector) with linear_map = the projected operator at the current position
        (psi -> optimal_map(P, psi)); returns the basis as an array [local dims..., rank]."""
        _, dims = self.local_info()
        n = int(np.prod(dims))
        cap = min(max_rank + oversample, n)
        Q = np.empty((n, max(cap, 1)), dtype=self.dtype, order="F")
        rank = C.c_int64()
        pr = None
        if probes is not None:
            pr = np.asfortranarray(np.reshape(probes, (n, -1), order="F"), dtype=self.dtype)
            assert pr.shape[1] >= cap
        self.ctx.check(self._lib.nsb_range_finder_heff(self.handle, pr.ctypes.data if pr is not None else None, seed, max_rank,
                                                        oversample, north_pass, orthogonal_threshold, cutoff, Q.ctypes.data, C.byref(rank)))
        return Q[:, : rank.value].reshape(list(dims) + [rank.value], order="F")

    def set_expand_probe(self, probe):
        """One-shot random tensor (basis size x expand_space) for the next "ortho" expansion (instead of device Philox)."""
        a = np.asfortranarray(probe, dtype=self.dtype)
        self.ctx.check(self._lib.nsb_expand_set_probe(self.handle, a.shape[0], a.shape[1], a.ctypes.data))

    def shard_emulate(self, nranks):
        """H_eff theta with the arithmetic of an nranks-way partition on this one device (test hook); returns (theta', mode)."""
        _, dims = self.local_info()
        out = np.empty(dims, dtype=self.dtype, order="F")
        mode = C.c_int32()
        self.ctx.check(self._lib.nsb_shard_emulate(self.handle, int(nranks), out.ctypes.data, C.byref(mode)))
        return out, mode.value

    def matvec_flops(self):
        f = C.c_double()
        self.ctx.check(self._lib.nsb_matvec_flops(self.handle, C.byref(f)))
        return f.value

    def matvec_flops_executed(self):
        """Flops actually issued per H_eff application (dense count minus skipped identity channels)."""
        f = C.c_double()
        self.ctx.check(self._lib.nsb_matvec_flops_executed(self.handle, C.byref(f)))
        return f.value
