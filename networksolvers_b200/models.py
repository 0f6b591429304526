"""Host-side tensor networks: site types, product / random states and the tree-tensor-network operator
of a sum of one- and two-site terms.  These stand in for the ITensor constructors the reference's callers
use before they reach the solver entry points (`siteinds`, `ttn(state, sites)`, `random_mps`,
`mpo(opsum, sites)` / `ttn(opsum, sites)`; examples/dmrg.jl:12-24,59-61).  Everything here is small
host data handed to the device library through `DeviceNetwork`.

A host tensor is a numpy array plus a list of legs, one per axis:
    ("site", v)            physical index of vertex v (ket / operator input)
    ("site_out", v)        primed physical index (operator output)
    ("link", v, n)         bond on the tree edge {v, n}
"""
from __future__ import annotations

import numpy as np

from .graphs import NamedGraph, default_root_vertex, _dfs


class SiteType:
    def __init__(self, name):
        self.name = name
        if name in ("S=1/2", "S=½"):
            sz = np.diag([0.5, -0.5])
            sp = np.array([[0.0, 1.0], [0.0, 0.0]])
            self.states = {"Up": 0, "Dn": 1}
        elif name == "S=1":
            sz = np.diag([1.0, 0.0, -1.0])
            sp = np.sqrt(2.0) * np.diag([1.0, 1.0], k=1)
            self.states = {"Up": 0, "Z0": 1, "Dn": 2}
        elif name == "Electron":
            # basis Emp, Up, Dn, UpDn; the in-site Jordan-Wigner sign sits in the Dn operators (ITensors convention)
            cup = np.zeros((4, 4)); cup[0, 1] = 1.0; cup[2, 3] = 1.0
            cdn = np.zeros((4, 4)); cdn[0, 2] = 1.0; cdn[1, 3] = -1.0
            F = np.diag([1.0, -1.0, -1.0, 1.0])
            self.dim = 4
            self.states = {"Emp": 0, "Up": 1, "Dn": 2, "UpDn": 3}
            self.ops = {"Id": np.eye(4), "F": F, "Cup": cup, "Cdagup": cup.T.copy(), "Cdn": cdn, "Cdagdn": cdn.T.copy(),
                        "Nup": cup.T @ cup, "Ndn": cdn.T @ cdn, "Nupdn": (cup.T @ cup) @ (cdn.T @ cdn),
                        "CdagupF": cup.T @ F, "CupF": -(cup @ F), "CdagdnF": cdn.T @ F, "CdnF": -(cdn @ F)}
            self.ops["Ntot"] = self.ops["Nup"] + self.ops["Ndn"]
            self.ops["Sz"] = 0.5 * (self.ops["Nup"] - self.ops["Ndn"])
            return
        else:
            raise ValueError(f"unknown site type {name}")
        sm = sp.T.copy()
        self.dim = sz.shape[0]
        self.ops = {"Id": np.eye(self.dim), "Sz": sz, "S+": sp, "S-": sm, "Sx": (sp + sm) / 2,
                    "X": (sp + sm) if self.dim == 2 else (sp + sm) / 2, "Z": 2 * sz if self.dim == 2 else sz}

    def op(self, name):
        return self.ops[name]


class GraphSites:
    """Site dimensions of an existing network (graph + uniform physical dimension), for constructors that only need
    `.graph` and `.dim`."""

    def __init__(self, graph, dim):
        self.graph, self.dim = graph, int(dim)

    @classmethod
    def of(cls, tn):
        v = tn.graph.vertices[0]
        return cls(tn.graph, tn.tensors[v].shape[tn.legs[v].index(("site", v))])


class SiteSet:
    """`siteinds(site_type, graph)`."""

    def __init__(self, site_type, graph, conserve_qns=False):
        self.graph = graph
        self.type = site_type if isinstance(site_type, SiteType) else SiteType(site_type)
        self.dim = self.type.dim
        self.conserve_qns = conserve_qns


def siteinds(site_type, graph_or_n, conserve_qns=False):
    from .graphs import path_graph
    g = path_graph(graph_or_n) if isinstance(graph_or_n, int) else graph_or_n
    return SiteSet(site_type, g, conserve_qns=conserve_qns)


class OpSum:
    def __init__(self):
        self.terms = []

    def add(self, coef, *ops_and_sites):
        """os.add(c, "Sz", i) / os.add(c, "S+", i, "S-", j) / ... any number of (operator name, vertex) pairs."""
        assert len(ops_and_sites) >= 2 and len(ops_and_sites) % 2 == 0
        self.terms.append((coef,) + tuple(ops_and_sites))
        return self

    __iadd__ = lambda self, t: self.add(*t)  # os += (c, "Sz", i, "Sz", j)


def heisenberg(graph):
    os = OpSum()
    for u, v in graph.edges:
        os.add(1.0, "Sz", u, "Sz", v)
        os.add(0.5, "S+", u, "S-", v)
        os.add(0.5, "S-", u, "S+", v)
    return os


def transverse_ising(graph, J=1.0, h=1.0):
    os = OpSum()
    for u, v in graph.edges:
        os.add(-J, "Z", u, "Z", v)
    for v in graph.vertices:
        os.add(-h, "X", v)
    return os


def hubbard(graph, t=1.0, U=4.0):
    """Nearest-neighbour Hubbard model on a chain-ordered tree edge list (u before v in Jordan-Wigner order):
    -t sum (c^dag_{u s} c_{v s} + h.c.) + U sum n_up n_dn, written with in-site string factors so that every term is a
    product of two local operators (MPO bond dimension 6)."""
    os = OpSum()
    for u, v in graph.edges:
        os.add(-t, "CdagupF", u, "Cup", v)
        os.add(-t, "CupF", u, "Cdagup", v)
        os.add(-t, "CdagdnF", u, "Cdn", v)
        os.add(-t, "CdnF", u, "Cdagdn", v)
    for v in graph.vertices:
        os.add(U, "Nupdn", v)
    return os


class HostTTN:
    """Tensors on the vertices of a tree: `tensors[v]` (numpy, C-contiguous logical layout) with `legs[v]`."""

    def __init__(self, graph, tensors, legs, ortho_region=None, site_dim=None, qn=None):
        self.qn = qn          # None, or dict(total=(nq,), site={v: (d, nq)}, link={(u, v): (dim, nq) on u's side})
        self.graph = graph
        self.tensors = dict(tensors)
        self.legs = {v: list(l) for v, l in legs.items()}
        self.ortho_region = list(ortho_region) if ortho_region is not None else list(graph.vertices)
        self.site_dim = site_dim

    def __getitem__(self, v):
        return self.tensors[v]

    def linkdim(self, u, v):
        return self.tensors[u].shape[self.legs[u].index(("link", u, v))]

    def maxlinkdim(self):
        return max([self.linkdim(u, v) for u, v in self.graph.edges] or [1])

    def dtype(self):
        return np.result_type(*[t.dtype for t in self.tensors.values()])

    def to_dense(self):
        """Contract to a dense vector (small networks only); site order = graph.vertices."""
        verts = self.graph.vertices
        post, parent = _dfs(self.graph, verts[0])
        letters = {}

        def sym(leg):
            key = leg if leg[0] != "link" else ("link",) + tuple(sorted(leg[1:], key=repr))
            return letters.setdefault(key, len(letters))

        operands = []
        for v in verts:
            operands += [self.tensors[v], [sym(l) for l in self.legs[v]]]
        out = [sym(("site", v)) for v in verts]
        return np.einsum(*operands, out, optimize="greedy").reshape(-1)


def canonical_legs(graph, v):
    """(first link, site, other links) -- the layout `permute_indices` produces (src/permute_indices.jl:4-18)."""
    nb = graph.neighbors(v)
    return ([("link", v, nb[0])] if nb else []) + [("site", v)] + [("link", v, n) for n in nb[1:]]


SITE_CHARGES = {"S=1/2": [[1], [-1]], "S=½": [[1], [-1]], "S=1": [[2], [0], [-2]],
                "Electron": [[0, 0], [1, 1], [1, -1], [2, 0]]}      # (2 Sz) for spins, (Nf, 2 Sz) for electrons


def _side_vertices(g, u, v):
    seen, todo = {u}, [u]
    while todo:
        x = todo.pop()
        for n in g.neighbors(x):
            if n not in seen and not (x == u and n == v):
                seen.add(n)
                todo.append(n)
    return seen


def product_state(sites: SiteSet, state, dtype=float, conserve_qns=None):
    """`ttn(state, sites)`: state maps vertex -> state name (e.g. "Up") or basis index.  With conserve_qns (or
    `siteinds(...; conserve_qns=True)`) the state carries its abelian charges and every later factorisation is done
    sector by sector."""
    g = sites.graph
    conserve = sites.conserve_qns if conserve_qns is None else conserve_qns
    tensors, legs = {}, {}
    chosen = {}
    for v in g.vertices:
        s = state[v] if not callable(state) else state(v)
        idx = sites.type.states[s] if isinstance(s, str) else int(s)
        chosen[v] = idx
        lg = canonical_legs(g, v)
        arr = np.zeros([sites.dim if l[0] == "site" else 1 for l in lg], dtype=dtype)
        arr.reshape(-1)[idx] = 1.0
        tensors[v], legs[v] = arr, lg
    qn = None
    if conserve:
        sc = np.array(SITE_CHARGES[sites.type.name], dtype=np.int64)
        total = sum(sc[chosen[v]] for v in g.vertices)
        link = {(u, v): sum(sc[chosen[x]] for x in _side_vertices(g, u, v))[None, :] for u, v in g.edges}
        qn = dict(total=np.asarray(total), site={v: sc for v in g.vertices}, link=link)
    return HostTTN(g, tensors, legs, site_dim=sites.dim, qn=qn)


def _side_size(g, u, v):
    seen, todo = {u}, [u]
    while todo:
        x = todo.pop()
        for n in g.neighbors(x):
            if n not in seen and not (x == u and n == v):
                seen.add(n)
                todo.append(n)
    return len(seen)


def bond_dims(graph, d, chi):
    nv = len(graph.vertices)
    out = {}
    for u, v in graph.edges:
        k = _side_size(graph, u, v)
        k = min(k, nv - k)
        out[(u, v)] = out[(v, u)] = int(min(chi, d ** min(k, 40)))
    return out


def random_state(sites: SiteSet, link_space, seed=1234, dtype=float):
    """`random_mps(sites; link_space)`-like synthetic state: i.i.d. N(0,1) entries scaled by 1/sqrt(size);
    the gauge is left to the first `extracter` call (orthogonality region = all vertices)."""
    g = sites.graph
    rng = np.random.default_rng(seed)
    dims = bond_dims(g, sites.dim, link_space)
    tensors, legs = {}, {}
    for v in g.vertices:
        lg = canonical_legs(g, v)
        shape = [sites.dim if l[0] == "site" else dims[(l[1], l[2])] for l in lg]
        arr = rng.standard_normal(shape)
        if np.issubdtype(np.dtype(dtype), np.complexfloating):
            arr = arr + 1j * rng.standard_normal(shape)
        tensors[v] = (arr / np.sqrt(arr.size / shape[0] if len(shape) > 1 else 1.0)).astype(dtype)
        legs[v] = lg
    return HostTTN(g, tensors, legs, site_dim=sites.dim)


def _is_local_opsum(opsum: OpSum, g) -> bool:
    for term in opsum.terms:
        if len(term) == 3:
            continue
        if len(term) == 5 and term[2] != term[4] and g.has_edge(term[2], term[4]):
            continue
        return False
    return True


def ttno(opsum: OpSum, sites: SiteSet, root=None, dtype=float, cutoff=1e-14):
    """Tree tensor network operator of an OpSum (`itn.ttn(opsum, sites)` / `itn.mpo`, examples/dmrg.jl:18,59).
    One-site and nearest-neighbour two-site terms: the exact finite-state-machine construction below.  Anything else
    (long-range or multi-site terms): sum of product operators compressed by SVD (`ttno_general`)."""
    if not _is_local_opsum(opsum, sites.graph):
        return ttno_general(opsum, sites, dtype=dtype, cutoff=cutoff)
    return _ttno_local(opsum, sites, root=root, dtype=dtype)


def _ttno_local(opsum: OpSum, sites: SiteSet, root=None, dtype=float):
    """Exact finite-state-machine TTNO for one-site and nearest-neighbour two-site terms.  Operator link
    states toward the root: 0 = identity so far, 1 = a complete term lies below, 2+k = term k of that edge
    started below.  Link dimension 2 + (#terms on the edge): 5 for Heisenberg, 3 for Ising."""
    g = sites.graph
    root = default_root_vertex(g) if root is None else root
    _, parent = _dfs(g, root)
    d, op = sites.dim, sites.type.op
    on_edge, on_site = {}, {}
    for term in opsum.terms:
        if len(term) == 3:
            c, a, v = term
            on_site[v] = on_site.get(v, 0) + c * op(a)
        else:
            c, a, u, b, v = term
            if not g.has_edge(u, v):
                raise ValueError(f"two-site term on {(u, v)} which is not an edge of the tree")
            if parent.get(u) == v:
                on_edge.setdefault(u, []).append((c, a, b))      # keyed by the child vertex
            else:
                on_edge.setdefault(v, []).append((c, b, a))
    wdim = lambda child: 2 + len(on_edge.get(child, []))
    tensors, legs = {}, {}
    for v in g.vertices:
        nb = g.neighbors(v)
        kids = [n for n in nb if parent.get(n) == v]
        up = parent[v]
        shape = [wdim(n) if n in kids else wdim(v) for n in nb]
        Wt = np.zeros(shape + [d, d], dtype=dtype)                # [..., out, in]

        def at(up_state, kid_states):
            return tuple((kid_states.get(n, 0) if n in kids else up_state) for n in nb)

        if up is not None:
            Wt[at(0, {})] += np.eye(d)
            for k, (c, a, b) in enumerate(on_edge.get(v, [])):
                Wt[at(2 + k, {})] += c * op(a)
        for n in kids:
            Wt[at(1, {n: 1})] += np.eye(d)
            for k, (c, a, b) in enumerate(on_edge.get(n, [])):
                Wt[at(1, {n: 2 + k})] += op(b)
        if v in on_site:
            Wt[at(1, {})] += on_site[v]
        tensors[v] = np.swapaxes(Wt, -1, -2).copy()               # [..., in, out]
        legs[v] = [("link", v, n) for n in nb] + [("site", v), ("site_out", v)]
    return HostTTN(g, tensors, legs, ortho_region=[], site_dim=d)


mpo = ttno


# ---- host-side observables for callbacks (`itn.expect(state(problem), "Sz", v)`, examples/quench_evolution.jl:41-47) ------
def _link_axis(psi: HostTTN, u, n):
    for i, l in enumerate(psi.legs[u]):
        if l[0] == "link" and set(l[1:]) == {u, n}:
            return i
    raise KeyError((u, n))


def _subtree_env(psi: HostTTN, u, p, cache):
    """E[bra, ket] of the subtree hanging off vertex u, seen through u's link toward p (<psi|psi> with that link open)."""
    if (u, p) in cache:
        return cache[(u, p)]
    T = psi.tensors[u]
    X = T
    for c in psi.graph.neighbors(u):
        if c == p:
            continue
        ax = _link_axis(psi, u, c)
        Ec = _subtree_env(psi, c, u, cache)                       # [bra, ket]
        X = np.moveaxis(np.tensordot(Ec, X, axes=(1, ax)), 0, ax)   # ket link -> bra link
    axp = _link_axis(psi, u, p)
    other = [i for i in range(T.ndim) if i != axp]
    E = np.tensordot(np.conj(T), X, axes=(other, other))         # [bra link, ket link]
    cache[(u, p)] = E
    return E


def _pair_env(a: HostTTN, b: HostTTN, u, p, cache):
    """E[link of a (bra), link of b (ket)] of the subtree at u seen through the link toward p, for <a|b>."""
    if (u, p) in cache:
        return cache[(u, p)]
    X = b.tensors[u]
    # bring b's tensor to a's leg order (same graph, same canonical legs in practice)
    if b.legs[u] != a.legs[u]:
        X = np.transpose(X, [b.legs[u].index(l) for l in a.legs[u]])
    for c in a.graph.neighbors(u):
        if c == p:
            continue
        ax = _link_axis(a, u, c)
        X = np.moveaxis(np.tensordot(_pair_env(a, b, c, u, cache), X, axes=(1, ax)), 0, ax)
    T = a.tensors[u]
    if p is None:
        return np.vdot(T, X)
    axp = _link_axis(a, u, p)
    other = [i for i in range(T.ndim) if i != axp]
    E = np.tensordot(np.conj(T), X, axes=(other, other))
    cache[(u, p)] = E
    return E


def inner(a: HostTTN, b: HostTTN):
    """<a|b> of two tree tensor network states on the same graph (`itn.inner(a, b; alg="exact")`), one pass over the tree."""
    return _pair_env(a, b, a.graph.vertices[0], None, {})


def expect(psi: HostTTN, op, v, sites: SiteSet = None):
    """<psi| op_v |psi> / <psi|psi> for a tree tensor network state on the host (`op`: operator name looked up in `sites`, or a
    d x d matrix <out|op|in>).  One pass over the tree, O(chi^3) per vertex."""
    M = np.asarray(sites.type.op(op) if isinstance(op, str) else op)
    cache = {}
    T = psi.tensors[v]
    X = T
    for c in psi.graph.neighbors(v):
        ax = _link_axis(psi, v, c)
        X = np.moveaxis(np.tensordot(_subtree_env(psi, c, v, cache), X, axes=(1, ax)), 0, ax)
    sax = psi.legs[v].index(("site", v))
    OX = np.moveaxis(np.tensordot(M, X, axes=(1, sax)), 0, sax)
    num = np.vdot(T, OX)
    den = np.vdot(T, X)
    val = num / den
    return float(val.real) if abs(val.imag) <= 1e-12 * max(1.0, abs(val)) else complex(val)


# ---- general OpSum -> compressed TTNO (operator construction with compression) ------------------
def _site_first(H: HostTTN) -> HostTTN:
    """Same operator with every tensor laid out [site, site_out, links in neighbour order]."""
    g = H.graph
    tensors, legs = {}, {}
    for v in g.vertices:
        want = [("site", v), ("site_out", v)] + [("link", v, n) for n in g.neighbors(v)]
        perm = [H.legs[v].index(l) for l in want]
        tensors[v] = np.ascontiguousarray(np.transpose(H.tensors[v], perm))
        legs[v] = want
    return HostTTN(g, tensors, legs, ortho_region=[], site_dim=H.site_dim)


def operator_direct_sum(A: HostTTN, B: HostTTN) -> HostTTN:
    """A + B as operators: block-diagonal operator links (link dimensions add), shared site legs."""
    A, B = _site_first(A), _site_first(B)
    g = A.graph
    if len(g.vertices) == 1:
        v = g.vertices[0]
        return HostTTN(g, {v: A.tensors[v] + B.tensors[v]}, A.legs, ortho_region=[], site_dim=A.site_dim)
    tensors = {}
    for v in g.vertices:
        ta, tb = A.tensors[v], B.tensors[v]
        la, lb = ta.shape[2:], tb.shape[2:]
        t = np.zeros(ta.shape[:2] + tuple(x + y for x, y in zip(la, lb)), dtype=np.result_type(ta, tb))
        t[(slice(None), slice(None)) + tuple(slice(0, x) for x in la)] = ta
        t[(slice(None), slice(None)) + tuple(slice(x, x + y) for x, y in zip(la, lb))] = tb
        tensors[v] = t
    return HostTTN(g, tensors, A.legs, ortho_region=[], site_dim=A.site_dim)


def compress_operator(H: HostTTN, cutoff=1e-14, maxdim=None) -> HostTTN:
    """Exact-to-`cutoff` compression of the operator links of a tree tensor network operator: QR sweep toward the root
    (every tensor orthonormal toward it), then an Euler tour from the root that truncates each edge by SVD with the centre
    on it (discarded squared singular values / total <= cutoff, at most maxdim kept) and returns by QR.  The operator is
    treated as a state with site dimension d^2 (Frobenius norm), as ITensor does for `MPO(opsum)`."""
    H = _site_first(H)
    g = H.graph
    if len(g.vertices) == 1:
        return H
    T = {v: np.array(t) for v, t in H.tensors.items()}
    root = g.vertices[0]
    post, parent = _dfs(g, root)

    def axis(v, n):
        return 2 + g.neighbors(v).index(n)

    def split_toward(v, n, truncate):
        """T[v] = (isometry toward n) * R; R is absorbed by n.  Centre moves v -> n."""
        ax = axis(v, n)
        t = np.moveaxis(T[v], ax, -1)
        shp = t.shape
        M = t.reshape(-1, shp[-1])
        if truncate:
            U, S, Vh = np.linalg.svd(M, full_matrices=False)
            p = S ** 2
            tot = p.sum()
            k = len(S)
            disc = 0.0
            while k > 1 and (disc + p[k - 1] <= cutoff * tot or (maxdim is not None and k > maxdim)):
                disc += p[k - 1]
                k -= 1
            Q, R = U[:, :k], S[:k, None] * Vh[:k]
        else:
            Q, R = np.linalg.qr(M)
        T[v] = np.moveaxis(Q.reshape(shp[:-1] + (Q.shape[1],)), -1, ax)
        an = axis(n, v)
        T[n] = np.moveaxis(np.tensordot(R, np.moveaxis(T[n], an, 0), axes=(1, 0)), 0, an)

    for v in post:                      # children before parents
        if parent[v] is not None:
            split_toward(v, parent[v], truncate=False)

    def tour(v):
        for c in g.neighbors(v):
            if parent.get(c) == v:
                split_toward(v, c, truncate=True)
                tour(c)
                split_toward(c, v, truncate=False)

    import sys
    lim = sys.getrecursionlimit()
    sys.setrecursionlimit(max(lim, 4 * len(g.vertices) + 100))
    try:
        tour(root)
    finally:
        sys.setrecursionlimit(lim)
    return HostTTN(g, T, H.legs, ortho_region=[], site_dim=H.site_dim)


def ttno_general(opsum: OpSum, sites: SiteSet, dtype=float, cutoff=1e-14, batch=24) -> HostTTN:
    """OpSum with arbitrary supports -> compressed TTNO: the terms enter `batch` at a time as a sum of product operators
    (one operator-link state per term), each partial sum is compressed before the next batch is added, so the link dimension
    never exceeds (compressed dimension + batch).  Operators repeated on one vertex are multiplied in the order written.
    Fermionic strings are the caller's business (as in `hubbard` above)."""
    op, d = sites.type.op, sites.dim
    terms = []
    for term in opsum.terms:
        c, rest = term[0], term[1:]
        ops = {}
        for name, v in zip(rest[0::2], rest[1::2]):
            m = np.asarray(op(name))
            ops[v] = m if v not in ops else ops[v] @ m
        terms.append((c, ops))
    if not terms:
        raise ValueError("empty OpSum")
    cplx = any(np.iscomplexobj(m) for _, ops in terms for m in ops.values()) or any(np.iscomplexobj(c) for c, _ in terms)
    dt = complex if (cplx or np.dtype(dtype).kind == "c") else dtype
    acc = None
    for i in range(0, len(terms), batch):
        part = _product_matrix_sum(sites, terms[i:i + batch], dt)
        acc = part if acc is None else operator_direct_sum(acc, part)
        acc = compress_operator(acc, cutoff=cutoff)
    # layout of ttno(): [links in neighbour order, site, site_out]
    g = sites.graph
    tensors, legs = {}, {}
    for v in g.vertices:
        nl = len(g.neighbors(v))
        tensors[v] = np.ascontiguousarray(np.moveaxis(acc.tensors[v], (0, 1), (nl, nl + 1)))
        legs[v] = [("link", v, n) for n in g.neighbors(v)] + [("site", v), ("site_out", v)]
    return HostTTN(g, tensors, legs, ortho_region=[], site_dim=d)


def _product_matrix_sum(sites: SiteSet, terms, dtype):
    """sum_k c_k prod_v M_k(v) with explicit matrices: one operator-link state per term (cf. product_operator_sum)."""
    g = sites.graph
    d = sites.dim
    nt = len(terms)
    root = g.vertices[0]
    tensors, legs = {}, {}
    for v in g.vertices:
        nb = g.neighbors(v)
        legs[v] = [("site", v), ("site_out", v)] + [("link", v, n) for n in nb]
        t = np.zeros([d, d] + [nt] * len(nb), dtype=dtype)
        for k, (c, ops) in enumerate(terms):
            m = ops[v] if v in ops else np.eye(d)
            if v == root:
                m = c * m
            if nb:
                t[(slice(None), slice(None)) + (k,) * len(nb)] = np.asarray(m, dtype=dtype).T      # (in, out): <out| m |in>
            else:
                t += np.asarray(m, dtype=dtype).T
        tensors[v] = t
    return HostTTN(g, tensors, legs, ortho_region=[], site_dim=d)


# ---- fitting helpers (src/fitting.jl:90-112) -----------------------------------------------------------------
def identity_operator(sites: SiteSet, dtype=float):
    """Operator network acting as the identity, every operator link of dimension 1 (the overlap network <psi|x> of
    `itn.truncate` has no operator layer)."""
    g = sites.graph
    d = sites.dim
    tensors, legs = {}, {}
    for v in g.vertices:
        nb = g.neighbors(v)
        legs[v] = [("site", v), ("site_out", v)] + [("link", v, n) for n in nb]
        tensors[v] = np.eye(d, dtype=dtype).reshape([d, d] + [1] * len(nb))
    return HostTTN(g, tensors, legs, ortho_region=[], site_dim=d)


def delta_state(sites: SiteSet, link_space, dtype=float):
    """`ITensorNetwork(v -> inds -> delta(inds), siteinds; link_space)` (src/fitting.jl:91-93, :106-108)."""
    g = sites.graph
    d = sites.dim
    tensors, legs = {}, {}
    for v in g.vertices:
        lg = canonical_legs(g, v)
        shape = [d if l[0] == "site" else int(link_space) for l in lg]
        t = np.zeros(shape, dtype=dtype)
        for i in range(min(shape)):
            t[(i,) * len(shape)] = 1.0
        tensors[v], legs[v] = t, lg
    return HostTTN(g, tensors, legs, ortho_region=list(g.vertices), site_dim=d)


def random_tensornetwork(sites: SiteSet, link_space, rng, dtype=float):
    """`itn.random_tensornetwork(rng, elt, s; link_space)`: i.i.d. normal entries, uniform link dimension, no gauge."""
    g = sites.graph
    d = sites.dim
    tensors, legs = {}, {}
    for v in g.vertices:
        lg = canonical_legs(g, v)
        shape = [d if l[0] == "site" else int(link_space) for l in lg]
        t = rng.standard_normal(shape)
        if np.issubdtype(np.dtype(dtype), np.complexfloating):
            t = t + 1j * rng.standard_normal(shape)
        tensors[v], legs[v] = t.astype(dtype), lg
    return HostTTN(g, tensors, legs, ortho_region=list(g.vertices), site_dim=d)


def product_operator_sum(sites: SiteSet, terms, dtype=float):
    """Operator network of sum_k c_k prod_v op_k(v) for arbitrary (not only neighbouring) supports: every operator link
    has one state per term.  `terms` = [(coef, {vertex: op name, ...}), ...].  Used where the reference builds
    `itn.ttn(opsum, sites)` with long-range terms (test/fitting/fitting_regression_test.jl:47-50)."""
    g = sites.graph
    d, op = sites.dim, sites.type.op
    nt = len(terms)
    root = g.vertices[0]
    tensors, legs = {}, {}
    for v in g.vertices:
        nb = g.neighbors(v)
        legs[v] = [("site", v), ("site_out", v)] + [("link", v, n) for n in nb]
        t = np.zeros([d, d] + [nt] * len(nb), dtype=dtype)
        for k, (c, ops) in enumerate(terms):
            m = op(ops[v]) if v in ops else np.eye(d)
            if v == root:
                m = c * m
            # site legs are (in, out): <out| m |in>
            t[(slice(None), slice(None)) + (k,) * len(nb)] = np.asarray(m, dtype=dtype).T
        tensors[v] = t
    return HostTTN(g, tensors, legs, ortho_region=[], site_dim=d)
