"""Site types, tree tensor-network states and TTNO (MPO-on-a-tree) builders (oracle; test-only).

Stands in for the UPSTREAM constructors the reference examples call
(`itn.siteinds`, `itn.mpo(os, s)` / `itn.ttn(os, s)`, `itn.ttn(state, s)`, `itn.random_mps`;
examples/dmrg.jl:12-24,59-61).  The TTNO is the exact finite-state-machine form for sums of
one- and two-site terms on tree edges: operator-link dimension = 2 + (#terms on the edge),
i.e. w = 5 for Heisenberg, w = 3 for transverse-field Ising.
"""
from __future__ import annotations

import numpy as np

from .graph import NamedGraph, dfs_parents, default_root_vertex, subtree_side
from .tensor import Tensor, site, link, oplink, qr, contract

# ---------------------------------------------------------------------------
# local operators: matrices O[out, in]
# ---------------------------------------------------------------------------
def spin_ops(site_type):
    if site_type in ("S=1/2", "S=½"):
        sz = np.diag([0.5, -0.5])
        sp = np.array([[0.0, 1.0], [0.0, 0.0]])
        states = {"Up": 0, "Dn": 1}
    elif site_type == "S=1":
        sz = np.diag([1.0, 0.0, -1.0])
        sp = np.sqrt(2.0) * np.array([[0, 1, 0], [0, 0, 1], [0, 0, 0]], dtype=float)
        states = {"Up": 0, "Z0": 1, "Dn": 2}
    else:
        raise ValueError(site_type)
    d = sz.shape[0]
    sm = sp.T.copy()
    ops = {
        "Id": np.eye(d),
        "Sz": sz,
        "S+": sp,
        "S-": sm,
        "Sx": 0.5 * (sp + sm),
        "X": (sp + sm) if d == 2 else 0.5 * (sp + sm),
        "Z": 2.0 * sz if d == 2 else sz,
    }
    return d, ops, states


def electron_ops():
    """ITensors "Electron" site: basis Emp, Up, Dn, UpDn; Jordan-Wigner sign on Dn operators."""
    d = 4
    cup = np.zeros((4, 4)); cup[0, 1] = 1.0; cup[2, 3] = 1.0            # Cup
    cdn = np.zeros((4, 4)); cdn[0, 2] = 1.0; cdn[1, 3] = -1.0           # Cdn (sign from up)
    F = np.diag([1.0, -1.0, -1.0, 1.0])
    ops = {
        "Id": np.eye(4), "F": F,
        "Cup": cup, "Cdagup": cup.T.copy(), "Cdn": cdn, "Cdagdn": cdn.T.copy(),
        "Nup": cup.T @ cup, "Ndn": cdn.T @ cdn,
    }
    ops["Nupdn"] = ops["Nup"] @ ops["Ndn"]
    states = {"Emp": 0, "Up": 1, "Dn": 2, "UpDn": 3}
    return d, ops, states


class OpSum:
    """Sum of one-site terms (c, op, v) and nearest-neighbour two-site terms (c, opA, u, opB, v)."""

    def __init__(self):
        self.one = []
        self.two = []

    def add(self, coef, *args):
        if len(args) == 2:
            self.one.append((coef, args[0], args[1]))
        elif len(args) == 4:
            self.two.append((coef, args[0], args[1], args[2], args[3]))
        else:
            raise ValueError(args)
        return self


def heisenberg_opsum(g: NamedGraph):
    """examples/dmrg.jl:12-17, test/dmrg/test_tree_dmrg.jl:21-27."""
    os = OpSum()
    for (u, v) in g.edges:
        os.add(1.0, "Sz", u, "Sz", v)
        os.add(0.5, "S+", u, "S-", v)
        os.add(0.5, "S-", u, "S+", v)
    return os


def ising_opsum(g: NamedGraph, J=1.0, h=1.0):
    os = OpSum()
    for (u, v) in g.edges:
        os.add(-J, "Z", u, "Z", v)
    for v in g.vertices:
        os.add(-h, "X", v)
    return os


def hubbard_chain_opsum(g: NamedGraph, t=1.0, U=4.0):
    """Nearest-neighbour Hubbard chain with in-site Jordan-Wigner factors (vertices ordered u<v)."""
    os = OpSum()
    for (u, v) in g.edges:
        # c^dag_{u,up} c_{v,up}: (Cdagup F)_u (Cup)_v ; h.c.: (F Cup... ) handled via explicit products
        os.add(-t, "CdagupF", u, "Cup", v)
        os.add(-t, "CupF", u, "Cdagup", v)
        os.add(-t, "CdagdnF", u, "Cdn", v)
        os.add(-t, "CdnF", u, "Cdagdn", v)
    for v in g.vertices:
        os.add(U, "Nupdn", v)
    return os


def _opmat(ops, name):
    if name in ops:
        return ops[name]
    if name == "CdagupF":
        return ops["Cdagup"] @ ops["F"]
    if name == "CupF":
        return -(ops["Cup"] @ ops["F"])
    if name == "CdagdnF":
        return ops["Cdagdn"] @ ops["F"]
    if name == "CdnF":
        return -(ops["Cdn"] @ ops["F"])
    raise KeyError(name)


class TTN:
    """Tree tensor network state/operator container (stands in for itn.TreeTensorNetwork)."""

    def __init__(self, graph, tensors, ortho_region=None, qn=None):
        self.graph = graph
        self.tensors = dict(tensors)
        self.ortho_region = list(ortho_region) if ortho_region is not None else list(graph.vertices)
        self.qn = qn            # oracle.qn.QNInfo or None (dense, no symmetry bookkeeping)

    def __getitem__(self, v):
        return self.tensors[v]

    def __setitem__(self, v, t):
        self.tensors[v] = t

    def copy(self):
        return TTN(self.graph, dict(self.tensors), list(self.ortho_region), self.qn.copy() if self.qn is not None else None)

    def linkdim(self, u, v):
        return self.tensors[u].dim(link(u, v))

    def maxlinkdim(self):
        return max([self.linkdim(u, v) for (u, v) in self.graph.edges] or [1])

    def linkdims(self):
        return {(u, v): self.linkdim(u, v) for (u, v) in self.graph.edges}


def ttno(os: OpSum, g: NamedGraph, ops, root=None, dtype=float):
    """Exact FSM tree-tensor-network operator.  W[v] labels: oplinks (adjacency order), s_in, s_out."""
    if root is None:
        root = default_root_vertex(g)
    _, parent = dfs_parents(g, root)
    d = ops["Id"].shape[0]
    # terms per undirected edge, oriented (child op, parent op)
    edge_terms = {}
    for (c, a, u, b, v) in os.two:
        assert g.has_edge(u, v), f"two-site term on non-edge {(u, v)}"
        if parent.get(u) == v:      # u is the child
            edge_terms.setdefault((u, v), []).append((c, a, b))
        else:
            assert parent.get(v) == u
            edge_terms.setdefault((v, u), []).append((c, b, a))
    onsite = {}
    for (c, a, v) in os.one:
        onsite[v] = onsite.get(v, 0) + c * _opmat(ops, a)
    I_, F_ = 0, 1   # link states: 0 = nothing yet (identity below), 1 = finished; 2.. = pending term k

    def wdim(c):    # dimension of link (c -> parent(c))
        return 2 + len(edge_terms.get((c, parent[c]), []))

    tensors = {}
    for v in g.vertices:
        nbrs = g.neighbors(v)
        dims = [wdim(n) if parent.get(n) == v else wdim(v) for n in nbrs]
        W = np.zeros(dims + [d, d], dtype=dtype)   # [..., out, in] filled below then transposed
        children = [n for n in nbrs if parent.get(n) == v]
        has_up = parent[v] is not None
        pos = {n: i for i, n in enumerate(nbrs)}

        def idx(up_state, downs):
            ix = [0] * len(nbrs)
            for n in children:
                ix[pos[n]] = downs.get(n, I_)
            if has_up:
                ix[pos[parent[v]]] = up_state
            return tuple(ix)

        Id = ops["Id"]
        if has_up:
            W[idx(I_, {})] += Id
            for k, (c, a, b) in enumerate(edge_terms.get((v, parent[v]), [])):
                W[idx(2 + k, {})] += c * _opmat(ops, a)
        # "finished" rows (for the root there is no up link; idx ignores up_state)
        for n in children:
            W[idx(F_, {n: F_})] += Id
            for k, (c, a, b) in enumerate(edge_terms.get((n, v), [])):
                W[idx(F_, {n: 2 + k})] += _opmat(ops, b)
        if v in onsite:
            W[idx(F_, {})] += onsite[v]
        elif not children and not has_up:
            pass
        labels = [oplink(v, n) for n in nbrs] + [site(v, 1), site(v, 0)]
        T = Tensor(W, labels)
        tensors[v] = T.permute([oplink(v, n) for n in nbrs] + [site(v, 0), site(v, 1)])
    return TTN(g, tensors, ortho_region=[])


def product_ttn(g: NamedGraph, d, state_index, dtype=float):
    """`itn.ttn(state, sites)` then permute_indices layout (first link, site, other links)
    (src/permute_indices.jl:4-18)."""
    tensors = {}
    for v in g.vertices:
        nbrs = g.neighbors(v)
        vec = np.zeros(d, dtype=dtype)
        vec[state_index[v]] = 1.0
        labels = ([link(v, nbrs[0])] if nbrs else []) + [site(v)] + [link(v, n) for n in nbrs[1:]]
        shape = [1 if l[0] == "l" else d for l in labels]
        tensors[v] = Tensor(vec.reshape(shape), labels)
    return TTN(g, tensors)


def _bond_dims(g: NamedGraph, d, chi):
    dims = {}
    nv = len(g.vertices)
    for (u, v) in g.edges:
        nu = len(subtree_side(g, u, v))
        k = min(nu, nv - nu)
        cap = d ** min(k, 40)
        dims[(u, v)] = dims[(v, u)] = int(min(chi, cap))
    return dims


def random_ttn(g: NamedGraph, d, chi, seed=1234, dtype=float, root=None, orthogonalize_to=None):
    """Synthetic state: i.i.d. N(0,1) tensors with uniform bond chi (capped by d^k), QR-gauged to
    `orthogonalize_to` (default: root).  Analogue of `itn.random_mps(s; link_space=chi)`
    (examples/dmrg.jl:61); SURVEY.md 8(d) config 2."""
    rng = np.random.default_rng(seed)
    dims = _bond_dims(g, d, chi)
    tensors = {}
    for v in g.vertices:
        nbrs = g.neighbors(v)
        labels = ([link(v, nbrs[0])] if nbrs else []) + [site(v)] + [link(v, n) for n in nbrs[1:]]
        shape = [d if l[0] == "s" else dims[(v, [n for n in nbrs if link(v, n) == l][0])] for l in labels]
        data = rng.standard_normal(shape)
        if np.issubdtype(np.dtype(dtype), np.complexfloating):
            data = data + 1j * rng.standard_normal(shape)
        tensors[v] = Tensor(data.astype(dtype), labels)
    psi = TTN(g, tensors)
    from .gauge import orthogonalize

    if root is None:
        root = default_root_vertex(g)
    tgt = [root] if orthogonalize_to is None else list(orthogonalize_to)
    psi = orthogonalize(psi, tgt)
    c = tgt[0]
    psi[c] = psi[c] / psi[c].norm()
    return psi


def site_charges(site_type):
    """Per-basis-state charges: S=1/2 -> (2 Sz); S=1 -> (2 Sz); Electron -> (Nf, 2 Sz)."""
    if site_type in ("S=1/2", "S=½"):
        return np.array([[1], [-1]])
    if site_type == "S=1":
        return np.array([[2], [0], [-2]])
    if site_type == "Electron":
        return np.array([[0, 0], [1, 1], [1, -1], [2, 0]])
    raise ValueError(site_type)


def product_ttn_qn(g: NamedGraph, site_type, d, state_index, dtype=float):
    """Product state with QN bookkeeping (`siteinds(...; conserve_qns=true)` + `ttn(state, sites)`)."""
    from .qn import product_state_qn
    psi = product_ttn(g, d, state_index, dtype=dtype)
    sc = site_charges(site_type)
    psi.qn = product_state_qn(g, {v: sc for v in g.vertices}, state_index, sc.shape[1])
    return psi
