"""Gauge moves (oracle; test-only).  UPSTREAM `itn.orthogonalize(psi, region)` called at
src/extracter.jl:6: QR walk over the Steiner tree joining the current
orthogonality region to the target, oriented toward first(region), skipping edges that lie
inside the region."""
from __future__ import annotations

from .graph import NamedGraph, tree_path
from .tensor import qr, contract, link

_TMP = ("x", "qr", 0)


def steiner_edges_toward(g: NamedGraph, terminals, target):
    """Edges (src->dst) of the Steiner tree of `terminals`, each oriented toward `target`,
    ordered so that every vertex's incoming edges precede its outgoing edge."""
    verts = set()
    for t in terminals:
        verts.update(tree_path(g, target, t))
    dist = {target: 0}
    frontier = [target]
    parent = {}
    while frontier:
        nxt = []
        for x in frontier:
            for n in g.neighbors(x):
                if n in verts and n not in dist:
                    dist[n] = dist[x] + 1
                    parent[n] = x
                    nxt.append(n)
        frontier = nxt
    order = sorted((v for v in verts if v != target), key=lambda v: -dist[v])
    return [(v, parent[v]) for v in order]


def qr_step(psi, a, b):
    """psi[a] -> Q (link label kept), psi[b] -> R * psi[b]."""
    A, B = psi[a], psi[b]
    l = link(a, b)
    left = [x for x in A.labels if x != l]
    if getattr(psi, "qn", None) is not None:
        import numpy as np
        from .qn import label_charges, multi_index_charges, qr_qn
        from .tensor import Tensor
        qn = psi.qn
        M = A.array(left + [l]).reshape(-1, A.dim(l))
        row_keys = multi_index_charges([label_charges(qn, a, x) for x in left])
        col_keys = qn.side_charge(b, a)
        Qm, Rm, newk = qr_qn(M, row_keys, col_keys)
        k = Qm.shape[1]
        Q = Tensor(Qm.reshape([A.dim(x) for x in left] + [k]), left + [_TMP])
        R = Tensor(Rm, [_TMP, l])
        qn.set_link(a, b, newk)
    else:
        Q, R = qr(A, left, _TMP)                 # Q: left + [TMP];  R: [TMP, l]
    psi[a] = Q.relabel({_TMP: l})
    RB = contract(R, B)                          # sums over l
    psi[b] = RB.relabel({_TMP: l}).permute(B.labels)


def orthogonalize(psi, region):
    region = list(region)
    if set(region) == set(psi.ortho_region):
        return psi
    g = psi.graph
    terminals = set(region) | set(psi.ortho_region)
    path = steiner_edges_toward(g, terminals, region[0])
    path = [(a, b) for (a, b) in path if not (a in region and b in region)]
    psi = psi.copy()
    for (a, b) in path:
        qr_step(psi, a, b)
    psi.ortho_region = region
    return psi
