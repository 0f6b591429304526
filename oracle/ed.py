"""Exact diagonalisation helpers used to pin the oracle (test-only).  Independent of the
tensor-network code path: builds H directly from the OpSum as a sparse matrix.
Counterpart of test/utilities/simple_ed_methods.jl:4-42."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import scipy.linalg

from .models import _opmat
from .tensor import contract, site


def dense_hamiltonian(os, g, ops, sparse=True):
    verts = list(g.vertices)
    d = ops["Id"].shape[0]
    n = len(verts)
    pos = {v: i for i, v in enumerate(verts)}

    def embed(mats):
        out = None
        for i in range(n):
            m = sp.csr_matrix(mats.get(i, np.eye(d)))
            out = m if out is None else sp.kron(out, m, format="csr")
        return out

    H = None
    for (c, a, u, b, v) in os.two:
        term = c * embed({pos[u]: _opmat(ops, a), pos[v]: _opmat(ops, b)})
        H = term if H is None else H + term
    for (c, a, v) in os.one:
        term = c * embed({pos[v]: _opmat(ops, a)})
        H = term if H is None else H + term
    return H if sparse else H.toarray()


def ed_ground_state(os, g, ops, k=1):
    H = dense_hamiltonian(os, g, ops)
    if H.shape[0] <= 512:
        w, v = np.linalg.eigh(H.toarray())
        return w[0], v[:, 0]
    w, v = spla.eigsh(H, k=max(k, 2), which="SA", tol=1e-13)
    i = np.argmin(w)
    return w[i], v[:, i]


def state_vector(psi):
    """Contract a TTN into a dense vector, site order = graph.vertices order (first = slowest)."""
    g = psi.graph
    verts = list(g.vertices)
    # contract along a DFS order so intermediate tensors stay connected
    from .graph import dfs_parents
    order, _ = dfs_parents(g, verts[0])
    T = psi[order[0]]
    for v in order[1:]:
        T = contract(T, psi[v])
    return T.array([site(v) for v in verts]).ravel()


def ed_time_evolution(os, g, ops, psi0_vec, time_points, normalize=False):
    """test/utilities/simple_ed_methods.jl:18-42: psi(t) = exp(-i H dt) ... applied step by step."""
    H = dense_hamiltonian(os, g, ops, sparse=False)
    psi = psi0_vec.astype(complex)
    ex = [0.0] + [-1j * t for t in time_points]
    steps = [ex[i + 1] - ex[i] for i in range(len(ex) - 1)][1:]
    for s in steps:
        psi = scipy.linalg.expm(H * s) @ psi
        if normalize:
            psi = psi / np.linalg.norm(psi)
    return psi


def ttno_dense(H, g, d):
    """Contract a TTNO into a dense matrix (rows = out, cols = in), site order = graph.vertices."""
    from .graph import dfs_parents
    verts = list(g.vertices)
    order, _ = dfs_parents(g, verts[0])
    T = H[order[0]]
    for v in order[1:]:
        T = contract(T, H[v])
    n = len(verts)
    M = T.array([site(v, 1) for v in verts] + [site(v, 0) for v in verts])
    return M.reshape(d**n, d**n)
