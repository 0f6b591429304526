"""Abelian quantum-number (QN) conservation for the oracle (test-only).

Restates the behaviour of the reference when it is run on QN-conserving ITensors (`siteinds(...; conserve_qns=true)`,
examples/dmrg.jl:10,18): block-sparse tensors whose factorisations never mix symmetry sectors, with the merged-spectrum
truncation of NDTensors (QN paragraph: global truncation, then every block keeps the values above
`docut`).

Storage stays dense (blocks forbidden by symmetry are exact zeros -- contractions of symmetric tensors keep them exact);
what QN conservation changes is only the three factorisations on the path (inserter `factorize`, expansion `eigen`,
gauge `qr`), which are done block by block.

Convention: every basis state of the link on edge {u, v} carries the charge of the subtree on the side of the vertex
recorded in `link_side`; the charge of the other side is `total - charge`.  This labelling does not depend on the
position of the orthogonality centre.  A site tensor psi[v] is non-zero only where
    q(site state) + sum_over_neighbours charge(subtree beyond that link) == total.
"""
from __future__ import annotations

import numpy as np

from .tensor import Tensor, edge_key, truncate_spectrum


class QNInfo:
    def __init__(self, total, site, link, link_side):
        self.total = np.asarray(total, dtype=np.int64)
        self.site = {v: np.asarray(a, dtype=np.int64) for v, a in site.items()}          # (d, nq)
        self.link = {k: np.asarray(a, dtype=np.int64) for k, a in link.items()}          # (dim, nq)
        self.link_side = dict(link_side)                                                 # edge_key -> vertex

    def copy(self):
        return QNInfo(self.total, self.site, self.link, self.link_side)

    def side_charge(self, v, n):
        """Charges (dim, nq) of the subtree on n's side of the edge {v, n}."""
        k = edge_key(v, n)
        a = self.link[k]
        return a if self.link_side[k] == n else self.total[None, :] - a

    def set_link(self, v, n, charges_on_v_side):
        k = edge_key(v, n)
        self.link[k] = np.asarray(charges_on_v_side, dtype=np.int64)
        self.link_side[k] = v


def label_charges(qn: QNInfo, owner, label):
    """Charge contribution (dim, nq) of one leg of the state tensor of vertex `owner`: the site charge, or the charge of
    the subtree hanging off that link (away from `owner`)."""
    kind, name, _ = label
    if kind == "s":
        return qn.site[name]
    assert kind == "l"
    u, v = name
    other = v if u == owner else u
    return qn.side_charge(owner, other)


def multi_index_charges(charge_list):
    """Charges of the C-order flattened multi-index (last label fastest)."""
    out = np.zeros((1, charge_list[0].shape[1]), dtype=np.int64) if charge_list else None
    for c in charge_list:
        out = (out[:, None, :] + c[None, :, :]).reshape(-1, c.shape[1])
    return out


def _blocks(row_keys, col_keys):
    rk = [tuple(r) for r in row_keys]
    ck = [tuple(c) for c in col_keys]
    rows, cols = {}, {}
    for i, k in enumerate(rk):
        rows.setdefault(k, []).append(i)
    for j, k in enumerate(ck):
        cols.setdefault(k, []).append(j)
    keys = [k for k in rows if k in cols]
    return [(k, np.array(rows[k]), np.array(cols[k])) for k in keys]


def check_block_structure(M, row_keys, col_keys, tol=1e-12):
    mask = (np.asarray(row_keys)[:, None, :] == np.asarray(col_keys)[None, :, :]).all(axis=2)
    off = np.abs(M[~mask]).max() if (~mask).any() else 0.0
    return off <= tol * max(np.abs(M).max(), 1e-300)


def truncate_merged(P_blocks, *, cutoff, mindim, maxdim):
    """NDTensors truncation of a block-sparse spectrum: merge, truncate globally (A.5), then each block keeps the values
    strictly above docut.  Returns (keep counts per block, truncerr)."""
    allP = np.sort(np.concatenate([p for p in P_blocks]) if P_blocks else np.zeros(0))[::-1]
    if len(allP) == 0:
        return [0 for _ in P_blocks], 0.0
    n, terr = truncate_spectrum(allP, cutoff=cutoff, mindim=mindim, maxdim=maxdim)
    docut = 0.0
    if n < len(allP):
        docut = (allP[n - 1] + allP[n]) / 2.0
        if abs(allP[n - 1] - allP[n]) < 1e-3 * allP[n - 1]:
            docut += 1e-3 * allP[n - 1]
    else:
        docut = -1.0
    keep = [int(np.count_nonzero(np.maximum(p, 0.0) > docut)) for p in P_blocks]
    if sum(keep) == 0:                       # never drop everything
        b = int(np.argmax([p.max() if len(p) else -1 for p in P_blocks]))
        keep[b] = 1
    return keep, terr


def svd_trunc_qn(M, row_keys, col_keys, *, cutoff, mindim, maxdim, use_eigen=False):
    """Block-wise truncated factorisation M = U S Vh.  Returns U (rows x k), s (k), Vh (k x cols), new_keys (k x nq), terr."""
    blks = _blocks(row_keys, col_keys)
    res = []
    for k, r, c in blks:
        B = M[np.ix_(r, c)]
        if use_eigen:
            w, V = np.linalg.eigh(0.5 * (B @ B.conj().T + (B @ B.conj().T).conj().T))
            order = np.argsort(-w, kind="stable")
            w, V = np.maximum(w[order], 0.0), V[:, order]
            kk = min(B.shape)
            U_b, P_b = V[:, :kk], w[:kk]
            R_b = U_b.conj().T @ B                     # = S Vh
            res.append((k, r, c, U_b, P_b, R_b))
        else:
            U_b, s_b, Vh_b = np.linalg.svd(B, full_matrices=False)
            res.append((k, r, c, U_b, s_b**2, s_b[:, None] * Vh_b))
    keep, terr = truncate_merged([x[4] for x in res], cutoff=cutoff, mindim=mindim, maxdim=maxdim)
    ktot = sum(keep)
    nq = np.asarray(row_keys).shape[1]
    U = np.zeros((M.shape[0], ktot), dtype=M.dtype)
    R = np.zeros((ktot, M.shape[1]), dtype=M.dtype)
    spec = np.zeros(ktot)
    new_keys = np.zeros((ktot, nq), dtype=np.int64)
    pos = 0
    for (k, r, c, U_b, P_b, R_b), nk in zip(res, keep):
        U[np.ix_(r, np.arange(pos, pos + nk))] = U_b[:, :nk]
        R[np.ix_(np.arange(pos, pos + nk), c)] = R_b[:nk]
        spec[pos:pos + nk] = P_b[:nk]
        new_keys[pos:pos + nk] = np.array(k)[None, :]
        pos += nk
    return U, spec, R, new_keys, terr


def eigen_trunc_qn(rho, keys, *, cutoff, mindim, maxdim):
    """Block-wise Hermitian eigendecomposition of rho (block diagonal in `keys`).  Returns D, U (n x k), new_keys, terr."""
    blks = _blocks(keys, keys)
    res = []
    for k, r, _ in blks:
        B = rho[np.ix_(r, r)]
        w, V = np.linalg.eigh(0.5 * (B + B.conj().T))
        order = np.argsort(-w, kind="stable")
        res.append((k, r, V[:, order], w[order]))
    keep, terr = truncate_merged([x[3] for x in res], cutoff=cutoff, mindim=mindim, maxdim=maxdim)
    ktot = sum(keep)
    nq = np.asarray(keys).shape[1]
    U = np.zeros((rho.shape[0], ktot), dtype=rho.dtype)
    D = np.zeros(ktot)
    new_keys = np.zeros((ktot, nq), dtype=np.int64)
    pos = 0
    for (k, r, V, w), nk in zip(res, keep):
        U[np.ix_(r, np.arange(pos, pos + nk))] = V[:, :nk]
        D[pos:pos + nk] = w[:nk]
        new_keys[pos:pos + nk] = np.array(k)[None, :]
        pos += nk
    return D, U, new_keys, terr


def qr_qn(M, row_keys, col_keys):
    """Block-wise thin QR.  Returns Q (rows x k), R (k x cols), new_keys (k x nq)."""
    blks = _blocks(row_keys, col_keys)
    parts = []
    for k, r, c in blks:
        Qb, Rb = np.linalg.qr(M[np.ix_(r, c)], mode="reduced")
        parts.append((k, r, c, Qb, Rb))
    ktot = sum(p[3].shape[1] for p in parts)
    nq = np.asarray(row_keys).shape[1]
    Q = np.zeros((M.shape[0], ktot), dtype=M.dtype)
    R = np.zeros((ktot, M.shape[1]), dtype=M.dtype)
    new_keys = np.zeros((ktot, nq), dtype=np.int64)
    pos = 0
    for k, r, c, Qb, Rb in parts:
        nk = Qb.shape[1]
        Q[np.ix_(r, np.arange(pos, pos + nk))] = Qb
        R[np.ix_(np.arange(pos, pos + nk), c)] = Rb
        new_keys[pos:pos + nk] = np.array(k)[None, :]
        pos += nk
    return Q, R, new_keys


def product_state_qn(g, site_charges, state_index, nq):
    """QNInfo of a product state: every link has one state whose charge is the sum of the site charges of a subtree."""
    from .graph import subtree_side
    total = np.zeros(nq, dtype=np.int64)
    for v in g.vertices:
        total += np.asarray(site_charges[v][state_index[v]], dtype=np.int64)
    link, side = {}, {}
    for (u, v) in g.edges:
        q = np.zeros(nq, dtype=np.int64)
        for x in subtree_side(g, u, v):
            q += np.asarray(site_charges[x][state_index[x]], dtype=np.int64)
        link[edge_key(u, v)] = q[None, :]
        side[edge_key(u, v)] = u
    return QNInfo(total, {v: np.asarray(site_charges[v]) for v in g.vertices}, link, side)


def check_state_symmetric(psi, tol=1e-12):
    """True if every site tensor obeys the selection rule of psi.qn."""
    qn = psi.qn
    for v in psi.graph.vertices:
        T = psi[v]
        ch = multi_index_charges([label_charges(qn, v, l) for l in T.labels])
        data = T.data.reshape(-1)
        bad = ~(ch == qn.total[None, :]).all(axis=1)
        if np.abs(data[bad]).max(initial=0.0) > tol * max(np.abs(data).max(), 1e-300):
            return False
    return True
