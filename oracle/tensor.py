"""Labelled dense tensors with ITensor-like contraction semantics (oracle; test-only).

An index label is a tuple ``(kind, name, plev)``:
  kind 's'  site index of vertex `name`
  kind 'l'  state (link) index on the undirected edge `name` = edge_key(u, v)
  kind 'm'  operator (MPO/TTNO) link on edge `name`
  kind 'x'  auxiliary index (e.g. the new index of an eigen decomposition)
`plev` is the prime level.  `A * B` contracts over all labels the operands share,
like `*` on ITensors (used throughout the reference, e.g. src/operator_map.jl:20-37).

Linear-algebra rules restated here are UPSTREAM (ITensors/NDTensors), SURVEY.md App. A.4,
A.5, A.8, A.9.
"""
from __future__ import annotations

import numpy as np


def edge_key(u, v):
    """Canonical undirected-edge name."""
    a, b = sorted((u, v), key=repr)
    return (a, b)


def site(v, p=0):
    return ("s", v, p)


def link(u, v, p=0):
    return ("l", edge_key(u, v), p)


def oplink(u, v, p=0):
    return ("m", edge_key(u, v), p)


class Tensor:
    __slots__ = ("data", "labels")

    def __init__(self, data, labels):
        data = np.asarray(data)
        labels = tuple(labels)
        assert data.ndim == len(labels), (data.shape, labels)
        assert len(set(labels)) == len(labels), labels
        self.data = data
        self.labels = labels

    # -- basic queries -------------------------------------------------
    @property
    def shape(self):
        return self.data.shape

    def dim(self, label):
        return self.data.shape[self.labels.index(label)]

    def copy(self):
        return Tensor(self.data.copy(), self.labels)

    def norm(self):
        return float(np.linalg.norm(self.data.ravel()))

    def scalar(self):
        assert self.data.size == 1
        return self.data.reshape(())[()]

    def permute(self, labels):
        labels = tuple(labels)
        perm = [self.labels.index(l) for l in labels]
        return Tensor(np.transpose(self.data, perm), labels)

    def array(self, labels):
        """Dense array with axes in the requested label order."""
        return np.ascontiguousarray(self.permute(labels).data)

    def relabel(self, mapping):
        return Tensor(self.data, tuple(mapping.get(l, l) for l in self.labels))

    # -- arithmetic ----------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Tensor):
            return contract(self, other)
        return Tensor(self.data * other, self.labels)

    __rmul__ = __mul__

    def __truediv__(self, s):
        return Tensor(self.data / s, self.labels)

    def __add__(self, other):
        return Tensor(self.data + other.array(self.labels), self.labels)

    def __sub__(self, other):
        return Tensor(self.data - other.array(self.labels), self.labels)

    def __neg__(self):
        return Tensor(-self.data, self.labels)

    def __repr__(self):
        return f"Tensor(shape={self.data.shape}, labels={self.labels})"


def contract(A: Tensor, B: Tensor) -> Tensor:
    """ITensor `*`: sum over shared labels; result = (unshared of A)..., (unshared of B)..."""
    shared = [l for l in A.labels if l in B.labels]
    ax_a = [A.labels.index(l) for l in shared]
    ax_b = [B.labels.index(l) for l in shared]
    out = np.tensordot(A.data, B.data, axes=(ax_a, ax_b))
    labels = tuple(l for l in A.labels if l not in shared) + tuple(
        l for l in B.labels if l not in shared
    )
    return Tensor(out, labels)


def prime(T: Tensor, which=None, inc=1) -> Tensor:
    """Raise the prime level of all labels, or of those in `which` (iterable of labels)."""
    if which is not None:
        which = set(which)
    return Tensor(
        T.data,
        tuple(
            (k, n, p + inc) if (which is None or (k, n, p) in which) else (k, n, p)
            for (k, n, p) in T.labels
        ),
    )


def noprime(T: Tensor) -> Tensor:
    return Tensor(T.data, tuple((k, n, 0) for (k, n, p) in T.labels))


def dag(T: Tensor) -> Tensor:
    return Tensor(np.conj(T.data), T.labels)


def inner(A: Tensor, B: Tensor):
    """<A|B> with matching labels."""
    return np.vdot(A.array(A.labels).ravel(), B.array(A.labels).ravel())


def commonlabels(A: Tensor, B: Tensor):
    return [l for l in A.labels if l in B.labels]


def uniquelabels(A: Tensor, B: Tensor):
    return [l for l in A.labels if l not in B.labels]


# ---------------------------------------------------------------------------
# Truncation rule -- UPSTREAM NDTensors `truncate!`
# ---------------------------------------------------------------------------
def truncate_spectrum(P, *, cutoff=0.0, mindim=1, maxdim=None):
    """P: eigenvalues of rho (= sigma^2), sorted descending.  Returns (nkeep, truncerr).

    defaults use_relative_cutoff=true, use_absolute_cutoff=false.
    """
    P = np.array(P, dtype=float)
    origm = len(P)
    if maxdim is None:
        maxdim = origm
    # zero out negative tail
    for n in range(origm - 1, -1, -1):
        if P[n] >= 0.0:
            break
        P[n] = 0.0
    if origm == 1:
        return 1, 0.0
    n = origm
    truncerr = 0.0
    while n > maxdim:
        truncerr += P[n - 1]
        n -= 1
    scale = P.sum()
    if scale == 0.0:
        scale = 1.0
    while n > mindim and (truncerr + P[n - 1] <= cutoff * scale):
        truncerr += P[n - 1]
        n -= 1
    truncerr /= scale
    if n < 1:
        n = 1
    return n, float(truncerr)


def _matricize(T: Tensor, left):
    left = list(left)
    right = [l for l in T.labels if l not in left]
    M = T.array(left + right)
    dl = int(np.prod([T.dim(l) for l in left])) if left else 1
    dr = int(np.prod([T.dim(l) for l in right])) if right else 1
    return M.reshape(dl, dr), left, right


def qr(T: Tensor, left, newlabel):
    """Thin QR.  UPSTREAM `qr(A, Linds)`: Q gets (left..., new), R (new, right...)."""
    M, left, right = _matricize(T, left)
    Q, R = np.linalg.qr(M, mode="reduced")
    k = Q.shape[1]
    Qt = Tensor(Q.reshape([T.dim(l) for l in left] + [k]), left + [newlabel])
    Rt = Tensor(R.reshape([k] + [T.dim(l) for l in right]), [newlabel] + right)
    return Qt, Rt


def svd_trunc(T: Tensor, left, newlabel, *, cutoff=None, mindim=1, maxdim=None):
    """Truncated SVD, truncation on sigma^2 (UPSTREAM `svd`; App. A.5).  Returns U, S(vector), V, truncerr."""
    M, left, right = _matricize(T, left)
    try:
        U, s, Vh = np.linalg.svd(M, full_matrices=False)
    except np.linalg.LinAlgError:  # UPSTREAM falls back from gesdd to gesvd
        import scipy.linalg

        U, s, Vh = scipy.linalg.svd(M, full_matrices=False, lapack_driver="gesvd")
    P = s**2
    if cutoff is None:
        n = min(len(P), maxdim if maxdim is not None else len(P))
        terr = float(P[n:].sum() / max(P.sum(), 1e-300))
    else:
        n, terr = truncate_spectrum(P, cutoff=cutoff, mindim=mindim, maxdim=maxdim)
    U, s, Vh = U[:, :n], s[:n], Vh[:n, :]
    Ut = Tensor(U.reshape([T.dim(l) for l in left] + [n]), left + [newlabel])
    Vt = Tensor(Vh.reshape([n] + [T.dim(l) for l in right]), [newlabel] + right)
    return Ut, s, Vt, terr


def eigen_trunc(rho: Tensor, left, newlabel, *, cutoff=None, mindim=1, maxdim=None, rows_primed=False):
    """Hermitian eigendecomposition, eigenvalues descending, truncated by A.5.

    UPSTREAM `eigen(rho; ishermitian=true, cutoff, mindim, maxdim)` (App. A.8).
    `rho` has labels left... and prime(left)....  Returns D (vector), U (left..., new), truncerr.
    """
    left = list(left)
    right = [(k, n, p + 1) for (k, n, p) in left]
    # the matrix whose eigenvectors are wanted: rows = the index set that rho maps *to*.  factorize builds
    # rho[l, l'] = theta theta^dagger (rows unprimed); the expansion builds rho[b', b] = S S^dagger (rows primed).
    M = rho.array(right + left) if rows_primed else rho.array(left + right)
    d = int(np.prod([rho.dim(l) for l in left]))
    M = M.reshape(d, d)
    M = 0.5 * (M + M.conj().T)
    w, V = np.linalg.eigh(M)
    order = np.argsort(-w, kind="stable")
    w, V = w[order], V[:, order]
    if cutoff is None and maxdim is None:
        n, terr = len(w), 0.0
    else:
        n, terr = truncate_spectrum(
            w, cutoff=0.0 if cutoff is None else cutoff, mindim=mindim, maxdim=maxdim
        )
    w, V = w[:n], V[:, :n]
    Ut = Tensor(V.reshape([rho.dim(l) for l in left] + [n]), left + [newlabel])
    return w, Ut, terr


AUTOMATIC_CUTOFF = 1e-12


def factorize(T: Tensor, left, newlabel, *, cutoff=None, mindim=1, maxdim=None):
    """UPSTREAM `ITensors.factorize(A, Linds; cutoff, mindim, maxdim, ortho="left")` decision
    rule; called from src/inserter.jl:23.  Returns L, R, info dict."""
    left = [l for l in T.labels if l in set(left)]
    right = [l for l in T.labels if l not in left]
    dL = int(np.prod([T.dim(l) for l in left])) if left else 1
    dR = int(np.prod([T.dim(l) for l in right])) if right else 1
    if maxdim is None:
        maxdim = min(dL, dR)
    maxdim = min(maxdim, dL, dR)
    might_truncate = (cutoff is not None) or maxdim < min(dL, dR)
    if not might_truncate:
        Q, R = qr(T, left, newlabel)
        return Q, R, {"decomp": "qr", "truncerr": 0.0, "spectrum": None}
    if cutoff is None or cutoff <= AUTOMATIC_CUTOFF:
        U, s, V, terr = svd_trunc(T, left, newlabel, cutoff=cutoff, mindim=mindim, maxdim=maxdim)
        R = Tensor(s.reshape([-1] + [1] * (V.data.ndim - 1)) * V.data, V.labels)
        return U, R, {"decomp": "svd", "truncerr": terr, "spectrum": s**2}
    # eigen route: rho = T T^dagger over the right labels
    Tp = prime(dag(T), left)
    rho = contract(T, Tp)
    w, U, terr = eigen_trunc(rho, left, newlabel, cutoff=cutoff, mindim=mindim, maxdim=maxdim)
    R = contract(dag(U), T)
    R = R.permute([newlabel] + right)
    return U, R, {"decomp": "eigen", "truncerr": terr, "spectrum": w}


def directsum(A: Tensor, a, B: Tensor, b, newlabel=None):
    """UPSTREAM `directsum(A=>a, B=>b)` (App. A.9): concatenate along a / b; other labels must match."""
    others = [l for l in A.labels if l != a]
    assert set(others) == set(l for l in B.labels if l != b)
    Aa = A.array(others + [a])
    Bb = B.array(others + [b])
    out = np.concatenate([Aa, Bb], axis=-1)
    lab = newlabel if newlabel is not None else a
    return Tensor(out, others + [lab]).permute([lab if l == a else l for l in A.labels])
