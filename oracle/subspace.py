"""Subspace expansion (oracle; test-only).  Restates src/subspace/subspace.jl:5-48,
src/subspace/densitymatrix.jl:5-74 and src/subspace/ortho_subspace.jl:4-77."""
from __future__ import annotations

import math
import sys

import numpy as np

from .tensor import (Tensor, contract, dag, prime, noprime, commonlabels, uniquelabels, eigen_trunc,
                     svd_trunc, directsum)
from .truncation_parameters import DEFAULT_MAXDIM, DEFAULT_CUTOFF, DEFAULT_MINDIM

DEFAULT_EXPANSION_FACTOR = 1.5
DEFAULT_MAX_EXPAND = sys.maxsize


def compute_expansion(current_dim, basis_size, *, expansion_factor=DEFAULT_EXPANSION_FACTOR,
                      max_expand=DEFAULT_MAX_EXPAND, maxdim=DEFAULT_MAXDIM):
    """src/subspace/subspace.jl:28-48."""
    expand_maxdim = math.ceil(expansion_factor * current_dim)
    expand_maxdim = min(max_expand, expand_maxdim)
    expand_maxdim = min(basis_size - current_dim, expand_maxdim)
    expand_maxdim = min(maxdim - current_dim, expand_maxdim)
    return max(0, expand_maxdim)


def subspace_expand(problem, local_state, region_iterator, *, subspace_algorithm=None, sweep=None,
                    trunc, **kws):
    """Dispatcher, src/subspace/subspace.jl:8-26."""
    if subspace_algorithm is None:
        return problem, local_state
    if subspace_algorithm == "densitymatrix":
        return subspace_expand_densitymatrix(problem, local_state, region_iterator, trunc=trunc, **kws)
    raise ValueError(
        "Subspace expansion (subspace_expand!) not defined for requested combination of "
        "subspace_algorithm and problem types")


_U = ("x", "u", 0)
_AX = ("x", "ax", 0)


def subspace_expand_densitymatrix(problem, local_state, region_iterator, *, north_pass=1,
                                  expansion_factor=DEFAULT_EXPANSION_FACTOR,
                                  max_expand=DEFAULT_MAX_EXPAND, trunc, **kws):
    """src/subspace/densitymatrix.jl:5-74."""
    region = region_iterator.current_region()
    psi = problem.state.copy()
    P = problem.operator
    prev_set = [v for v in P.pos if v not in region]
    if len(prev_set) != 1 or P.on_edge():
        return problem, local_state
    prev_vertex = prev_set[0]
    A = psi[prev_vertex]
    next_vertices = [v for v in region if commonlabels(psi[v], A)]
    if not next_vertices:
        return problem, local_state
    assert len(next_vertices) == 1
    next_vertex = next_vertices[0]
    C = psi[next_vertex]
    common = commonlabels(A, C)
    if not common:
        return problem, local_state
    a = common[0]
    basis = uniquelabels(A, C)
    basis_size = int(np.prod([A.dim(l) for l in basis]))
    expanded_maxdim = compute_expansion(A.dim(a), basis_size, expansion_factor=expansion_factor,
                                        max_expand=max_expand, maxdim=trunc["maxdim"])
    if expanded_maxdim <= 0:
        return problem, local_state
    trunc = dict(trunc, maxdim=expanded_maxdim)

    sqrt_rho = A
    for e in P.incident_edges():
        if e[0] in region or e[1] in region:
            continue
        sqrt_rho = contract(sqrt_rho, P.environments[e])
    sqrt_rho = contract(sqrt_rho, P.operator[prev_vertex])

    Ap = prime(A)

    def conj_proj_A(T):
        return T - contract(Ap, contract(dag(Ap), T))

    for _ in range(north_pass):
        sqrt_rho = conj_proj_A(sqrt_rho)
    rho = contract(sqrt_rho, dag(noprime(sqrt_rho)))       # labels: basis' (plev 1), basis (plev 0)
    qn = getattr(psi, "qn", None)
    if qn is not None:
        from .qn import label_charges, multi_index_charges, eigen_trunc_qn
        basis_p = [(k, n, p + 1) for (k, n, p) in basis]
        nbasis = int(np.prod([A.dim(l) for l in basis]))
        Mr = rho.array(basis_p + list(basis)).reshape(nbasis, nbasis)
        keys = multi_index_charges([label_charges(qn, prev_vertex, x) for x in basis])
        D, Um, newk, _ = eigen_trunc_qn(Mr, keys, **trunc)
        U = Tensor(Um.reshape([A.dim(l) for l in basis] + [Um.shape[1]]), list(basis) + [_U])
    else:
        D, U, _ = eigen_trunc(rho, basis, _U, rows_primed=True, **trunc)     # U: basis (plev 0) + [u]

    Apa = prime(A, [a])

    def Uproj(T):
        return T - contract(Apa, contract(dag(Apa), T))

    for _ in range(north_pass):
        U = Uproj(U)
    ovl = contract(dag(U), A).norm()
    if ovl > 1e-10:
        print("Warning: |U*A| = %.3E in subspace expansion" % ovl)
        return problem, local_state

    Ax = directsum(A, a, U, _U, newlabel=_AX)
    expander = contract(dag(Ax), A)                        # labels: [ax, a]
    if qn is not None:
        old = qn.side_charge(next_vertex, prev_vertex)     # charges of the subtree on prev's side
        qn.set_link(prev_vertex, next_vertex, np.concatenate([old, newk], axis=0))
    psi[prev_vertex] = Ax.relabel({_AX: a})
    tmp = ("x", "a_old", 0)
    exp2 = expander.relabel({a: tmp}).relabel({_AX: a})    # [a(new), a_old]
    Cn = contract(exp2, C.relabel({a: tmp}))
    psi[next_vertex] = Cn.permute(C.labels)
    ls = contract(exp2, local_state.relabel({a: tmp}))
    local_state = ls.permute(local_state.labels)
    return problem.setproperties(state=psi), local_state


def expand_space(chi, expansion_factor):
    """src/subspace/ortho_subspace.jl:4."""
    return max(chi + 1, math.floor(expansion_factor * chi))


def subspace_expand_ortho(problem, local_tensor, region_iterator, *, cutoff=DEFAULT_CUTOFF,
                          maxdim=DEFAULT_MAXDIM, mindim=DEFAULT_MINDIM,
                          expansion_factor=DEFAULT_EXPANSION_FACTOR, max_expand=DEFAULT_MAX_EXPAND,
                          rng=None, **kws):
    """src/subspace/ortho_subspace.jl:19-77 (`subspace_expand!`, Backend"ortho"; unreachable from the
    reference's dispatcher, restated for completeness).  Mutates problem.state like the original."""
    prev_region = region_iterator.previous_region()
    region = region_iterator.current_region()
    if prev_region is None:
        return local_tensor
    prev_set = [v for v in prev_region if v not in region]
    if len(prev_set) != 1:
        return local_tensor
    prev_vertex = prev_set[0]
    psi = problem.state
    A = psi[prev_vertex]
    next_vertices = [v for v in region if commonlabels(psi[v], A)]
    if not next_vertices:
        return local_tensor
    next_vertex = next_vertices[0]
    C = psi[next_vertex]
    common = commonlabels(A, C)
    if not common:
        return local_tensor
    a = common[0]
    basis = uniquelabels(A, C)
    basis_size = int(np.prod([A.dim(l) for l in basis]))
    ax_dim = expand_space(basis_size, expansion_factor)
    rng = rng or np.random.default_rng(0)

    def linear_map(w):
        return w - contract(A, contract(dag(A), w))

    shape = [A.dim(l) for l in basis] + [ax_dim]
    rnd = rng.standard_normal(shape)
    if np.iscomplexobj(A.data):
        rnd = rnd + 1j * rng.standard_normal(shape)
    Y = linear_map(Tensor(rnd, list(basis) + [_AX]))
    expand_maxdim = compute_expansion(A.dim(a), basis_size, expansion_factor=expansion_factor,
                                      max_expand=max_expand, maxdim=maxdim)
    if Y.norm() <= 1e-15 or expand_maxdim <= 0:
        return local_tensor
    Ux, S, V, _ = svd_trunc(Y, basis, _U, cutoff=1e-14, maxdim=expand_maxdim)
    Ux = linear_map(Ux)
    Ax = directsum(A, a, Ux, _U, newlabel=_AX)
    expander = contract(dag(Ax), A)
    tmp = ("x", "a_old", 0)
    exp2 = expander.relabel({a: tmp}).relabel({_AX: a})
    psi[prev_vertex] = Ax.relabel({_AX: a})
    psi[next_vertex] = contract(exp2, C.relabel({a: tmp})).permute(C.labels)
    out = contract(exp2, local_tensor.relabel({a: tmp})).permute(local_tensor.labels)
    return out
