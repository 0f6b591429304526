"""Invariants of the subspace expansion (src/subspace/densitymatrix.jl:5-74, src/subspace/ortho_subspace.jl:19-77) that hold
whatever the implementation, checked on the oracle's restatement through its own extracter: the expansion changes the *basis* on
the bond behind the region, never the state; the enlarged basis tensor stays an isometry containing the old one; the growth obeys
`compute_expansion` (src/subspace/subspace.jl:31-48)."""
import numpy as np
import pytest

from oracle.ed import state_vector
from oracle.graph import path_graph, named_comb_tree
from oracle.models import heisenberg_opsum, random_ttn, spin_ops, ttno
from oracle.projttn import ProjTTN
from oracle.subspace import compute_expansion
from oracle.sweep import EigsolveProblem, RegionIterator, extracter
from oracle.tensor import contract, dag, link, prime


def _plan(regions):
    return [(list(r), {}) for r in regions]


@pytest.mark.parametrize("graph,regions,bond", [
    (path_graph(8), [[3], [4]], (3, 4)),
    (path_graph(8), [[6], [5]], (6, 5)),
    (named_comb_tree([2, 3, 2]), [[(2, 2)], [(2, 1)]], ((2, 2), (2, 1))),
])
@pytest.mark.parametrize("alg", ["densitymatrix", "ortho"])
@pytest.mark.parametrize("factor,max_expand", [(1.5, 10**9), (1.1, 10**9), (2.0, 1)])
def test_expansion_changes_the_basis_not_the_state(graph, regions, bond, alg, factor, max_expand):
    d, ops, _ = spin_ops("S=1/2")
    H = ttno(heisenberg_opsum(graph), graph, ops)
    chi = 2                                                   # small on purpose: there is room to expand
    psi0 = random_ttn(graph, d, chi, seed=21)
    prob = EigsolveProblem(psi0, ProjTTN(H))
    ri = RegionIterator(prob, _plan(regions))
    trunc = dict(cutoff=1e-12, maxdim=50)
    # first region without expansion (the operator now sits there), then the expansion toward the previous vertex
    ri.which_region = 1
    prob, _ = extracter(prob, ri, sweep=1, trunc=trunc)
    ri.problem = prob
    ri.which_region = 2
    before = state_vector(prob.state)
    prev_v, next_v = bond
    dim_before = prob.state.linkdim(prev_v, next_v)
    if alg == "densitymatrix":
        prob2, local = extracter(prob, ri, sweep=1, trunc=trunc, subspace_algorithm=alg, expansion_factor=factor, max_expand=max_expand)
    else:
        # `subspace_expand!` with Backend"ortho" is not reachable from the reference's dispatcher (src/subspace/subspace.jl:8-14 calls
        # the bang-less name): the oracle restates it as a function of its own, applied after the plain extracter
        from oracle.subspace import subspace_expand_ortho
        prob2, local = extracter(prob, ri, sweep=1, trunc=trunc)
        prob2 = prob2.setproperties(state=prob2.state.copy())
        local = subspace_expand_ortho(prob2, local, ri, maxdim=trunc["maxdim"], expansion_factor=factor, max_expand=max_expand,
                                      rng=np.random.default_rng(3))
    after = state_vector(prob2.state)
    # (1) same state (the local tensor returned is the region's tensor of that state)
    assert np.abs(after - before).max() <= 1e-12 * np.abs(before).max()
    assert np.abs(local.array(prob2.state[next_v].labels) - prob2.state[next_v].data).max() <= 1e-13
    # (2) the basis tensor behind the region is still an isometry onto the (possibly larger) bond
    A = prob2.state[prev_v]
    a = link(prev_v, next_v)
    G = contract(dag(prime(A, [a])), A).array([(a[0], a[1], 1), a])
    assert np.abs(G - np.eye(G.shape[0])).max() <= 1e-12
    # (3) growth bounded by the rule
    dim_after = prob2.state.linkdim(prev_v, next_v)
    basis = int(np.prod([A.dim(l) for l in A.labels if l != a]))
    bound = compute_expansion(dim_before, basis, expansion_factor=factor, max_expand=max_expand, maxdim=trunc["maxdim"])
    assert dim_before <= dim_after <= min(basis, dim_before + bound)      # both algorithms add at most compute_expansion(...)
    if basis > dim_before and bound > 0:
        assert dim_after > dim_before                                     # a generic state has something to add
