"""Properties of the oracle's gauge walk (UPSTREAM `itn.orthogonalize`; called from src/extracter.jl:6): the
state is unchanged, every tensor outside the region is an isometry toward the region, a second call is a no-op, and moving the
region only re-factorises the tensors on the path between the two regions."""
import numpy as np
import pytest

from oracle.ed import state_vector
from oracle.gauge import orthogonalize
from oracle.graph import path_graph, named_comb_tree, tree_path
from oracle.models import random_ttn, spin_ops
from oracle.tensor import contract, dag, link, prime

GRAPHS = [path_graph(7), named_comb_tree([2, 3, 1, 2])]


def _toward(g, v, region):
    """Neighbour of v on the path to the region."""
    best = min((tree_path(g, v, r) for r in region), key=len)
    return best[1]


@pytest.mark.parametrize("g", GRAPHS)
@pytest.mark.parametrize("cplx", [False, True])
def test_orthogonalize_keeps_the_state_and_makes_isometries(g, cplx):
    d, _, _ = spin_ops("S=1/2")
    psi0 = random_ttn(g, d, 3, seed=4, dtype=complex if cplx else float)
    ref = state_vector(psi0)
    verts = list(g.vertices)
    regions = [[verts[0]], [verts[-1]], [verts[2]]] + [[a, b] for a in verts for b in g.neighbors(a)][:4]
    psi = psi0
    for region in regions:
        before = {v: psi[v] for v in verts}
        old_region = list(psi.ortho_region)
        psi = orthogonalize(psi, region)
        assert set(psi.ortho_region) == set(region)                # (a region that is set-equal to the current one is a no-op)
        out = state_vector(psi)
        assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
        for v in verts:
            if v in region:
                continue
            n = _toward(g, v, region)
            a = link(v, n)
            G = contract(dag(prime(psi[v], [a])), psi[v]).array([(a[0], a[1], 1), a])
            assert np.abs(G - np.eye(G.shape[0])).max() <= 1e-12, (region, v)
        # only tensors on the paths between the old and the new region were touched
        if old_region:
            on_path = set(region) | set(old_region)
            for r0 in old_region:
                for r1 in region:
                    on_path |= set(tree_path(g, r0, r1))
            for v in verts:
                if v not in on_path:
                    assert psi[v] is before[v], (region, v)
        assert orthogonalize(psi, region) is psi                  # already there: no work, same object
