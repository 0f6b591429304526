"""The oracle's H_eff application (restating src/operator_map.jl:3-42) against the *definition* of the projected operator,
computed densely and without any environment: H_eff = B^dagger H B, where B embeds the local tensor of the region into the full
Hilbert space through all the other tensors of the state (orthonormal or not) and H is the dense Hamiltonian built from the
operator sum (oracle/ed.py, independent of the TTNO and of every contraction order).  Chain and tree, 2-site, 1-site and the
0-site (on-edge) map of the 1-site TDVP backward step (src/applyexp.jl:37-40), real and complex."""
import numpy as np
import pytest

from oracle.ed import dense_hamiltonian
from oracle.gauge import orthogonalize
from oracle.graph import path_graph, named_comb_tree
from oracle.models import heisenberg_opsum, random_ttn, spin_ops, ttno
from oracle.operator_map import operator_map, optimal_map
from oracle.projttn import ProjTTN, position
from oracle.tensor import Tensor, contract, site

CASES = [(path_graph(6), [3, 4]), (path_graph(6), [1, 2]), (path_graph(6), [4]), (named_comb_tree([2, 3, 2]), [(2, 1), (2, 2)]),
         (named_comb_tree([2, 3, 2]), [(2, 1)]), (named_comb_tree([2, 3, 2]), [(1, 1), (2, 1)])]


def _rest_of_state(psi, region):
    """Product of every tensor of the state outside the region (open: the region's links and the other vertices' sites)."""
    rest = None
    for v in psi.graph.vertices:
        if v in region:
            continue
        rest = psi.tensors[v] if rest is None else contract(rest, psi.tensors[v])
    return rest


@pytest.mark.parametrize("graph,region", CASES)
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("gauge", [True, False])
def test_heff_equals_the_dense_projected_hamiltonian(graph, region, cplx, gauge):
    d, ops, _ = spin_ops("S=1/2")
    os_ = heisenberg_opsum(graph)
    H = ttno(os_, graph, ops, dtype=complex if cplx else float)
    psi = random_ttn(graph, d, 3, seed=11, dtype=complex if cplx else float)
    if gauge:
        psi = orthogonalize(psi, region)
    P = position(ProjTTN(H), psi, region)
    theta = psi.tensors[region[0]]
    for v in region[1:]:
        theta = contract(theta, psi.tensors[v])
    rng = np.random.default_rng(5)
    x = rng.standard_normal(theta.shape) + (1j * rng.standard_normal(theta.shape) if cplx else 0.0)
    theta = Tensor(x, theta.labels)                                    # any local tensor, not only the state's own

    verts = list(graph.vertices)
    Hd = dense_hamiltonian(os_, graph, ops, sparse=False)
    rest = _rest_of_state(psi, region)
    full = contract(rest, theta).array([site(v) for v in verts])
    y = (Hd @ full.ravel()).reshape(full.shape)
    Y = Tensor(y, [site(v) for v in verts])
    ref = contract(Tensor(np.conj(rest.data), rest.labels), Y).array(theta.labels)     # B^dagger (H (B theta))

    for fmap in (optimal_map, operator_map):
        out = fmap(P, theta).array(theta.labels)
        assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max(), fmap.__name__


@pytest.mark.parametrize("graph,v1,v2", [(path_graph(6), 3, 4), (path_graph(6), 2, 1), (named_comb_tree([2, 3, 2]), (2, 1), (2, 2)),
                                         (named_comb_tree([2, 3, 2]), (2, 1), (1, 1))])
@pytest.mark.parametrize("cplx", [False, True])
def test_on_edge_map_equals_the_dense_projected_hamiltonian(graph, v1, v2, cplx):
    """The 0-site map of the backward step (src/applyexp.jl:37-40, src/operator_map.jl:17-22): psi[v1] = Q of a QR toward v2,
    the operator positioned on the edge, applied to the bond tensor R."""
    from oracle.sweep import _QR
    from oracle.tensor import qr, uniquelabels
    d, ops, _ = spin_ops("S=1/2")
    os_ = heisenberg_opsum(graph)
    H = ttno(os_, graph, ops, dtype=complex if cplx else float)
    psi = orthogonalize(random_ttn(graph, d, 3, seed=7, dtype=complex if cplx else float), [v1])
    Q, R = qr(psi.tensors[v1], uniquelabels(psi.tensors[v1], psi.tensors[v2]), _QR)
    psi = psi.copy()
    psi[v1] = Q
    P = position(ProjTTN(H), psi, [("edge", v1, v2)])
    assert P.on_edge()
    rng = np.random.default_rng(9)
    x = rng.standard_normal(R.shape) + (1j * rng.standard_normal(R.shape) if cplx else 0.0)
    Rx = Tensor(x, R.labels)

    verts = list(graph.vertices)
    Hd = dense_hamiltonian(os_, graph, ops, sparse=False)
    rest = _rest_of_state(psi, [])
    full = contract(rest, Rx).array([site(v) for v in verts])
    Y = Tensor((Hd @ full.ravel()).reshape(full.shape), [site(v) for v in verts])
    ref = contract(Tensor(np.conj(rest.data), rest.labels), Y).array(Rx.labels)
    for fmap in (optimal_map, operator_map):
        out = fmap(P, Rx).array(Rx.labels)
        assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max(), fmap.__name__
