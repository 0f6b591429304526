"""Parity of the two randomised pieces of the path against the oracle with the SAME random numbers:
the host injects the probes, so device and oracle follow the same rule on the same draws.

  range_finder      src/sketched_linear_algebra/range_finder.jl:6-64  vs oracle/range_finder.py
                    (matrix map and the projected operator psi -> optimal_map(P, psi) as the linear map)
  "ortho" expansion src/subspace/ortho_subspace.jl:19-77              vs oracle/subspace.py::subspace_expand_ortho

The range finder's vectors are unique given the probe sequence (q_k = normalised Gram-Schmidt residual), so they are compared
entry by entry (1e-10).  The expansion's new basis is unique up to the sign convention of the SVD, so the projector onto the
enlarged basis is compared (1e-10) together with the new bond dimension."""
import numpy as np
import pytest

from helpers import _olabel, to_oracle_ttn

pytestmark = pytest.mark.gpu


def _rand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    return a + 1j * rng.standard_normal(shape) if cplx else a


class _Draws:
    """random_vector() / rng stand-in that replays a fixed sequence of draws."""

    def __init__(self, cols):
        self.cols, self.i = cols, 0

    def __call__(self):
        v = self.cols[:, self.i].copy()
        self.i += 1
        return v


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("case", ["exhaust", "capped", "panels", "cutoff"])
def test_range_finder_matrix_matches_oracle(cplx, case):
    import networksolvers_b200 as ns
    from oracle.range_finder import range_finder_map
    ctx = ns.default_context()
    rng = np.random.default_rng(41)
    if case == "exhaust":        # rank 6 map: stops when the residual norm drops below 1e-12
        # (entries scaled so that the rounding noise of an exhausted residual, eps |A| |x|, is far below the absolute 1e-12 test)
        A, max_rank, kw = 1e-2 * _rand(rng, (80, 6), cplx) @ _rand(rng, (6, 50), cplx), 20, {}
    elif case == "capped":       # max_rank + oversample vectors
        A, max_rank, kw = _rand(rng, (64, 40), cplx), 7, dict(oversample=3)
    elif case == "panels":       # several 32-vector panels, north_pass = 1
        A, max_rank, kw = _rand(rng, (300, 120), cplx), 70, dict(north_pass=1)
    else:                        # experimental cutoff rule: keep the first vector below the cutoff, then stop
        A = _rand(rng, (60, 60), cplx) * (10.0 ** -np.arange(60))[None, :]
        max_rank, kw = 30, dict(cutoff=1e-3)
    m, n = A.shape
    ndraw = min(max_rank + kw.get("oversample", 2), m, n) + 1
    draws = _rand(rng, (n, ndraw), cplx)
    ref = range_finder_map(lambda x: A @ x, _Draws(draws), max_rank=max_rank, **kw)
    Q = ctx.range_finder(A, max_rank=max_rank, probes=draws[:, 1:], **kw)      # the first draw only sizes the domain (:55)
    assert Q.shape[1] == len(ref), (Q.shape, len(ref))
    Qr = np.stack(ref, axis=1) if ref else np.zeros((m, 0))
    assert np.abs(Q - Qr).max() <= 1e-10, np.abs(Q - Qr).max()
    assert np.abs(Q.conj().T @ Q - np.eye(Q.shape[1])).max() < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
def test_range_finder_projected_operator_matches_oracle(cplx):
    """linear_map = the H_eff closure of src/eigsolve.jl:22 at a two-site position of a chain."""
    import networksolvers_b200 as ns
    from oracle.operator_map import optimal_map
    from oracle.projttn import ProjTTN, position
    from oracle.range_finder import range_finder_map
    from oracle.tensor import Tensor
    g = ns.path_graph(8)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi = ns.random_state(sites, 6, seed=9, dtype=complex if cplx else float)
    net = ns.EigsolveProblem(state=psi, operator=H).net
    region = [4, 5]
    net.extract(region)
    legs, dims = net.local_info()
    n = int(np.prod(dims))
    rng = np.random.default_rng(43)
    max_rank = 9
    draws = _rand(rng, (n, max_rank + 3), cplx)
    psio = to_oracle_ttn(net.to_host())
    P = position(ProjTTN(to_oracle_ttn(H, operator=True)), psio, region)
    labels = [_olabel(l) for l in legs]

    def lin(x):
        th = Tensor(np.reshape(x, dims, order="F"), labels)
        return optimal_map(P, th).array(labels).reshape(-1, order="F")

    ref = range_finder_map(lin, _Draws(draws), max_rank=max_rank)
    Q = net.range_finder(max_rank, probes=draws[:, 1:])
    assert Q.shape[-1] == len(ref) == max_rank + 2
    Qm = Q.reshape(n, -1, order="F")
    assert np.abs(Qm - np.stack(ref, axis=1)).max() <= 1e-10
    # device Philox probes: same guarantees, no oracle comparison possible
    Q2 = net.range_finder(5, seed=3).reshape(n, -1, order="F")
    assert Q2.shape[1] == 7 and np.abs(Q2.conj().T @ Q2 - np.eye(7)).max() < 1e-12


class _RegionIter:
    def __init__(self, prev, cur):
        self._p, self._c = prev, cur

    def previous_region(self):
        return self._p

    def current_region(self):
        return self._c


class _FixedNormal:
    def __init__(self, arr_re, arr_im=None):
        self.q = [arr_re] + ([arr_im] if arr_im is not None else [])

    def standard_normal(self, shape):
        a = self.q.pop(0)
        assert list(a.shape) == list(shape), (a.shape, shape)
        return a


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("graph_kind", ["chain", "tree"])
def test_ortho_expansion_matches_oracle(graph_kind, cplx):
    import networksolvers_b200 as ns
    from networksolvers_b200 import _lib as L
    from networksolvers_b200.models import canonical_legs
    from oracle.gauge import orthogonalize
    from oracle.subspace import compute_expansion, expand_space, subspace_expand_ortho
    if graph_kind == "chain":
        g, prev, cur = ns.path_graph(8), 4, 5
    else:
        g, prev, cur = ns.star_of_chains(3, 3), (0, 0), (1, 1)     # degree-3 centre: the bond is its FIRST leg
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi = ns.random_state(sites, 4, seed=21, dtype=complex if cplx else float)
    prob = ns.EigsolveProblem(state=psi, operator=H)
    net = prob.net
    net.extract([prev])
    net.insert((0.0, 1, 1 << 30))                 # position = [prev], centre on prev (one-site region: plain write-back)
    before = net.to_host()
    # sizes as the reference computes them (ortho_subspace.jl:44-60)
    legs_prev = canonical_legs(g, prev)
    a_leg = ("link", prev, cur)
    basis_legs = [l for l in legs_prev if l != a_leg]
    tprev, lp = before.tensors[prev], before.legs[prev]
    shape = {l: tprev.shape[lp.index(l)] for l in lp}
    nb, cur_dim = int(np.prod([shape[l] for l in basis_legs])), shape[a_leg]
    max_expand, maxdim, factor = 2, 64, 1.5
    axd = expand_space(nb, factor)
    assert compute_expansion(cur_dim, nb, expansion_factor=factor, max_expand=max_expand, maxdim=maxdim) == 2 < nb - cur_dim
    rng = np.random.default_rng(77)
    pshape = [shape[l] for l in basis_legs] + [axd]
    pre, pim = rng.standard_normal(pshape), (rng.standard_normal(pshape) if cplx else None)
    probe = pre + 1j * pim if cplx else pre
    # device
    net.set_expand_probe(np.reshape(np.asfortranarray(probe), (nb, axd), order="F"))
    info = net.extract([cur], (0.0, 1, maxdim), dict(algorithm=L.NSB_EXPAND_ORTHO, north_pass=1, expansion_factor=factor, max_expand=max_expand))
    assert info.expanded == 1
    after = net.to_host()
    Ad, ld = after.tensors[prev], after.legs[prev]
    Ad = np.transpose(Ad, [ld.index(l) for l in basis_legs + [a_leg]]).reshape(nb, -1, order="F")
    assert Ad.shape[1] == cur_dim + 2
    # oracle on the same pre-expansion state with the same draw
    psio = orthogonalize(to_oracle_ttn(before), [cur])

    class _P:
        state = psio
    local = psio[cur]
    out = subspace_expand_ortho(_P, local, _RegionIter([prev], [cur]), maxdim=maxdim, expansion_factor=factor, max_expand=max_expand,
                                rng=_FixedNormal(pre, pim))
    Ao = psio[prev]
    Ao = Ao.array([_olabel(l) for l in basis_legs + [a_leg]]).reshape(nb, -1, order="F")
    assert Ao.shape[1] == Ad.shape[1]
    Pd, Po = Ad @ Ad.conj().T, Ao @ Ao.conj().T
    assert np.abs(Pd - Po).max() <= 1e-10, np.abs(Pd - Po).max()
    assert np.abs(Ad.conj().T @ Ad - np.eye(Ad.shape[1])).max() < 1e-12
    # the local tensor is the same physical tensor: contract the enlarged basis back in
    th, lt = net.local_download()
    ia = [i for i, l in enumerate(lt) if l[0] == "link" and set(l[1:]) == {cur, prev}]
    assert len(ia) == 1 and th.shape[ia[0]] == cur_dim + 2
