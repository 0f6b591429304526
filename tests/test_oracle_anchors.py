"""Pins the oracle on everything the reference itself pins for the hot path:
the shipped constant (examples/dmrg.jl:41-43), the tree-DMRG energy check
(test/dmrg/test_tree_dmrg.jl:53,67), the TDVP checks (test/tdvp/test_tree_tdvp.jl:65-77), the
Euler-tour properties (test/test_euler_tour.jl:13-25), plus independent ED / dense expm."""
import numpy as np
import pytest

from oracle.ed import ed_ground_state, ed_time_evolution, state_vector, dense_hamiltonian, ttno_dense
from oracle.graph import build_tree, chain_plus_ancilla, path_graph, named_comb_tree
from oracle.models import heisenberg_opsum, ising_opsum, product_ttn, spin_ops, ttno, random_ttn
from oracle.region_plans import euler_tour_edges, euler_tour_vertices, tdvp_regions
from oracle.sweep import dmrg, tdvp
from oracle.local_solvers import exponentiate_solver, runge_kutta_solver


def neel(g, st, even_up=True):
    out = {}
    for j, v in enumerate(g.vertices, start=1):
        up = (j % 2 == 0) if even_up else (j % 2 == 1)
        out[v] = st["Up"] if up else st["Dn"]
    return out


def test_euler_tour_properties():
    g = build_tree(3, 3)
    tour = euler_tour_edges(g, (0, 0))
    for a, b in zip(tour[:-1], tour[1:]):
        assert a[1] == b[0]
    for (u, v) in g.edges:
        assert (u, v) in tour and (v, u) in tour
    assert len(tour) == 2 * len(g.edges)
    vt = euler_tour_vertices(g, (0, 0))
    for v in g.vertices:
        assert v in vt


@pytest.mark.parametrize("graph", [path_graph(6), build_tree(3, 2), named_comb_tree([2, 3, 1])])
def test_ttno_equals_dense_hamiltonian(graph):
    d, ops, _ = spin_ops("S=1/2")
    for os_ in (heisenberg_opsum(graph), ising_opsum(graph, 1.0, 0.7)):
        H = ttno(os_, graph, ops)
        assert np.abs(ttno_dense(H, graph, d) - dense_hamiltonian(os_, graph, ops, sparse=False)).max() < 1e-13


def test_reference_constant_s1_n10():
    """examples/dmrg.jl:26-43 shape (2-site, no expansion): Exact energy = -12.8945601."""
    g = path_graph(10)
    d, ops, st = spin_ops("S=1")
    H = ttno(heisenberg_opsum(g), g, ops)
    psi0 = product_ttn(g, d, neel(g, st))
    trunc = dict(cutoff=1e-12, maxdim=[10, 40, 80, 160])
    E, psi = dmrg(H, psi0, nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc))
    assert abs(E - (-12.8945601)) < 5e-8
    assert abs(E - (-12.894560132211)) < 5e-9


def test_tree_dmrg_two_site_and_one_site_expansion():
    """test/dmrg/test_tree_dmrg.jl:15-67."""
    g = build_tree(3, 3)
    d, ops, st = spin_ops("S=1/2")
    os_ = heisenberg_opsum(g)
    H = ttno(os_, g, ops)
    psi0 = product_ttn(g, d, neel(g, st))
    Ex, _ = ed_ground_state(os_, g, ops)
    assert abs(Ex - (-4.046057854359)) < 1e-10
    trunc = dict(cutoff=1e-5, maxdim=40)
    E, _ = dmrg(H, psi0, nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc))
    assert abs(E - Ex) < 1e-5
    E, _ = dmrg(H, psi0, nsweeps=5, nsites=1,
                extracter_kwargs=dict(trunc=trunc, subspace_algorithm="densitymatrix"),
                inserter_kwargs=dict(trunc=trunc))
    assert abs(E - Ex) < 1e-5


def test_config1_heisenberg_n20_energy():
    """BASELINE config 1: S=1/2 N=20, 2-site, maxdim 100, cutoff 1e-12; ED -8.682473334399."""
    g = path_graph(20)
    d, ops, st = spin_ops("S=1/2")
    H = ttno(heisenberg_opsum(g), g, ops)
    psi0 = product_ttn(g, d, neel(g, st))
    E, psi = dmrg(H, psi0, nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=dict(cutoff=1e-12, maxdim=100)))
    assert abs(E - (-8.682473334399)) < 1e-9
    assert psi.maxlinkdim() <= 100


@pytest.mark.parametrize("order", [2, 4])
def test_tdvp_two_site_vs_dense_expm(order):
    """examples/quench_evolution.jl:20-84 shape; fidelity target >= 1 - 1e-8."""
    g = path_graph(8)
    d, ops, st = spin_ops("S=1/2")
    os_ = heisenberg_opsum(g)
    H = ttno(os_, g, ops)
    psi0 = product_ttn(g, d, neel(g, st, even_up=False))
    tp = list(np.arange(0, 0.5 + 1e-9, 0.05))
    vx = ed_time_evolution(os_, g, ops, state_vector(psi0), tp, normalize=True)
    psit = tdvp(H, psi0, tp, nsites=2, tdvp_order=order,
                updater_kwargs=dict(solver=runge_kutta_solver, order=4),
                inserter_kwargs=dict(trunc=dict(maxdim=5000, cutoff=1e-14), normalize=True))
    assert 1 - abs(np.vdot(vx, state_vector(psit))) < 1e-8


def test_tdvp_exponentiate_solver_matches_rk4():
    g = path_graph(6)
    d, ops, st = spin_ops("S=1/2")
    os_ = heisenberg_opsum(g)
    H = ttno(os_, g, ops)
    psi0 = product_ttn(g, d, neel(g, st, even_up=False))
    tp = [0.0, 0.05, 0.1]
    vx = ed_time_evolution(os_, g, ops, state_vector(psi0), tp, normalize=True)
    psit = tdvp(H, psi0, tp, nsites=2, tdvp_order=2, updater_kwargs=dict(solver=exponentiate_solver),
                inserter_kwargs=dict(trunc=dict(cutoff=1e-14), normalize=True))
    assert 1 - abs(np.vdot(vx, state_vector(psit))) < 1e-9


def test_tree_tdvp_regression():
    """test/tdvp/test_tree_tdvp.jl:24-77 (chain + ancilla): norms, overlaps and accumulated phase."""
    N = 10
    g = chain_plus_ancilla(N)
    d, ops, st = spin_ops("S=1/2")
    from oracle.models import OpSum
    os_ = OpSum()
    for j in range(1, N):
        os_.add(1.0, "Sz", j, "Sz", j + 1)
        os_.add(0.5, "S+", j, "S-", j + 1)
        os_.add(0.5, "S-", j, "S+", j + 1)
    H = ttno(os_, g, ops)
    psi0 = product_ttn(g, d, neel(g, st))
    trunc = dict(cutoff=1e-10, maxdim=100)
    E, gs = dmrg(H, psi0, nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc))
    tmax = 0.10
    tp = list(np.arange(0, tmax + 1e-9, 0.02))
    psi1 = tdvp(H, gs, tp, nsites=1, inserter_kwargs=dict(trunc=trunc))
    psi2 = tdvp(H, gs, tp, nsites=2, inserter_kwargs=dict(trunc=trunc))
    v0, v1, v2 = state_vector(gs), state_vector(psi1), state_vector(psi2)
    assert np.linalg.norm(v1) > 0.999 and np.linalg.norm(v2) > 0.999
    assert abs(np.vdot(v1, v0)) > 0.99
    assert abs(np.vdot(v1, v2)) > 0.99
    z = np.vdot(v1, v0)
    assert abs(np.arctan(z.imag / z.real) - E * tmax) < 1e-4


def test_tdvp_plan_structure():
    g = path_graph(5)
    plan = tdvp_regions(g, 0.1, updater_kwargs={}, tdvp_order=2, nsites=2, sweep=1)
    regs = [r for r, _ in plan]
    # forward half-sweep: (N-1) two-site + (N-2) one-site regions; second half reversed
    assert len(regs) == 2 * (4 + 3)
    assert regs[len(regs) // 2:] == [list(reversed(r)) for r in reversed(regs[: len(regs) // 2])]
    ts = [k["updater_kwargs"]["time_step"] for _, k in plan[:7]]
    assert ts == [0.05, -0.05, 0.05, -0.05, 0.05, -0.05, 0.05]


def test_qn_oracle_hubbard_sector_ground_state():
    """QN-conserving oracle (block-wise factorisations, merged truncation): Hubbard N=6 stays in the (3,3) sector of
    the start state and converges to that sector's exact ground state; dense and QN runs agree on the S=1 chain."""
    import itertools
    from oracle.models import electron_ops, hubbard_chain_opsum, product_ttn_qn
    from oracle.qn import check_state_symmetric
    g = path_graph(6)
    d, ops, st = electron_ops()
    H = ttno(hubbard_chain_opsum(g, 1.0, 4.0), g, ops)
    psi0 = product_ttn_qn(g, "Electron", 4, {v: (st["Up"] if v % 2 else st["Dn"]) for v in g.vertices})
    trunc = dict(cutoff=1e-10, maxdim=[10, 20, 60])
    ek = dict(trunc=trunc, subspace_algorithm="densitymatrix", expansion_factor=1.1)
    E, psi = dmrg(H, psi0, nsweeps=6, nsites=2, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc))
    Hd = ttno_dense(H, g, 4)
    sector = [i for i, c in enumerate(itertools.product(range(4), repeat=6))
              if sum(x in (1, 3) for x in c) == 3 and sum(x in (2, 3) for x in c) == 3]
    Esec = np.linalg.eigvalsh(Hd[np.ix_(sector, sector)])[0]
    assert abs(E - Esec) < 1e-8
    assert check_state_symmetric(psi)
    g = path_graph(10)
    d, ops, st = spin_ops("S=1")
    H = ttno(heisenberg_opsum(g), g, ops)
    idx = neel(g, st)
    trunc = dict(cutoff=1e-12, maxdim=[10, 40, 80, 160])
    Eq, _ = dmrg(H, product_ttn_qn(g, "S=1", 3, idx), nsweeps=4, nsites=2, inserter_kwargs=dict(trunc=trunc))
    Ed, _ = dmrg(H, product_ttn(g, d, idx), nsweeps=4, nsites=2, inserter_kwargs=dict(trunc=trunc))
    assert abs(Eq - Ed) < 1e-9


@pytest.mark.parametrize("graph,region", [(path_graph(8), [4, 5]), (path_graph(8), [3]), (named_comb_tree([2, 3, 2]), [(2, 1)])])
def test_environments_of_an_orthonormal_state_have_one_identity_channel(graph, region):
    """Basis of the device's identity-channel skipping (csrc/net.cu, Net::prepare_identity_skip): between orthonormal
    bases every environment of the operator network has exactly one operator-link channel that is the identity
    (the pass-through channel of the sum-of-products operator), so contracting it returns the local tensor itself.
    Checked on the oracle's environments; replacing that channel by an exact identity leaves H_eff theta unchanged."""
    from oracle.gauge import orthogonalize
    from oracle.operator_map import optimal_map
    from oracle.projttn import ProjTTN, position
    from oracle.tensor import contract
    d, ops, _ = spin_ops("S=1/2")
    H = ttno(heisenberg_opsum(graph), graph, ops)
    psi = orthogonalize(random_ttn(graph, d, 6, seed=3), region)
    P = position(ProjTTN(H), psi, region)
    theta = psi.tensors[region[0]]
    for v in region[1:]:
        theta = contract(theta, psi.tensors[v])
    ref = optimal_map(P, theta)
    assert len(P.environments) >= 1
    for key, E in list(P.environments.items()):
        ol = [l for l in E.labels if l[0] == "m"]
        assert len(ol) == 1
        links = [l for l in E.labels if l[0] == "l"]
        A = E.array([links[0], ol[0], links[1]])
        n = A.shape[0]
        dev = [np.abs(A[:, w, :] - np.eye(n)).max() for w in range(A.shape[1])]
        ident = [w for w, x in enumerate(dev) if x < 1e-12]
        assert len(ident) == 1, (key, dev)
        A2 = A.copy()
        A2[:, ident[0], :] = np.eye(n)
        P.environments[key] = type(E)(A2, [links[0], ol[0], links[1]])
    out = optimal_map(P, theta)
    assert np.abs(out.array(ref.labels) - ref.array(ref.labels)).max() < 1e-13 * np.abs(ref.array(ref.labels)).max()
