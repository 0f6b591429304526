"""GPU parity of the sweep hot path against the oracle, through the reference-shaped Python API which
forwards to the C ABI (nsb_extract / nsb_update_* / nsb_insert / nsb_matvec_host).

Tolerances (BASELINE.json north_star): energies 1e-10 relative, truncation errors 1e-8, TDVP fidelity
>= 1 - 1e-8; matvec max rel err <= 1e-13 sqrt(K)."""
import numpy as np
import pytest

from helpers import SweepRecorder, neel, oracle_array, to_oracle_ttn

pytestmark = pytest.mark.gpu


def _ns():
    import networksolvers_b200 as ns
    return ns


def _oracle_sweeps(H, psi0, **kw):
    """Run the oracle DMRG recording per-sweep energies / per-region truncation errors."""
    from oracle import sweep as osw
    rec = {"E": [], "terr": [], "maxdim": []}
    osw.COUNTERS.clear()

    def sweep_cb(region_iter, **k):
        rec["E"].append(region_iter.problem.eigenvalue)
        rec["maxdim"].append(region_iter.problem.state.maxlinkdim())

    E, psi = osw.dmrg(H, psi0, sweep_callback=sweep_cb, **kw)
    rec["terr"] = list(osw.COUNTERS.get("truncerrs", []))
    return E, psi, rec


def _oracle_problem_from_device(net, Ho):
    """Oracle ProjTTN + state built from the tensors currently on the device."""
    from oracle.projttn import ProjTTN
    host = net.to_host()
    return to_oracle_ttn(host), ProjTTN(Ho)


@pytest.mark.parametrize("graph_kind,chi,region", [("chain", 16, (4, 5)), ("chain", 16, (6, 5)), ("chain", 12, (1, 2)),
                                                   ("chain", 12, (10, 9)), ("tree", 6, ((0, 0), (1, 1))),
                                                   ("tree", 6, ((2, 2), (2, 1)))])
@pytest.mark.parametrize("cplx", [False, True])
def test_matvec_matches_oracle(graph_kind, chi, region, cplx):
    ns = _ns()
    from oracle.operator_map import optimal_map, operator_map
    from oracle.projttn import position
    from oracle.tensor import Tensor
    g = ns.path_graph(10) if graph_kind == "chain" else ns.star_of_chains(3, 2)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi = ns.random_state(sites, chi, seed=5, dtype=complex if cplx else float)
    prob = ns.EigsolveProblem(state=psi, operator=H)
    net = prob.net
    net.extract(list(region))          # includes the initial gauge walk (QR steps may permute)
    net.ctx.reset_counters()
    legs, dims = net.local_info()
    theta, _ = net.local_download()
    out = net.matvec_host(theta)
    # oracle on the gauge-moved tensors that are on the device now
    Ho = to_oracle_ttn(H, operator=True)
    psio, P = _oracle_problem_from_device(net, Ho)
    P = position(P, psio, list(region))
    from helpers import _olabel
    th = Tensor(np.array(theta), [_olabel(l) for l in legs])
    ref = optimal_map(P, th).array(th.labels)
    ref2 = operator_map(P, th).array(th.labels)
    K = max(dims) * 5
    tol = 1e-13 * np.sqrt(K) * np.abs(ref).max()
    assert np.abs(ref - ref2).max() <= tol
    assert np.abs(out - ref).max() <= tol, np.abs(out - ref).max()
    if graph_kind == "chain":
        assert net.ctx.counters()["permute_bytes"] == 0, "chain hot path must be permutation-free"


def test_dmrg_reference_example_s1_n10():
    """examples/dmrg.jl:10-43 shape (2-site): per-sweep energies vs oracle to 1e-10, final vs -12.8945601."""
    ns = _ns()
    from oracle.models import spin_ops, heisenberg_opsum, ttno as ottno, product_ttn
    from oracle.graph import path_graph as opath
    g = ns.path_graph(10)
    sites = ns.siteinds("S=1", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-12, maxdim=[10, 40, 80, 160])
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc), sweep_callback=rec.sweep,
                     region_callback=rec.region)
    Eo, _, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=5, nsites=2,
                                 inserter_kwargs=dict(trunc=trunc))
    assert abs(E - (-12.8945601)) < 5e-8
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-10 * abs(b), (rec.energies, orec["E"])
    assert rec.maxlinkdims == orec["maxdim"]
    terr = [t for t in rec.truncerrs if t is not None]
    assert len(terr) == len(orec["terr"])
    assert np.abs(np.array(terr) - np.array(orec["terr"])).max() <= 1e-8


def test_dmrg_config1_heisenberg_n20():
    """BASELINE config 1: S=1/2 N=20, 2-site, maxdim 100, cutoff 1e-12 (SVD route)."""
    ns = _ns()
    g = ns.path_graph(20)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-12, maxdim=100)
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=4, nsites=2, inserter_kwargs=dict(trunc=trunc), sweep_callback=rec.sweep)
    Eo, _, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=4, nsites=2,
                                 inserter_kwargs=dict(trunc=trunc))
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-10 * abs(b), (rec.energies, orec["E"])
    assert abs(E - (-8.682473334399)) < 1e-8
    assert psi.maxlinkdim() <= 100


def test_dmrg_eigen_route_cutoff():
    """cutoff 1e-9 > 1e-12 takes the reference's "eigen" factorize branch (examples/timed_dmrg/timed_dmrg.jl:29)."""
    ns = _ns()
    g = ns.path_graph(12)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-9, maxdim=[10, 40])
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=3, nsites=2, inserter_kwargs=dict(trunc=trunc), sweep_callback=rec.sweep,
                     region_callback=rec.region)
    Eo, _, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=3, nsites=2,
                                 inserter_kwargs=dict(trunc=trunc))
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-10 * abs(b)
    terr = [t for t in rec.truncerrs if t is not None]
    assert np.abs(np.array(terr) - np.array(orec["terr"])).max() <= 1e-8
    assert rec.maxlinkdims == orec["maxdim"]


def test_tree_dmrg_two_site_and_one_site_expansion():
    """test/dmrg/test_tree_dmrg.jl:15-67: |E - E_ED| < 1e-5 (reference's own bar) and 1e-10 vs the oracle."""
    ns = _ns()
    g = ns.star_of_chains(3, 3)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    Ex = -4.046057854359
    trunc = dict(cutoff=1e-5, maxdim=40)
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc), sweep_callback=rec.sweep)
    assert abs(E - Ex) < 1e-5
    Eo, _, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=5, nsites=2,
                                 inserter_kwargs=dict(trunc=trunc))
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-10 * abs(b), (rec.energies, orec["E"])
    rec = SweepRecorder()
    ek = dict(trunc=trunc, subspace_algorithm="densitymatrix")
    E, psi = ns.dmrg(H, psi0, nsweeps=5, nsites=1, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc),
                     sweep_callback=rec.sweep)
    assert abs(E - Ex) < 1e-5
    Eo, _, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=5, nsites=1, extracter_kwargs=ek,
                                 inserter_kwargs=dict(trunc=trunc))
    assert rec.maxlinkdims == orec["maxdim"]
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-9 * abs(b), (rec.energies, orec["E"])


def test_one_site_expansion_chain_example_settings():
    """examples/dmrg.jl:26-36: 1-site + densitymatrix expansion (expansion_factor 1.1) on the S=1 chain."""
    ns = _ns()
    g = ns.path_graph(10)
    sites = ns.siteinds("S=1", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-12, maxdim=[10, 40, 80, 160])
    ek = dict(trunc=trunc, subspace_algorithm="densitymatrix", expansion_factor=1.1)
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=5, nsites=1, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc),
                     sweep_callback=rec.sweep)
    Eo, _, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=5, nsites=1, extracter_kwargs=ek,
                                 inserter_kwargs=dict(trunc=trunc))
    assert rec.maxlinkdims == orec["maxdim"]
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-9 * abs(b), (rec.energies, orec["E"])
    assert abs(E - (-12.8945601)) < 1e-6


@pytest.mark.parametrize("solver_name", ["rk4", "krylov"])
@pytest.mark.parametrize("order", [2, 4])
def test_tdvp_two_site_fidelity(order, solver_name):
    """examples/quench_evolution.jl:20-84 shape: fidelity vs dense expm >= 1 - 1e-8, and vs the oracle state."""
    ns = _ns()
    from oracle.ed import ed_time_evolution, state_vector
    from oracle.models import heisenberg_opsum, spin_ops
    from oracle import sweep as osw
    from oracle.local_solvers import runge_kutta_solver as o_rk, exponentiate_solver as o_exp
    g = ns.path_graph(8)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g, even_up=False))
    tp = list(np.arange(0, 0.3 + 1e-9, 0.05))
    ik = dict(trunc=dict(maxdim=5000, cutoff=1e-14), normalize=True)
    uk = dict(solver=ns.runge_kutta_solver, order=4) if solver_name == "rk4" else dict(solver=ns.exponentiate_solver)
    psit = ns.tdvp(H, psi0, tp, nsites=2, tdvp_order=order, updater_kwargs=uk, inserter_kwargs=ik)
    v = psit.to_host().to_dense()
    og = to_oracle_ttn(psi0).graph
    d, ops, _ = spin_ops("S=1/2")
    vx = ed_time_evolution(heisenberg_opsum(og), og, ops, psi0.to_dense(), tp, normalize=True)
    assert 1 - abs(np.vdot(vx, v)) < 1e-8
    ouk = dict(solver=o_rk, order=4) if solver_name == "rk4" else dict(solver=o_exp)
    po = osw.tdvp(to_oracle_ttn(H, True), to_oracle_ttn(psi0), tp, nsites=2, tdvp_order=order, updater_kwargs=ouk,
                  inserter_kwargs=ik)
    vo = state_vector(po)
    assert 1 - abs(np.vdot(vo, v)) < 1e-10
    assert psit.maxlinkdim() == po.maxlinkdim()


def test_error_behaviour():
    """Error paths mirror the reference: unsupported region length (src/inserter.jl:26), unknown expansion
    backend (src/subspace/subspace.jl:16-20), bad RK order (src/local_solvers/runge_kutta.jl:22)."""
    ns = _ns()
    g = ns.path_graph(6)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    prob = ns.EigsolveProblem(state=psi0, operator=H)
    with pytest.raises(ns.NsbError) as ei:
        prob.net.extract([1, 2, 3])
    assert ei.value.code == -6
    with pytest.raises(ValueError):
        ns.dmrg(H, psi0, nsweeps=1, nsites=1, extracter_kwargs=dict(subspace_algorithm="nonsense"))
    with pytest.raises(ValueError):
        ns.tdvp(H, psi0, [0.0, 0.1], nsites=2, updater_kwargs=dict(solver=ns.runge_kutta_solver, order=3))
    with pytest.raises(ValueError):
        ns.tdvp_sub_time_steps(3)


@pytest.mark.parametrize("cplx", [False, True])
def test_chain_region_steps_are_permutation_free(cplx):
    """After the initial gauge walk, extract (theta build + environment update), the Lanczos update and the
    insert of consecutive 2-site chain regions move no data through layout permutes, in both directions."""
    ns = _ns()
    g = ns.path_graph(10)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi = ns.random_state(sites, 16, seed=5, dtype=complex if cplx else float)
    net = ns.EigsolveProblem(state=psi, operator=H).net
    net.extract([4, 5]); net.update_eigsolve(); net.insert((0.0, 1, 16))
    net.ctx.reset_counters()
    for region in ([5, 6], [6, 7], [7, 6], [6, 5], [5, 4]):
        info = net.extract(region)
        assert info.qr_steps == 0 and info.env_builds <= 1
        val, sinfo = net.update_eigsolve()
        assert sinfo.nmatvec == 3
        net.insert((0.0, 1, 16))
    c = net.ctx.counters()
    assert c["permute_bytes"] == 0
    assert c["matvecs"] == 15


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("nsites", [2, 1])
def test_identity_channel_skipping(cplx, nsites):
    """ctx option skip_identity: the channel of the left / right environment that is the identity (orthonormal bases,
    MPO with a pass-through channel) is not contracted -- theta is spliced in instead.  Same H_eff theta as the dense
    contraction to rounding, fewer flops issued, in both sweep directions; a non-canonical state has no such channel."""
    ns = _ns()
    g = ns.path_graph(10)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi = ns.random_state(sites, 24, seed=11, dtype=complex if cplx else float)
    net = ns.EigsolveProblem(state=psi, operator=H).net
    ctx = net.ctx
    regions = ([4, 5], [5, 6], [6, 5], [5, 4]) if nsites == 2 else ([4], [5], [6], [5], [4])
    try:
        for region in regions:
            outs, flops = {}, {}
            for skip in (1, 0):
                ctx.set_option("skip_identity", skip)
                net.extract(region)          # (re-positions: the compact copy of the first environment follows the option)
                theta, _ = net.local_download()
                outs[skip] = net.matvec_host(theta)
                flops[skip] = (net.matvec_flops(), net.matvec_flops_executed())
            err = np.abs(outs[1] - outs[0]).max() / np.abs(outs[0]).max()
            assert err < 1e-13, (region, err)
            assert flops[0][0] == flops[0][1] == flops[1][0]
            assert flops[1][1] < 0.9 * flops[1][0], (region, flops)      # one of w = 5 channels on each side
            ctx.set_option("skip_identity", 1)
            net.extract(region); net.update_eigsolve(); net.insert((0.0, 1, 24))
    finally:
        ctx.set_option("skip_identity", 1)


def test_sharded_matvec_two_gpus():
    """theta sharded along its last bond over 2 ranks + NCCL all-reduce == single-GPU matvec."""
    import os, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(root, "tests", "_nccl_shard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_SHARD_OK 2" in r.stdout


def test_tree_tdvp_one_site_and_two_site_regression():
    """test/tdvp/test_tree_tdvp.jl:24-77 (chain + ancilla tree): 1-site TDVP (QR split + on-edge backward step,
    src/applyexp.jl:30-42) and 2-site TDVP from the DMRG ground state; the reference's own asserts plus parity
    with the oracle states."""
    ns = _ns()
    from oracle import sweep as osw
    from oracle.ed import state_vector
    N = 10
    g = ns.NamedGraph()
    for j in range(1, N + 1):
        g.add_vertex(j)
    for j in range(1, N):
        g.add_edge(j, j + 1)
    g.add_vertex(0)
    g.add_edge(0, N // 2)
    sites = ns.siteinds("S=1/2", g)
    os_ = ns.OpSum()
    for j in range(1, N):
        os_.add(1.0, "Sz", j, "Sz", j + 1)
        os_.add(0.5, "S+", j, "S-", j + 1)
        os_.add(0.5, "S-", j, "S+", j + 1)
    H = ns.ttno(os_, sites)
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-10, maxdim=100)
    E, gs = ns.dmrg(H, psi0, nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc))
    gs_host = gs.to_host()
    tmax = 0.10
    tp = list(np.arange(0, tmax + 1e-9, 0.02))
    psi1 = ns.tdvp(H, gs_host, tp, nsites=1, inserter_kwargs=dict(trunc=trunc))
    psi2 = ns.tdvp(H, gs_host, tp, nsites=2, inserter_kwargs=dict(trunc=trunc))
    v0, v1, v2 = gs_host.to_dense(), psi1.to_host().to_dense(), psi2.to_host().to_dense()
    assert np.linalg.norm(v1) > 0.999 and np.linalg.norm(v2) > 0.999
    assert abs(np.vdot(v1, v0)) > 0.99
    assert abs(np.vdot(v1, v2)) > 0.99
    z = np.vdot(v1, v0)
    assert abs(np.arctan(z.imag / z.real) - E * tmax) < 1e-4
    # parity with the oracle evolving the same initial state
    Ho, gso = to_oracle_ttn(H, True), to_oracle_ttn(gs_host)
    o1 = state_vector(osw.tdvp(Ho, gso, tp, nsites=1, inserter_kwargs=dict(trunc=trunc)))
    o2 = state_vector(osw.tdvp(Ho, gso, tp, nsites=2, inserter_kwargs=dict(trunc=trunc)))
    assert 1 - abs(np.vdot(o1, v1)) / (np.linalg.norm(o1) * np.linalg.norm(v1)) < 1e-8
    assert 1 - abs(np.vdot(o2, v2)) / (np.linalg.norm(o2) * np.linalg.norm(v2)) < 1e-8
    assert abs(np.linalg.norm(o1) - np.linalg.norm(v1)) < 1e-8


def test_one_site_tdvp_chain_krylov_solver():
    """1-site TDVP with the Krylov exponentiate solver on a chain, vs the oracle."""
    ns = _ns()
    from oracle import sweep as osw
    from oracle.ed import state_vector
    from oracle.local_solvers import exponentiate_solver as o_exp
    g = ns.path_graph(6)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi = ns.random_state(sites, 8, seed=3, dtype=complex)
    tp = [0.0, 0.05, 0.1]
    out = ns.tdvp(H, psi, tp, nsites=1, tdvp_order=2, updater_kwargs=dict(solver=ns.exponentiate_solver))
    v = out.to_host().to_dense()
    vo = state_vector(osw.tdvp(to_oracle_ttn(H, True), to_oracle_ttn(psi), tp, nsites=1, tdvp_order=2,
                               updater_kwargs=dict(solver=o_exp)))
    assert 1 - abs(np.vdot(vo, v)) / (np.linalg.norm(vo) * np.linalg.norm(v)) < 1e-9
    assert abs(np.linalg.norm(vo) - np.linalg.norm(v)) < 1e-9 * np.linalg.norm(vo)


def test_dmrg_hubbard_dense_matches_oracle_and_ed():
    """BASELINE config 3 Hamiltonian (Hubbard chain, d = 4, w = 6) in dense (no-QN) form, 2-site DMRG with
    density-matrix expansion.  Region energies equal the oracle's to 1e-10 until the first truncation that cuts
    through a symmetry-degenerate multiplet (there the kept basis is not unique: LAPACK and Jacobi pick different,
    equally valid vectors, and a dense run -- oracle and device alike -- eventually leaks out of the particle-number
    sector of the start state through rounding noise, which is why the reference's own examples use QN-conserving
    tensors for such models).  The energy is variational and ends at the ground state of the start sector or,
    after leaking, of a lower sector."""
    import itertools
    ns = _ns()
    g = ns.path_graph(6)
    sites = ns.siteinds("Electron", g)
    H = ns.ttno(ns.hubbard(g, 1.0, 4.0), sites)
    psi0 = ns.product_state(sites, {v: ("Up" if v % 2 else "Dn") for v in g.vertices})
    trunc = dict(cutoff=1e-10, maxdim=[10, 20, 60])
    ek = dict(trunc=trunc, subspace_algorithm="densitymatrix", expansion_factor=1.1)
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=8, nsites=2, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc),
                     sweep_callback=rec.sweep, region_callback=rec.region)
    from oracle import sweep as osw
    oreg = []
    osw.dmrg(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=1, nsites=2, extracter_kwargs=ek,
             inserter_kwargs=dict(trunc=trunc), region_callback=lambda p, **k: oreg.append(p.eigenvalue))
    for a, b in list(zip(rec.region_energies, oreg))[:3]:
        assert abs(a - b) <= 1e-10 * abs(b), (rec.region_energies[:5], oreg[:5])
    from oracle.ed import ttno_dense
    Hd = ttno_dense(to_oracle_ttn(H, True), to_oracle_ttn(psi0).graph, 4)
    sector = [i for i, conf in enumerate(itertools.product(range(4), repeat=6))
              if sum(c in (1, 3) for c in conf) == 3 and sum(c in (2, 3) for c in conf) == 3]
    Esec = np.linalg.eigvalsh(Hd[np.ix_(sector, sector)])[0]
    w = np.linalg.eigvalsh(Hd)
    assert E >= w[0] - 1e-9                                  # variational
    assert min(abs(rec.energies[2] - Esec), abs(rec.energies[2] - w[0])) < 1e-3 or E < Esec
    assert all(b <= a + 1e-9 for a, b in zip(rec.energies[:-1], rec.energies[1:]))   # monotone over sweeps


def test_tdvp_two_site_with_expansion_quench_example():
    """examples/quench_evolution.jl:20-57 exactly: 2-site TDVP, tdvp_order 4, RK4, densitymatrix expansion
    (expansion_factor 1.2, max_expand 4), cutoff 1e-14, normalize -- complex arithmetic through the expansion path."""
    ns = _ns()
    from oracle.ed import ed_time_evolution, state_vector
    from oracle.models import heisenberg_opsum, spin_ops
    from oracle import sweep as osw
    from oracle.local_solvers import runge_kutta_solver as o_rk
    g = ns.path_graph(8)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g, even_up=False))
    tp = list(np.arange(0, 0.2 + 1e-9, 0.05))
    ek = dict(subspace_algorithm="densitymatrix", expansion_factor=1.2, max_expand=4)
    ik = dict(trunc=dict(maxdim=5000, cutoff=1e-14), normalize=True)
    psit = ns.tdvp(H, psi0, tp, nsites=2, tdvp_order=4, extracter_kwargs=ek, updater_kwargs=dict(solver=ns.runge_kutta_solver, order=4),
                   inserter_kwargs=ik)
    v = psit.to_host().to_dense()
    og = to_oracle_ttn(psi0).graph
    d, ops, _ = spin_ops("S=1/2")
    vx = ed_time_evolution(heisenberg_opsum(og), og, ops, psi0.to_dense(), tp, normalize=True)
    assert 1 - abs(np.vdot(vx, v)) < 1e-8
    po = osw.tdvp(to_oracle_ttn(H, True), to_oracle_ttn(psi0), tp, nsites=2, tdvp_order=4, extracter_kwargs=ek,
                  updater_kwargs=dict(solver=o_rk, order=4), inserter_kwargs=ik)
    vo = state_vector(po)
    assert 1 - abs(np.vdot(vo, v)) < 1e-9


def test_matvec_host_slab_equals_matvec_host_when_unsharded():
    """nsb_matvec_host_slab / nsb_shard_range on one GPU: the slab is the whole last mode and the call is nsb_matvec_host
    (the 2-GPU form of the same call is checked by tests/_nccl_shard_worker.py)."""
    ns = _ns()
    g = ns.path_graph(10)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi = ns.random_state(sites, 24, seed=3)
    net = ns.EigsolveProblem(state=psi, operator=H).net
    net.extract([5, 6])
    _, dims = net.local_info()
    assert net.shard_range() == (0, dims[-1], dims[-1])
    th, _ = net.local_download()
    assert np.array_equal(net.matvec_host_slab(th), net.matvec_host(th))


def test_dmrg_with_compressed_long_range_ttno():
    """A Hamiltonian the nearest-neighbour builder cannot express (J1-J2 chain, N = 12) through the general OpSum -> compressed
    TTNO construction: per-sweep energies equal the oracle's on the same operator tensors (1e-10), final energy = ED."""
    ns = _ns()
    g = ns.path_graph(12)
    V = g.vertices
    sites = ns.siteinds("S=1/2", g)
    os_ = ns.OpSum()
    for r, J in ((1, 1.0), (2, 0.35)):
        for i in range(len(V) - r):
            os_.add(J, "Sz", V[i], "Sz", V[i + r])
            os_.add(J / 2, "S+", V[i], "S-", V[i + r])
            os_.add(J / 2, "S-", V[i], "S+", V[i + r])
    H = ns.ttno(os_, sites)
    assert H.maxlinkdim() == 8
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-12, maxdim=64)
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc), sweep_callback=rec.sweep)
    Eo, _, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=5, nsites=2, inserter_kwargs=dict(trunc=trunc))
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-10 * abs(b), (rec.energies, orec["E"])
    from functools import reduce
    op, d = sites.type.op, 2
    Hd = np.zeros((d ** 12,) * 2)
    for term in os_.terms:
        mats = {v: np.asarray(op(nm)).real for nm, v in zip(term[1::2], term[2::2])}
        Hd += term[0] * reduce(np.kron, [mats.get(v, np.eye(d)) for v in V])
    assert abs(E - np.linalg.eigvalsh(Hd)[0]) < 1e-8
