"""QN-conserving runs (BASELINE config 3 shape: Hubbard chain with (Nf, Sz) conservation + density-matrix expansion)
against the QN-aware oracle and exact diagonalisation inside the symmetry sector."""
import itertools

import numpy as np
import pytest

from helpers import SweepRecorder, neel, to_oracle_ttn, to_oracle_ttn_qn

pytestmark = pytest.mark.gpu


def _oracle_sweeps(H, psi0, **kw):
    from oracle import sweep as osw
    rec = {"E": [], "maxdim": []}
    osw.COUNTERS.clear()

    def sweep_cb(region_iter, **k):
        rec["E"].append(region_iter.problem.eigenvalue)
        rec["maxdim"].append(region_iter.problem.state.maxlinkdim())

    E, psi = osw.dmrg(H, psi0, sweep_callback=sweep_cb, **kw)
    rec["terr"] = list(osw.COUNTERS.get("truncerrs", []))
    return E, psi, rec


def test_hubbard_qn_dmrg_with_expansion():
    import networksolvers_b200 as ns
    from oracle.ed import ttno_dense
    from oracle.qn import check_state_symmetric
    g = ns.path_graph(6)
    sites = ns.siteinds("Electron", g, conserve_qns=True)
    H = ns.ttno(ns.hubbard(g, 1.0, 4.0), sites)
    psi0 = ns.product_state(sites, {v: ("Up" if v % 2 else "Dn") for v in g.vertices})
    assert psi0.qn["total"].tolist() == [6, 0]
    trunc = dict(cutoff=1e-10, maxdim=[10, 20, 60])
    ek = dict(trunc=trunc, subspace_algorithm="densitymatrix", expansion_factor=1.1)
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=6, nsites=2, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc),
                     sweep_callback=rec.sweep, region_callback=rec.region)
    Eo, psio, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn_qn(psi0), nsweeps=6, nsites=2, extracter_kwargs=ek,
                                    inserter_kwargs=dict(trunc=trunc))
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-9 * abs(b), (rec.energies, orec["E"])
    assert rec.maxlinkdims == orec["maxdim"]
    terr = [t for t in rec.truncerrs if t is not None]
    assert np.abs(np.array(terr) - np.array(orec["terr"])).max() <= 1e-8
    # exact ground state of the (N_up, N_dn) = (3, 3) sector: no leakage with QN conservation
    Hd = ttno_dense(to_oracle_ttn(H, True), to_oracle_ttn(psi0).graph, 4)
    sector = [i for i, c in enumerate(itertools.product(range(4), repeat=6))
              if sum(x in (1, 3) for x in c) == 3 and sum(x in (2, 3) for x in c) == 3]
    Esec = np.linalg.eigvalsh(Hd[np.ix_(sector, sector)])[0]
    assert abs(E - Esec) < 1e-8
    final = to_oracle_ttn_qn(psi.to_host())
    assert check_state_symmetric(final, tol=1e-13)          # every site tensor obeys the selection rule exactly
    assert {k: v for k, v in psi.linkdims().items()} == {k: v for k, v in psio.linkdims().items()}


@pytest.mark.parametrize("nsites", [2, 1])
def test_heisenberg_s1_qn_matches_oracle(nsites):
    """examples/dmrg.jl with conserve_qns=true: 2-site, and 1-site + densitymatrix expansion (expansion_factor 1.1)."""
    import networksolvers_b200 as ns
    from oracle.qn import check_state_symmetric
    g = ns.path_graph(10)
    sites = ns.siteinds("S=1", g, conserve_qns=True)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-12, maxdim=[10, 40, 80, 160])
    ek = dict(trunc=trunc, subspace_algorithm="densitymatrix", expansion_factor=1.1) if nsites == 1 else {}
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=5, nsites=nsites, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc),
                     sweep_callback=rec.sweep)
    Eo, psio, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn_qn(psi0), nsweeps=5, nsites=nsites,
                                    extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc))
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-9 * abs(b), (rec.energies, orec["E"])
    assert rec.maxlinkdims == orec["maxdim"]
    assert abs(E - (-12.8945601)) < 1e-6
    assert check_state_symmetric(to_oracle_ttn_qn(psi.to_host()), tol=1e-13)


def test_tdvp_qn_two_site_fidelity():
    import networksolvers_b200 as ns
    from oracle.ed import ed_time_evolution
    from oracle.models import heisenberg_opsum, spin_ops
    from oracle.qn import check_state_symmetric
    g = ns.path_graph(8)
    sites = ns.siteinds("S=1/2", g, conserve_qns=True)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g, even_up=False))
    tp = list(np.arange(0, 0.3 + 1e-9, 0.05))
    ik = dict(trunc=dict(maxdim=5000, cutoff=1e-14), normalize=True)
    psit = ns.tdvp(H, psi0, tp, nsites=2, tdvp_order=2, updater_kwargs=dict(solver=ns.runge_kutta_solver, order=4), inserter_kwargs=ik)
    host = psit.to_host()
    og = to_oracle_ttn(psi0).graph
    d, ops, _ = spin_ops("S=1/2")
    vx = ed_time_evolution(heisenberg_opsum(og), og, ops, psi0.to_dense(), tp, normalize=True)
    assert 1 - abs(np.vdot(vx, host.to_dense())) < 1e-8
    assert check_state_symmetric(to_oracle_ttn_qn(host), tol=1e-13)
