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


@pytest.mark.parametrize("model", ["hubbard", "s1"])
def test_block_sparse_engine_matches_dense_storage(model):
    """K13: environments, local tensor and Krylov vectors as symmetry blocks with grouped sector GEMMs (ctx option
    qn_block_sparse, default on) against the dense-storage path of the same library on the same state: H_eff application to
    1e-13, Ritz value / kept dimension / truncation error of a whole region step, both sweep directions; the sector GEMMs
    execute a fraction of the dense-equivalent flops."""
    import networksolvers_b200 as ns
    g = ns.path_graph(10)
    if model == "hubbard":
        sites = ns.siteinds("Electron", g, conserve_qns=True)
        H = ns.ttno(ns.hubbard(g, 1.0, 4.0), sites)
        psi0 = ns.product_state(sites, {v: ("Up" if v % 2 else "Dn") for v in g.vertices})
    else:
        sites = ns.siteinds("S=1", g, conserve_qns=True)
        H = ns.ttno(ns.heisenberg(g), sites)
        psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-12, maxdim=[8, 16, 48])
    ctx = ns.default_context()
    prob = ns.EigsolveProblem(state=psi0, operator=H)
    ns.dmrg(prob, nsweeps=3, nsites=2, inserter_kwargs=dict(trunc=trunc))       # grows the bonds (block-sparse path)
    host = prob.net.to_host()
    res = {}
    try:
        for bs in (1, 0):
            ctx.set_option("qn_block_sparse", bs)
            net = ns.EigsolveProblem(state=host, operator=H).net
            out = []
            for reg in ([5, 6], [6, 7], [7, 6], [6, 5]):
                net.extract(reg)
                th, _ = net.local_download()
                hv = net.matvec_host(th)
                fl = (net.matvec_flops_executed(), net.matvec_flops())
                if bs == 1:
                    # the same local tensor through dense-storage environments of the same network
                    ctx.set_option("qn_block_sparse", 0)
                    net.env_drop_all()
                    net.extract(reg)
                    hv_dense = net.matvec_host(th)
                    ctx.set_option("qn_block_sparse", 1)
                    net.env_drop_all()
                    net.extract(reg)
                    assert np.abs(hv - hv_dense).max() <= 1e-13 * np.abs(hv_dense).max() * 10, reg
                val, info = net.update_eigsolve()
                th2, _ = net.local_download()
                ins = net.insert((1e-12, 1, 48))
                ev = np.vdot(th, hv).real / np.vdot(th, th).real
                out.append((ev, fl, val, np.linalg.norm(th2), ins.newdim, ins.truncerr, info.nmatvec))
            res[bs] = out
    finally:
        ctx.set_option("qn_block_sparse", 1)
    # gauge-invariant quantities of the two storage forms (degenerate multiplets may rotate inside their subspace)
    for a, b in zip(res[1], res[0]):
        assert abs(a[0] - b[0]) <= 1e-11 * max(1.0, abs(b[0]))
        assert abs(a[2] - b[2]) <= 1e-11 * max(1.0, abs(b[2]))
        assert abs(a[3] - 1.0) <= 1e-12 and a[6] == b[6] == 3
        assert a[4] == b[4] and abs(a[5] - b[5]) <= 1e-11
        assert a[1][1] == b[1][1]                                     # same dense-equivalent count
        assert a[1][0] < 0.5 * a[1][1], a[1]                          # the sector GEMMs skip the symmetry-forbidden blocks
