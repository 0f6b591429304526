"""GPU parity of the device Hermitian eigensolver (csrc/eigh.cu: blocked tridiagonalisation + divide & conquer +
WY back-transformation) and of the Gram + eigh route of the truncating factorisation, against NumPy/LAPACK and the
oracle's truncation rule, called through the C ABI (nsb_eigh_host / nsb_factorize_host / the DMRG hooks)."""
import numpy as np
import pytest

from helpers import SweepRecorder, neel, to_oracle_ttn

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def ctx():
    import networksolvers_b200 as ns
    return ns.default_context()


@pytest.fixture
def eigh_route(ctx):
    """Force the Gram + eigh route for every matrix size, restore the default afterwards."""
    ctx.set_option("eigh_min_n", 2)
    yield ctx
    ctx.set_option("eigh_min_n", 1024)
    ctx.set_option("eigh_nb", 64)


def _cases(n, rng):
    M = rng.standard_normal((n, n))
    yield "gauss", M + M.T
    G = rng.standard_normal((n, 2 * n))
    yield "wishart", G @ G.T
    Mc = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    yield "complex", Mc + Mc.conj().T
    if n < 8:
        return
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    yield "graded", (Q * np.exp(-40.0 * np.arange(n) / n)) @ Q.T
    yield "lowrank", (Q[:, :3] * np.array([1.0, 0.5, 1e-3])) @ Q[:, :3].T
    yield "identity+rank1", np.eye(n) + 1e-3 * np.outer(Q[:, 0], Q[:, 0])
    cl = np.repeat(np.arange(1, n // 8 + 2), 8)[:n].astype(float)
    yield "clustered", (Q * cl) @ Q.T
    Gc = rng.standard_normal((n, n + 3)) + 1j * rng.standard_normal((n, n + 3))
    yield "complex gram", Gc @ Gc.conj().T


@pytest.mark.parametrize("coop", [1, 0])
@pytest.mark.parametrize("n,nb", [(1, 64), (2, 64), (5, 64), (64, 64), (129, 64), (300, 32), (700, 64), (1100, 128)])
def test_eigh_matches_lapack(ctx, n, nb, coop):
    """coop = 1: one cooperative kernel per tridiagonalisation panel; coop = 0: five launches per column."""
    rng = np.random.default_rng(n)
    ctx.set_option("eigh_nb", nb)
    ctx.set_option("eigh_coop", coop)
    try:
        for name, A in _cases(n, rng):
            w, U = ctx.eigh(A)
            nrm = max(np.linalg.norm(A, 2), 1e-300)
            res = np.linalg.norm(A @ U - U * w[None, :]) / (nrm * n)
            orth = np.linalg.norm(U.conj().T @ U - np.eye(n)) / n
            err = np.abs(w - np.linalg.eigvalsh(A)).max() / nrm
            assert res < 30 * EPS and orth < 30 * EPS and err < 100 * EPS, (name, n, res, orth, err)
    finally:
        ctx.set_option("eigh_nb", 64)
        ctx.set_option("eigh_coop", 1)


@pytest.mark.parametrize("tc", [0, 16, 32, 64, 128])
@pytest.mark.parametrize("n,nb", [(4, 64), (66, 64), (300, 32), (1100, 128), (1538, 64), (2600, 64)])
def test_eigh_symmetric_panel_kernel(ctx, n, nb, tc):
    """Real FP64, even n: the half-traffic panel kernel (lower-triangle work units, per-unit partials summed in the next
    phase) for every unit width, against LAPACK and against the full-square kernel."""
    rng = np.random.default_rng(7 * n + tc)
    ctx.set_option("eigh_nb", nb)
    ctx.set_option("eigh_sym_tc", tc)
    try:
        for name, A in _cases(n, rng):
            if np.iscomplexobj(A) or (n > 2000 and name != "gauss"):
                continue
            ctx.set_option("eigh_sym", 1)
            w, U = ctx.eigh(A)
            if name == "gauss":   # the symmetric kernel and its lower-triangular trailing update never read the upper triangle
                Ap = np.tril(A) + np.triu(np.full((n, n), np.nan), 1)
                assert np.array_equal(ctx.eigh(Ap, vectors=False), w), (name, n, tc)
            ctx.set_option("eigh_sym", 0)
            w0 = ctx.eigh(A, vectors=False)
            nrm = max(np.linalg.norm(A, 2), 1e-300)
            res = np.linalg.norm(A @ U - U * w[None, :]) / (nrm * n)
            orth = np.linalg.norm(U.T @ U - np.eye(n)) / n
            err = np.abs(w - np.linalg.eigvalsh(A)).max() / nrm
            err0 = np.abs(w - w0).max() / nrm
            tol = 100 * EPS * max(1.0, n / 1000)   # eigenvalue error bound grows with n
            assert res < 30 * EPS and orth < 30 * EPS and err < tol and err0 < tol, (name, n, tc, res, orth, err, err0)
    finally:
        ctx.set_option("eigh_nb", 64)
        ctx.set_option("eigh_sym", 1)
        ctx.set_option("eigh_sym_tc", 0)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(8, 8), (20, 33), (33, 20), (96, 96), (200, 333), (260, 200), (640, 640)])
def test_factorize_eigh_route_full_spectrum(eigh_route, cplx, shape):
    """Eigen route, no truncation: spectrum = sigma^2 of LAPACK (absolute accuracy eps sigma_1^2), U orthonormal,
    U C = M."""
    rng = np.random.default_rng(23)
    M = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0.0)
    U, Cm, spec, info = eigh_route.factorize(M, cutoff=0.0)
    s = np.linalg.svd(M, compute_uv=False)
    k = min(shape)
    assert info["newdim"] == k
    assert np.abs(spec - s**2).max() <= 1e-12 * s[0] ** 2
    assert np.abs(U.conj().T @ U - np.eye(k)).max() < 1e-12
    assert np.abs(U @ Cm - M).max() < 1e-11 * max(1.0, s[0])


@pytest.mark.parametrize("cutoff,maxdim", [(1e-12, None), (1e-8, None), (1e-4, None), (0.0, 10), (1e-6, 7)])
def test_factorize_eigh_route_truncation_rule(eigh_route, cutoff, maxdim):
    """Truncation rule (NDTensors truncate!) vs the oracle on a decaying spectrum."""
    from oracle.tensor import truncate_spectrum
    rng = np.random.default_rng(17)
    n = 160
    Uo, _ = np.linalg.qr(rng.standard_normal((n, n)))
    Vo, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sig = np.exp(-0.2 * np.arange(n))
    M = (Uo * sig) @ Vo.T
    U, Cm, spec, info = eigh_route.factorize(M, cutoff=cutoff, maxdim=maxdim)
    nk, terr = truncate_spectrum(np.maximum(np.linalg.eigvalsh(M @ M.T)[::-1], 0.0), cutoff=cutoff, mindim=1, maxdim=maxdim)
    assert info["newdim"] == nk
    assert abs(info["truncerr"] - terr) <= 1e-8 * max(terr, 1e-30) + 1e-15
    best = (Uo[:, :nk] * sig[:nk]) @ Vo[:, :nk].T
    assert np.abs(U @ Cm - best).max() < 1e-7 * max(sig[nk - 1], 1e-8) + 1e-10


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


@pytest.mark.parametrize("cutoff", [1e-12, 1e-9])
def test_dmrg_energies_with_eigh_route(eigh_route, cutoff):
    """2-site DMRG (S=1/2 Heisenberg N=14) with every factorisation on the Gram + eigh route: per-sweep energies
    within 1e-10 relative and truncation errors within 1e-8 of the oracle (LAPACK svd / eigh)."""
    import networksolvers_b200 as ns
    g = ns.path_graph(14)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=cutoff, maxdim=[10, 20, 40])
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=3, nsites=2, inserter_kwargs=dict(trunc=trunc), sweep_callback=rec.sweep,
                     region_callback=rec.region)
    Eo, _, orec = _oracle_sweeps(to_oracle_ttn(H, True), to_oracle_ttn(psi0), nsweeps=3, nsites=2,
                                 inserter_kwargs=dict(trunc=trunc))
    for a, b in zip(rec.energies, orec["E"]):
        assert abs(a - b) <= 1e-10 * abs(b), (rec.energies, orec["E"])
    terr = [t for t in rec.truncerrs if t is not None]
    assert np.abs(np.array(terr) - np.array(orec["terr"])).max() <= 1e-8
    assert rec.maxlinkdims == orec["maxdim"]


def test_eigh_large_residual(ctx):
    """n = 2048 Wishart matrix: residual and orthogonality at LAPACK level, timing printed for the log."""
    import time
    rng = np.random.default_rng(3)
    n = 2048
    G = rng.standard_normal((n, n))
    A = G @ G.T
    t0 = time.perf_counter()
    w, U = ctx.eigh(A)
    dt = time.perf_counter() - t0
    nrm = np.linalg.norm(A, 2)
    res = np.linalg.norm(A @ U - U * w[None, :]) / (nrm * n)
    orth = np.linalg.norm(U.T @ U - np.eye(n)) / n
    print(f"eigh n={n}: {dt:.3f} s (incl. H2D/D2H) resid {res:.2e} orth {orth:.2e}")
    assert res < 30 * EPS and orth < 30 * EPS


def test_one_site_expansion_with_eigh_route(eigh_route):
    """1-site DMRG + densitymatrix expansion (examples/dmrg.jl:26-36) with the expansion's eigen(rho) and the gauge
    factorisations on the device eigensolver (direct Hermitian path of factorize_left)."""
    import networksolvers_b200 as ns
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


def test_tdvp_complex_with_eigh_route(eigh_route):
    """2-site TDVP (complex128, examples/quench_evolution.jl shape) with every factorisation on the eigh route:
    fidelity against dense expm >= 1 - 1e-8 and against the oracle state >= 1 - 1e-10."""
    import networksolvers_b200 as ns
    from oracle.ed import ed_time_evolution, state_vector
    from oracle.models import heisenberg_opsum, spin_ops
    from oracle import sweep as osw
    from oracle.local_solvers import runge_kutta_solver as o_rk
    g = ns.path_graph(8)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g, even_up=False))
    tp = list(np.arange(0, 0.2 + 1e-9, 0.05))
    ik = dict(trunc=dict(maxdim=5000, cutoff=1e-14), normalize=True)
    psit = ns.tdvp(H, psi0, tp, nsites=2, tdvp_order=2, updater_kwargs=dict(solver=ns.runge_kutta_solver, order=4), inserter_kwargs=ik)
    v = psit.to_host().to_dense()
    og = to_oracle_ttn(psi0).graph
    d, ops, _ = spin_ops("S=1/2")
    vx = ed_time_evolution(heisenberg_opsum(og), og, ops, psi0.to_dense(), tp, normalize=True)
    assert 1 - abs(np.vdot(vx, v)) < 1e-8
    po = osw.tdvp(to_oracle_ttn(H, True), to_oracle_ttn(psi0), tp, nsites=2, tdvp_order=2, updater_kwargs=dict(solver=o_rk, order=4),
                  inserter_kwargs=ik)
    assert 1 - abs(np.vdot(state_vector(po), v)) < 1e-10


def test_ortho_expansion_backend(ctx):
    """`subspace_algorithm="ortho"` (src/subspace/ortho_subspace.jl:19-77): random directions orthogonal to the basis
    of the previous vertex.  The random numbers differ from the oracle's, so the checks are the properties the method
    guarantees: the state is unchanged by an expansion, the enlarged basis stays orthonormal, the bond grows by
    compute_expansion, and 1-site DMRG with it reaches the exact ground-state energy (which 1-site DMRG without
    expansion cannot, starting from a product state)."""
    import networksolvers_b200 as ns
    from oracle.ed import ed_ground_state
    from oracle.models import heisenberg_opsum, spin_ops
    g = ns.path_graph(10)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi0 = ns.product_state(sites, neel(g))
    trunc = dict(cutoff=1e-12, maxdim=[4, 8, 16, 32, 32, 32])
    ek = dict(trunc=trunc, subspace_algorithm="ortho", expansion_factor=1.5)
    rec = SweepRecorder()
    E, psi = ns.dmrg(H, psi0, nsweeps=6, nsites=1, extracter_kwargs=ek, inserter_kwargs=dict(trunc=trunc), sweep_callback=rec.sweep)
    og = to_oracle_ttn(psi0).graph
    d, ops, _ = spin_ops("S=1/2")
    E0 = float(np.atleast_1d(ed_ground_state(heisenberg_opsum(og), og, ops)[0])[0])
    assert abs(E - E0) < 1e-7, (rec.energies, E0)
    assert rec.maxlinkdims[0] > 1 and max(rec.maxlinkdims) <= 32
    host = psi.to_host()
    v = host.to_dense()
    assert abs(np.vdot(v, v) - 1.0) < 1e-10
    # one expansion step leaves the state invariant and the previous vertex orthonormal
    prob = ns.EigsolveProblem(host, H)
    net = prob.net
    net.extract([5, 6])
    net.update_eigsolve()
    net.insert((1e-12, 1, 32))
    before = net.to_host().to_dense()
    info = net.extract([6], (1e-12, 1, 64), dict(algorithm=2, north_pass=1, expansion_factor=1.5, max_expand=2**62))
    after_host = net.to_host()
    assert info.expanded == 1
    A = np.asarray(after_host.tensors[5])
    legs = after_host.legs[5]
    bond = legs.index(("link", 5, 6))
    Am = np.moveaxis(A, bond, -1).reshape(-1, A.shape[bond])
    assert np.abs(Am.conj().T @ Am - np.eye(Am.shape[1])).max() < 1e-10
    after = after_host.to_dense()
    assert abs(abs(np.vdot(before, after)) - 1.0) < 1e-10 and abs(np.vdot(after, after) - 1.0) < 1e-10
