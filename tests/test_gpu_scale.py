"""Size-independent properties of the hot path at the BASELINE sizes (config 2, chi up to 4096), where the
oracle is too slow to be the checker: Hermiticity and linearity of H_eff, position independence of the energy
expectation (environment updates + gauge consistency), and the factorisation round trip."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _net(chi, nsites, dtype=np.float64, seed=1234):
    import networksolvers_b200 as ns
    g = ns.path_graph(nsites)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    mid = nsites // 2
    net = ns.DeviceNetwork.synthetic(H, sites, chi, seed=seed, dtype=dtype, ortho_region=[mid, mid + 1])
    return ns, net, [mid, mid + 1]


@pytest.mark.parametrize("chi,nsites,cplx", [(1024, 22, False), (512, 20, True), (4096, 26, False)])
def test_heff_hermitian_and_linear(chi, nsites, cplx):
    dt = np.complex128 if cplx else np.float64
    ns, net, region = _net(chi, nsites, dt)
    net.extract(region)
    net.ctx.reset_counters()             # the default context is shared by all tests of the session
    legs, dims = net.local_info()
    assert dims == [chi, 2, 2, chi]
    rng = np.random.default_rng(1)

    def rnd():
        a = rng.standard_normal(dims)
        if cplx:
            a = a + 1j * rng.standard_normal(dims)
        return np.asfortranarray(a.astype(dt))

    x, y = rnd(), rnd()
    Hx, Hy = net.matvec_host(x), net.matvec_host(y)
    lhs, rhs = np.vdot(x, Hy), np.vdot(Hx, y)
    scale = np.linalg.norm(x) * np.linalg.norm(Hy)
    assert abs(lhs - rhs) <= 1e-11 * scale, (lhs, rhs)              # <x|H y> = <H x|y>
    a, b = 0.7, -1.3
    Hz = net.matvec_host(np.asfortranarray(a * x + b * y))
    assert np.abs(Hz - (a * Hx + b * Hy)).max() <= 1e-11 * np.abs(Hz).max()
    assert net.ctx.counters()["permute_bytes"] == 0
    # analytic flop count of the fixed-order matvec
    w, d = 5, 2
    assert abs(net.matvec_flops() - (4 * w * d * d * chi**3 + 4 * w * w * d**3 * chi**2) * (4 if cplx else 1)) < 1.0


def test_energy_expectation_is_position_independent():
    """<theta|H_eff|theta> / <theta|theta> must not depend on the bond at which it is evaluated: exercises theta
    build, incremental environment updates in both directions, QR gauge moves and a non-truncating insert."""
    ns, net, region = _net(96, 14)
    net.set_ortho_region(list(net.graph.vertices))     # synthetic tensors are not orthonormal: force the gauge walk
    vals = []
    for reg in ([7, 8], [8, 9], [9, 10], [10, 9], [9, 8], [8, 7], [7, 6]):
        net.extract(reg)
        th, _ = net.local_download()
        Hth = net.matvec_host(th)
        vals.append(np.vdot(th, Hth).real / np.vdot(th, th).real)
        net.insert((0.0, 1, 10**9))                    # no truncation
    assert np.abs(np.array(vals) - vals[0]).max() <= 1e-10 * abs(vals[0]), vals


@pytest.mark.parametrize("n,cplx", [(1024, False), (512, True)])
def test_factorize_round_trip_at_scale(n, cplx):
    import networksolvers_b200 as ns
    ctx = ns.default_context()
    rng = np.random.default_rng(2)
    Uo, _ = np.linalg.qr(rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0))
    Vo, _ = np.linalg.qr(rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0))
    sig = np.exp(-10.0 * np.arange(n) / n)
    M = (Uo * sig) @ Vo.conj().T
    k = n // 2
    U, C, spec, info = ctx.factorize(M, cutoff=0.0, maxdim=k)
    assert info["newdim"] == k
    assert np.abs(U.conj().T @ U - np.eye(k)).max() < 5e-12
    terr = (sig[k:] ** 2).sum() / (sig**2).sum()
    assert abs(info["truncerr"] - terr) <= 1e-8 * terr + 1e-16
    resid = np.linalg.norm(U @ C - M) ** 2 / np.linalg.norm(M) ** 2
    assert abs(resid - terr) <= 1e-6 * terr + 1e-14
    assert np.abs(spec[:k] - sig[:k] ** 2).max() <= 5e-12          # LAPACK-level absolute accuracy on sigma^2


def test_synthetic_setup_is_bitwise_reproducible():
    """Two builds of the same synthetic state (Philox fill, gauge walk = one blocked compact-WY QR per edge, environments) hold
    bit-identical tensors and give a bit-identical H_eff application: what the multi-GPU path relies on when every rank builds
    its replica independently (guards the write-after-read hazard the QR panel kernel once had on row j of the updated column)."""
    import networksolvers_b200 as ns
    g = ns.path_graph(24)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    ctx = ns.default_context()
    outs = []
    for _ in range(3):
        net = ns.DeviceNetwork.synthetic(H, sites, 1024, seed=1234, ctx=ctx, ortho_region=[12, 13], canonical=True)
        net.extract([12, 13])
        theta, _ = net.local_download()
        y = net.matvec_device(1, download=True)
        outs.append((np.array(theta), np.array(y), np.array(net.site(3)[0]), np.array(net.site(20)[0])))
        net.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
