"""Oracle parity at the sizes where the persistent TMA + DMMA GEMM, grouped tile rasterisation, batched tensor maps and
the identity-channel splice are really in play (chi = 512 / 1024: every GEMM of the matvec spans many 128 x 128 tiles and
several waves of the 148-CTA grid), through the C ABI.  The oracle (NumPy/BLAS) needs a few seconds per case here.

Checked: the H_eff application (src/operator_map.jl:3-10), one full region step -- Ritz value of the 3-matvec Lanczos
(src/eigsolve.jl:14-28), kept dimension and truncation error of the truncating factorisation (src/inserter.jl:20-24) -- with
the identity channels skipped and not skipped, real and complex; a graded spectrum (12 decades) with cutoff 1e-14 through
nsb_insert at n = 1024 against the oracle's SVD route; the `eager` Lanczos rule against the oracle.

Tolerances: matvec 1e-13 sqrt(K) max-rel, Ritz value 1e-10 relative, truncation error 1e-8 absolute (BASELINE north_star)."""
import numpy as np
import pytest

from helpers import _olabel, to_oracle_ttn

pytestmark = pytest.mark.gpu


def _problem(chi, nsites, cplx, seed=7):
    import networksolvers_b200 as ns
    g = ns.path_graph(nsites)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi = ns.random_state(sites, chi, seed=seed, dtype=complex if cplx else float)
    return ns, H, ns.EigsolveProblem(state=psi, operator=H).net


@pytest.mark.parametrize("chi,nsites,cplx", [(512, 20, False), (1024, 22, False), (512, 20, True)])
@pytest.mark.parametrize("skip", [1, 0])
def test_matvec_and_region_step_match_oracle_at_scale(chi, nsites, cplx, skip):
    from oracle.local_solvers import lanczos_eigsolve
    from oracle.operator_map import optimal_map
    from oracle.projttn import ProjTTN, position
    from oracle.tensor import Tensor, factorize, link, site
    ns, H, net = _problem(chi, nsites, cplx)
    ctx = net.ctx
    ctx.set_option("skip_identity", skip)
    try:
        mid = nsites // 2
        region = [mid, mid + 1]
        net.extract(region)                      # gauge walk from both ends (blocked device QR) + environments
        legs, dims = net.local_info()
        assert dims == [chi, 2, 2, chi]
        ctx.reset_counters()
        theta, _ = net.local_download()
        out = net.matvec_host(theta)
        executed, dense = net.matvec_flops_executed(), net.matvec_flops()
        if skip:
            assert executed < 0.85 * dense, "canonical state: both environments carry an identity channel"
        else:
            assert executed == dense
        assert ctx.counters()["permute_bytes"] == 0
        # oracle on the tensors that are on the device now
        psio = to_oracle_ttn(net.to_host())
        P = position(ProjTTN(to_oracle_ttn(H, operator=True)), psio, region)
        th = Tensor(np.array(theta), [_olabel(l) for l in legs])
        ref = optimal_map(P, th).array(th.labels)
        tol = 1e-13 * np.sqrt(5 * chi) * np.abs(ref).max()
        assert np.abs(out - ref).max() <= tol, (np.abs(out - ref).max(), tol)
        # full region step
        val, info = net.update_eigsolve()
        oval, ovec, _ = lanczos_eigsolve(lambda x: optimal_map(P, x), th)
        assert info.nmatvec == 3
        assert abs(val - oval) <= 1e-10 * abs(oval), (val, oval)
        th2, _ = net.local_download()
        ov = np.vdot(ovec.array(th.labels), th2)
        assert abs(abs(ov) - 1.0) <= 1e-10                                   # same Ritz vector (up to a phase)
        ins = net.insert((0.0, 1, chi))
        L_, R_, finfo = factorize(ovec, [link(mid - 1, mid), site(mid)], link(mid, mid + 1), cutoff=0.0, maxdim=chi)
        assert ins.newdim == L_.dim(link(mid, mid + 1)) == chi
        assert abs(ins.truncerr - finfo["truncerr"]) <= 1e-8, (ins.truncerr, finfo["truncerr"])
    finally:
        ctx.set_option("skip_identity", 1)


@pytest.mark.parametrize("cutoff", [1e-14, 1e-12])
def test_graded_spectrum_small_cutoff_through_insert(cutoff):
    """TDVP quench shape (examples/quench_evolution.jl:21: cutoff 1e-14): a two-site tensor with singular values graded
    over 12 decades, 1024 x 1024, through nsb_insert.  The reference takes LAPACK SVD here (cutoff <= 1e-12); the device
    takes Gram + eigh with Rayleigh-quotient refinement of the spectrum (decomp = 3).  Kept dimension within the states whose
    weight is at the cutoff (a 1 % band), truncation error to 1e-8 absolute (in fact ~1e-16)."""
    from oracle.tensor import Tensor, factorize, link, site
    chi, nsites = 512, 20
    ns, H, net = _problem(chi, nsites, False)
    mid = nsites // 2
    net.extract([mid, mid + 1])
    legs, dims = net.local_info()
    n = 2 * chi
    rng = np.random.default_rng(11)
    Uo, _ = np.linalg.qr(rng.standard_normal((n, n)))
    Vo, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sig = 10.0 ** (-6.0 * np.arange(n) / n)           # sigma over 6 decades = sigma^2 (the truncation weights) over 12
    M = (Uo * sig) @ Vo.T
    theta = np.asfortranarray(M.reshape(dims, order="F"))
    net.local_upload(theta)
    ins = net.insert((cutoff, 1, n))
    assert ins.decomp == 3, "n >= 1024 with cutoff <= 1e-12: Gram + eigh route with refined spectrum"
    th = Tensor(theta, [_olabel(l) for l in legs])
    L_, R_, finfo = factorize(th, [link(mid - 1, mid), site(mid)], link(mid, mid + 1), cutoff=cutoff, maxdim=n)
    assert finfo["decomp"] == "svd"
    kref = L_.dim(link(mid, mid + 1))
    P = sig**2 / (sig**2).sum()
    # states whose inclusion changes the discarded weight by less than 2 % of the cutoff cannot be told apart
    tail = np.cumsum(P[::-1])[::-1]
    lo = int(np.searchsorted(-tail, -1.02 * cutoff))
    hi = int(np.searchsorted(-tail, -0.98 * cutoff))
    assert lo - 1 <= ins.newdim <= hi + 1, (ins.newdim, kref, lo, hi)
    assert abs(ins.newdim - kref) <= max(2, hi - lo), (ins.newdim, kref)
    assert abs(ins.truncerr - finfo["truncerr"]) <= 1e-8
    assert abs(ins.truncerr - finfo["truncerr"]) <= 0.05 * cutoff + 1e-16, (ins.truncerr, finfo["truncerr"])
    # the factors reproduce theta to the discarded weight, U is orthonormal
    host = net.to_host()
    A, la = host.tensors[mid], host.legs[mid]
    k = ins.newdim
    Am = np.transpose(A, [la.index(("link", mid, mid - 1)), la.index(("site", mid)), la.index(("link", mid, mid + 1))]).reshape(n, k, order="F")
    assert np.abs(Am.T @ Am - np.eye(k)).max() < 1e-11


@pytest.mark.parametrize("krylovdim", [3, 8])
def test_eager_lanczos_matches_oracle(krylovdim):
    """KrylovKit's `eager`: an early convergence test after every expansion, not an unconditional stop.  With
    tol = 1e-14 on a non-converged state it must still do `krylovdim` matvecs; with a loose tol it leaves early -- both as
    the oracle does."""
    from oracle.local_solvers import lanczos_eigsolve
    from oracle.operator_map import optimal_map
    from oracle.projttn import ProjTTN, position
    from oracle.tensor import Tensor
    ns, H, net = _problem(16, 10, False, seed=3)
    region = [5, 6]
    for tol in (1e-14, 5e-2):
        net.extract(region)
        legs, dims = net.local_info()
        theta, _ = net.local_download()
        psio = to_oracle_ttn(net.to_host())
        P = position(ProjTTN(to_oracle_ttn(H, operator=True)), psio, region)
        th = Tensor(np.array(theta), [_olabel(l) for l in legs])
        val, info = net.update_eigsolve(krylovdim=krylovdim, tol=tol, eager=True)
        oval, _, oinfo = lanczos_eigsolve(lambda x: optimal_map(P, x), th, krylovdim=krylovdim, tol=tol, eager=True)
        assert info.nmatvec == oinfo["numops"], (info.nmatvec, oinfo)
        assert abs(val - oval) <= 1e-11 * max(1.0, abs(oval))
        if tol == 1e-14:
            assert info.nmatvec == krylovdim          # made progress: not the one-matvec exit of round 1
        net.insert((0.0, 1, 16))
