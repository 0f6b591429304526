"""GPU parity of the fitting path (src/fitting.jl: FittingProblem extracter / updater, fit_tensornetwork, itn.truncate,
itn.apply) through the C ABI (nsb_fit_target_upload / nsb_extract / nsb_update_fit / nsb_insert).  The assertions are
the reference's own (test/fitting/test_fitting.jl:20-51, test/fitting/fitting_regression_test.jl:44-60) plus equality
with the oracle restatement; the device library computes in FP64 / complex128 only (the reference also runs Float32)."""
import numpy as np
import pytest

from helpers import neel, to_oracle_ttn

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _ns():
    import networksolvers_b200 as ns
    return ns


def _dense_op(H):
    from oracle.ed import ttno_dense
    Ho = to_oracle_ttn(H, True)
    v = H.graph.vertices[0]
    d = H.tensors[v].shape[H.legs[v].index(("site", v))]
    return ttno_dense(Ho, Ho.graph, d)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_reference_fitting_assertions(dtype):
    ns = _ns()
    from oracle import fitting as OF
    g = ns.named_comb_tree((3, 2))
    s = ns.siteinds("S=1/2", g)
    rng = np.random.default_rng(1234)
    # one-site truncation (test_fitting.jl:20-26)
    a = ns.random_tensornetwork(s, 3, rng, dtype)
    b = ns.truncate(a, maxdim=3).to_host()
    va, vb = a.to_dense(), b.to_dense()
    f = np.vdot(va, vb) / np.sqrt(np.vdot(va, va) * np.vdot(vb, vb))
    assert abs(abs(f) - 1.0) <= 50 * EPS
    bo = OF.truncate(to_oracle_ttn(a), maxdim=3)
    from oracle.ed import state_vector
    vo = state_vector(bo)
    assert abs(abs(np.vdot(vo, vb)) / np.sqrt(np.vdot(vo, vo).real * np.vdot(vb, vb).real) - 1.0) <= 50 * EPS
    # two-site truncation (:28-35)
    a = ns.random_tensornetwork(s, 3, rng, dtype)
    b = ns.truncate(a, maxdim=3, cutoff=1e-16, nsites=2).to_host()
    va, vb = a.to_dense(), b.to_dense()
    f = np.vdot(va, vb) / np.sqrt(np.vdot(va, va) * np.vdot(vb, vb))
    assert abs(abs(f) - 1.0) <= 50 * EPS and b.maxlinkdim() <= 3
    # one-site apply, no normalisation (:37-43)
    H = ns.ttno(ns.heisenberg(g), s, dtype=dtype)
    Hd = _dense_op(H)
    a = ns.random_tensornetwork(s, 2, rng, dtype)
    Ha = ns.apply(H, a, maxdim=4, nsites=1, normalize=False).to_host()
    va = a.to_dense()
    f = np.vdot(Ha.to_dense(), va) / np.vdot(va, Hd @ va)
    assert abs(f - 1.0) <= 200 * EPS
    # two-site apply (:45-51)
    a = ns.random_tensornetwork(s, 2, rng, dtype)
    Ha = ns.apply(H, a, maxdim=4, cutoff=1e-16, nsites=2, normalize=False).to_host()
    va = a.to_dense()
    f = np.vdot(Ha.to_dense(), va) / np.vdot(va, Hd @ va)
    assert abs(f - 1.0) <= 200 * EPS
    # the fitted state itself equals A|a> (the tree admits every state at link dimension 4)
    w = Hd @ va
    assert np.abs(Ha.to_dense() - w).max() <= 1e-12 * np.abs(w).max()


def test_fitting_regression_apply_on_dmrg_state():
    """fitting_regression_test.jl:44-60: apply O = S+_3 S-_5 + S-_3 S+_5 to a DMRG state of the N = 8 chain."""
    ns = _ns()
    g = ns.path_graph(8)
    s = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), s)
    psi0 = ns.product_state(s, neel(g))
    trunc = dict(maxdim=50, cutoff=1e-5)
    E, psi = ns.dmrg(H, psi0, nsweeps=5, nsites=2, extracter_kwargs=dict(trunc=trunc), inserter_kwargs=dict(trunc=trunc))
    psi = psi.to_host()
    O = ns.product_operator_sum(s, [(1.0, {3: "S+", 5: "S-"}), (1.0, {3: "S-", 5: "S+"})])
    Opsi = ns.apply(O, psi, maxdim=60, nsites=2, normalize=False).to_host()
    v = psi.to_dense()
    f = np.vdot(Opsi.to_dense(), v) / np.vdot(v, _dense_op(O) @ v)
    assert abs(f - 1.0) <= 1e-10


def test_fitting_below_exact_rank_matches_oracle():
    """Truncating below the exact rank: same overlap history end point as the oracle restatement."""
    ns = _ns()
    from oracle import fitting as OF
    from oracle.ed import state_vector
    g = ns.path_graph(6)
    s = ns.siteinds("S=1/2", g)
    a = ns.random_state(s, 8, seed=7)
    b = ns.truncate(a, maxdim=3, nsweeps=12).to_host()
    bo = OF.truncate(to_oracle_ttn(a), maxdim=3, nsweeps=12)
    va, vb, vo = a.to_dense(), b.to_dense(), state_vector(bo)
    fid = abs(np.vdot(va, vb)) ** 2 / (np.vdot(va, va).real * np.vdot(vb, vb).real)
    fido = abs(np.vdot(va, vo)) ** 2 / (np.vdot(va, va).real * np.vdot(vo, vo).real)
    assert abs(fid - fido) <= 1e-10
    assert abs(abs(np.vdot(vo, vb)) - 1.0) <= 1e-9        # both normalised: the same state up to a sign
