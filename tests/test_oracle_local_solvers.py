"""Known-answer tests of the oracle's local solvers and small rules against closed forms computed independently here (dense
linear algebra on a small Hermitian matrix), so that the restated UPSTREAM rules (KrylovKit Lanczos / exponentiate, App. A.6 / A.7)
and the reference's own Runge-Kutta arithmetic (src/local_solvers/runge_kutta.jl:2-25) are pinned on mathematics rather than on
the restatement itself."""
import math

import numpy as np
import pytest
import scipy.linalg

from oracle.local_solvers import (lanczos_eigsolve, eigsolve_solver, runge_kutta_2, runge_kutta_4, runge_kutta_solver,
                                  exponentiate_solver)
from oracle.tensor import Tensor, link, site
from oracle.truncation_parameters import get_or_last, truncation_parameters
from oracle.subspace import compute_expansion
from oracle.region_plans import tdvp_sub_time_steps

LABELS = [link(1, 2), site(2), site(3), link(3, 4)]
SHAPE = (3, 2, 2, 3)
N = int(np.prod(SHAPE))


def _problem(cplx, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N, N)) + (1j * rng.standard_normal((N, N)) if cplx else 0.0)
    Hm = (A + A.conj().T) / 2
    v = rng.standard_normal(N) + (1j * rng.standard_normal(N) if cplx else 0.0)

    def op(x: Tensor) -> Tensor:
        return Tensor((Hm @ x.array(LABELS).reshape(N)).reshape(SHAPE), LABELS)

    return Hm, v, op, Tensor(v.reshape(SHAPE), LABELS)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("K", [1, 2, 3, 5])
def test_lanczos_ritz_value_is_the_lowest_eigenvalue_of_the_projected_operator(cplx, K):
    Hm, v, op, t0 = _problem(cplx)
    val, vec, info = lanczos_eigsolve(op, t0, krylovdim=K, maxiter=1, tol=1e-14)
    # independent: orthonormal basis of span{v, Hv, ..., H^(K-1) v} by QR, Rayleigh-Ritz on it
    Kry = np.stack([np.linalg.matrix_power(Hm, j) @ v for j in range(K)], axis=1)
    Q, _ = np.linalg.qr(Kry)
    w, y = np.linalg.eigh(Q.conj().T @ Hm @ Q)
    assert abs(val - w[0]) < 1e-11 * max(1.0, abs(w[0]))
    ritz = Q @ y[:, 0]
    got = vec.array(LABELS).reshape(N)
    assert abs(abs(np.vdot(ritz, got)) - 1.0) < 1e-10            # same vector up to a phase, unit norm
    assert info["numops"] == K and info["krylov"] == K


def test_lanczos_stops_on_an_invariant_subspace():
    Hm, v, op, _ = _problem(False)
    w, U = np.linalg.eigh(Hm)
    t0 = Tensor((U[:, 0] + U[:, 3]).reshape(SHAPE), LABELS)       # two-dimensional invariant subspace
    val, vec, info = lanczos_eigsolve(op, t0, krylovdim=6, maxiter=1, tol=1e-12)
    assert abs(val - w[0]) < 1e-11
    assert info["numops"] == 2


def test_eigsolve_solver_defaults_are_three_matvecs():
    Hm, v, op, t0 = _problem(False)
    count = [0]

    def counted(x):
        count[0] += 1
        return op(x)

    out = eigsolve_solver(counted, t0)
    assert count[0] == 3                                           # krylovdim = 3, maxiter = 1 (src/local_solvers/eigsolve.jl:8-9)
    val = out[0]
    assert val <= np.vdot(v, Hm @ v).real / np.vdot(v, v).real + 1e-12     # variational: not above the start's Rayleigh quotient


@pytest.mark.parametrize("cplx", [False, True])
def test_runge_kutta_steps_are_the_taylor_polynomials(cplx):
    Hm, v, op, t0 = _problem(cplx)
    t = -0.07j if cplx else 0.05
    taylor = lambda order: sum((t ** k / math.factorial(k)) * (np.linalg.matrix_power(Hm, k) @ v) for k in range(order + 1))
    r2 = runge_kutta_2(op, t, t0).array(LABELS).reshape(N)
    r4 = runge_kutta_4(op, t, t0).array(LABELS).reshape(N)
    assert np.abs(r2 - taylor(2)).max() < 1e-13 * np.abs(v).max()
    assert np.abs(r4 - taylor(4)).max() < 1e-13 * np.abs(v).max()
    for order, ref in ((2, r2), (4, r4)):
        out = runge_kutta_solver(op, t, t0, order=order)
        assert np.abs(out[0].array(LABELS).reshape(N) - ref).max() == 0.0
    with pytest.raises(Exception, match="must specify `order`"):
        runge_kutta_solver(op, t, t0, order=3)


@pytest.mark.parametrize("cplx", [False, True])
def test_exponentiate_matches_dense_expm(cplx):
    Hm, v, op, t0 = _problem(cplx)
    t = -0.3j
    th = Tensor(v.astype(complex).reshape(SHAPE), LABELS)
    out = exponentiate_solver(op, t, th)
    got = out[0].array(LABELS).reshape(N)
    ref = scipy.linalg.expm(t * Hm) @ v
    assert np.abs(got - ref).max() < 1e-10 * np.abs(ref).max()


def test_small_rules_known_answers():
    # src/truncation_parameters.jl:5: get_or_last(x, i) = (i >= length(x)) ? last(x) : x[i]   (1-based sweep index)
    assert [get_or_last([10, 20, 30], s) for s in (1, 2, 3, 4, 9)] == [10, 20, 30, 30, 30]
    assert get_or_last(7, 5) == 7
    tp = truncation_parameters(2, cutoff=[1e-6, 1e-9], maxdim=[10, 20, 40])
    assert (tp["cutoff"], tp["maxdim"], tp["mindim"]) == (1e-9, 20, 1)
    # src/subspace/subspace.jl:31-48
    assert compute_expansion(10, 40, expansion_factor=1.5) == 15
    assert compute_expansion(10, 40, expansion_factor=1.5, max_expand=4) == 4
    assert compute_expansion(10, 18, expansion_factor=1.5) == 8                  # basis_size - current_dim
    assert compute_expansion(10, 40, expansion_factor=1.5, maxdim=12) == 2       # maxdim - current_dim
    assert compute_expansion(10, 40, expansion_factor=1.5, maxdim=8) == 0        # never negative
    assert compute_expansion(3, 40, expansion_factor=1.1) == 4                   # ceil
    # src/region_plans/tdvp_region_plans.jl:1-13
    assert tdvp_sub_time_steps(1) == [1.0] and tdvp_sub_time_steps(2) == [0.5, 0.5]
    s = 1.0 / (2.0 - 2.0 ** (1.0 / 3.0))
    w4 = tdvp_sub_time_steps(4)
    assert np.allclose(w4, [s / 2, s / 2, 0.5 - s, 0.5 - s, s / 2, s / 2], rtol=0, atol=1e-16) and abs(sum(w4) - 1.0) < 1e-15
    with pytest.raises(Exception, match="not supported"):
        tdvp_sub_time_steps(3)
