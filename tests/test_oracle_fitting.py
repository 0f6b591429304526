"""Oracle anchor for the fitting path (src/fitting.jl): the reference's own assertions of test/fitting/test_fitting.jl:
`truncate` to the original link dimension reproduces the state (fidelity 1) and `apply` with enough link dimension
reproduces <a|H|a> -- on the comb tree (3, 2), S=1/2, real and complex, 1-site and 2-site."""
import numpy as np
import pytest

from oracle import fitting as F
from oracle.graph import named_comb_tree
from oracle.models import heisenberg_opsum, spin_ops, ttno

EPS = np.finfo(float).eps


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_reference_fitting_assertions(dtype):
    g = named_comb_tree((3, 2))
    d, ops, _ = spin_ops("S=1/2")
    rng = np.random.default_rng(1234)
    a = F.random_tensornetwork(g, d, 3, rng, dtype)
    b = F.truncate(a, maxdim=3)                                  # test_fitting.jl:20-26
    f = F.inner(a, b) / np.sqrt(F.inner(a, a) * F.inner(b, b))
    assert abs(abs(f) - 1.0) <= 50 * EPS
    a = F.random_tensornetwork(g, d, 3, rng, dtype)
    b = F.truncate(a, maxdim=3, cutoff=1e-16, nsites=2)          # :28-35
    f = F.inner(a, b) / np.sqrt(F.inner(a, a) * F.inner(b, b))
    assert abs(abs(f) - 1.0) <= 50 * EPS and b.maxlinkdim() <= 3
    H = ttno(heisenberg_opsum(g), g, ops, dtype=dtype)
    a = F.random_tensornetwork(g, d, 2, rng, dtype)
    Ha = F.apply(H, a, maxdim=4, nsites=1, normalize=False)      # :37-43
    assert abs(F.inner(Ha, a) / F.inner(a, a, H) - 1.0) <= 100 * EPS
    a = F.random_tensornetwork(g, d, 2, rng, dtype)
    Ha = F.apply(H, a, maxdim=4, cutoff=1e-16, nsites=2, normalize=False)   # :45-51
    assert abs(F.inner(Ha, a) / F.inner(a, a, H) - 1.0) <= 100 * EPS


def test_truncation_below_exact_rank_is_variationally_optimal_on_a_chain():
    """On a chain, fitting with maxdim below the exact rank cannot beat (and should reach within 1e-10) the fidelity of
    the canonical SVD truncation sweep from the same state."""
    from oracle.graph import path_graph
    from oracle.models import random_ttn
    from oracle.ed import state_vector
    g = path_graph(6)
    a = random_ttn(g, 2, 8, seed=7)
    b = F.truncate(a, maxdim=3, nsweeps=12)
    va, vb = state_vector(a), state_vector(b)
    fid = abs(np.vdot(va, vb)) ** 2 / (np.vdot(va, va).real * np.vdot(vb, vb).real)
    # best rank-3 approximation at the middle cut bounds every MPS of bond dimension 3
    s = np.linalg.svd(va.reshape(8, 8), compute_uv=False)
    assert fid <= np.sum(s[:3] ** 2) / np.sum(s ** 2) + 1e-12
    assert fid > 0.5
