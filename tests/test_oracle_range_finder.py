"""Properties of the oracle's range finder (src/sketched_linear_algebra/range_finder.jl:6-64) on matrices whose range is known:
what it returns is an orthonormal basis, it stops at the numerical rank (the next sample has nothing left after two passes of
Gram-Schmidt), the sketch is capped at max_rank + oversample, and the argument checks of the linear-map form."""
import numpy as np
import pytest

from oracle.range_finder import range_finder, range_finder_map


def _low_rank(m, n, r, rng, cplx=False):
    g = lambda *s: rng.standard_normal(s) + (1j * rng.standard_normal(s) if cplx else 0.0)
    return g(m, r) @ g(r, n)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("m,n,r", [(40, 30, 5), (25, 60, 12), (16, 16, 16)])
def test_basis_is_orthonormal_and_spans_the_range(m, n, r, cplx):
    rng = np.random.default_rng(1)
    A = _low_rank(m, n, r, rng, cplx)
    rv = lambda: rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0.0)
    Q = np.stack(range_finder_map(lambda x: A @ x, rv), axis=1)
    assert Q.shape[1] == r                                        # stops at the rank: the (r + 1)-th sample is inside the span
    assert np.abs(Q.conj().T @ Q - np.eye(r)).max() < 1e-12
    assert np.linalg.norm(A - Q @ (Q.conj().T @ A)) < 1e-10 * np.linalg.norm(A)


def test_sketch_size_is_max_rank_plus_oversample():
    rng = np.random.default_rng(2)
    A = rng.standard_normal((50, 50))
    for max_rank, oversample, expect in [(4, 2, 6), (4, 0, 4), (48, 5, 50), (0, 2, 0), (-1, 2, 0)]:
        Q = range_finder_map(lambda x: A @ x, lambda: rng.standard_normal(50), max_rank=max_rank, oversample=oversample)
        assert len(Q) == expect


def test_zero_map_and_argument_checks():
    rng = np.random.default_rng(3)
    assert range_finder_map(lambda x: 0.0 * x, lambda: rng.standard_normal(8)) == []
    with pytest.raises(ValueError, match="should equal domain_size"):
        range_finder_map(lambda x: x, lambda: rng.standard_normal(8), domain_size=9)
    # sample form with a known range size: no first probe is spent on measuring it
    calls = [0]

    def sample():
        calls[0] += 1
        return rng.standard_normal(6)

    Q = range_finder(sample, range_size=6, max_rank=3, oversample=1)
    assert len(Q) == 4 and calls[0] == 4


def test_cutoff_stops_after_the_first_small_vector():
    """The experimental cutoff (range_finder.jl:41): the vector whose remainder falls below it is still kept, then the loop ends."""
    rng = np.random.default_rng(4)
    U, _ = np.linalg.qr(rng.standard_normal((30, 30)))
    A = (U * np.concatenate([np.ones(3), 1e-6 * np.ones(27)])) @ U.T
    Q = range_finder_map(lambda x: A @ x, lambda: rng.standard_normal(30), cutoff=1e-3)
    assert len(Q) == 4
