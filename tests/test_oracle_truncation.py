"""Known-answer tests of the oracle's truncation rule (UPSTREAM NDTensors `truncate!!`, SURVEY.md App. A.5: relative cutoff on
the discarded weight, maxdim first, mindim last, negative tail zeroed, and for block-sparse spectra the `docut` rule).  The
expected values are worked out by hand from the rule's statement; the device follows the same rule
(`truncate_spectrum` in csrc/linalg.cu, `truncate_blocks` in csrc/net_qn.cu) and is compared with this oracle in the
`-m gpu` tests."""
import numpy as np
import pytest

from oracle.tensor import truncate_spectrum
from oracle.qn import truncate_merged


@pytest.mark.parametrize("P,kw,keep,terr", [
    # relative cutoff: discard while the discarded weight / total stays <= cutoff
    ([0.5, 0.3, 0.15, 0.05], dict(cutoff=0.06), 3, 0.05),
    ([0.5, 0.3, 0.15, 0.05], dict(cutoff=0.2), 2, 0.2),            # boundary: "<=" keeps discarding at equality
    ([0.5, 0.3, 0.15, 0.05], dict(cutoff=0.19999), 3, 0.05),
    # cutoff 0 discards exact zeros only
    ([0.7, 0.3, 0.0, 0.0], dict(cutoff=0.0), 2, 0.0),
    # maxdim is enforced first and its discarded weight counts toward the error
    ([0.5, 0.3, 0.15, 0.05], dict(cutoff=0.0, maxdim=2), 2, 0.2),
    ([0.5, 0.3, 0.15, 0.05], dict(cutoff=0.0, maxdim=10), 4, 0.0),
    # mindim stops the cutoff loop (not the maxdim loop)
    ([0.5, 0.3, 0.15, 0.05], dict(cutoff=0.9, mindim=3), 3, 0.05),
    ([0.5, 0.3, 0.15, 0.05], dict(cutoff=0.9, mindim=3, maxdim=2), 2, 0.2),
    # the scale is the sum of the spectrum, whatever its normalisation
    ([5.0, 3.0, 1.5, 0.5], dict(cutoff=0.06), 3, 0.05),
    # a negative tail (round-off of an eigen route) is zeroed before anything else
    ([1.0, 1e-18, -1e-17], dict(cutoff=0.0), 2, 0.0),
    ([1.0, -1e-17, -2e-17], dict(cutoff=0.0), 1, 0.0),
    # all-zero spectrum: scale falls back to 1, everything but mindim goes
    ([0.0, 0.0, 0.0], dict(cutoff=0.0), 1, 0.0),
    # a single value is always kept
    ([0.3], dict(cutoff=1.0), 1, 0.0),
    # never fewer than one -- the rule restores n = 1 after the loop and leaves the accumulated error as it is
    ([0.6, 0.4], dict(cutoff=1.0, mindim=0), 1, 1.0),
])
def test_truncation_rule_known_answers(P, kw, keep, terr):
    n, t = truncate_spectrum(np.array(P), **kw)
    assert n == keep
    assert t == pytest.approx(terr, abs=1e-15)


def test_block_sparse_spectrum_uses_docut():
    # merged spectrum [0.5, 0.3, 0.15, 0.05]: keep 3 -> docut = (0.15 + 0.05) / 2 = 0.1: every block keeps its values above it
    keep, terr = truncate_merged([np.array([0.5, 0.05]), np.array([0.3, 0.15])], cutoff=0.06, mindim=1, maxdim=10)
    assert keep == [1, 2] and terr == pytest.approx(0.05)
    # a degenerate pair straddling the cut is dropped as a whole (docut is raised by 1e-3 of the kept value): 2 kept, not 3
    keep, terr = truncate_merged([np.array([0.5, 0.1]), np.array([0.3, 0.1])], cutoff=0.0, mindim=1, maxdim=3)
    assert keep == [1, 1] and terr == pytest.approx(0.1)
    # nothing truncated: every non-negative value stays, zeros included
    keep, terr = truncate_merged([np.array([0.5, 0.0]), np.array([0.5])], cutoff=0.0, mindim=3, maxdim=10)
    assert keep == [2, 1] and terr == 0.0
    # empty block list / empty blocks
    assert truncate_merged([], cutoff=0.0, mindim=1, maxdim=4) == ([], 0.0)
    keep, _ = truncate_merged([np.zeros(0), np.array([1.0])], cutoff=0.0, mindim=1, maxdim=4)
    assert keep == [0, 1]


def test_factorize_route_follows_the_cutoff_label():
    """UPSTREAM ITensors `factorize` (App. A.4): cutoff <= 1e-12 -> SVD, larger -> eigen of the density matrix; both give the same
    truncation on a well-conditioned spectrum."""
    from oracle.tensor import Tensor, factorize, link, site
    rng = np.random.default_rng(0)
    th = Tensor(rng.standard_normal((4, 2, 2, 4)), [link(1, 2), site(2), site(3), link(3, 4)])
    outs = []
    for cutoff in (1e-13, 1e-8):
        U, C, spec = factorize(th, [link(1, 2), site(2)], link(2, 3), cutoff=cutoff, maxdim=5)
        outs.append((U.dim(link(2, 3)), spec))
    assert outs[0][0] == outs[1][0] == 5
