"""Experimental device code that the hooks do not use yet (round-2 groundwork); kept in a file that sorts after the parity
tests so that `pytest -x` reaches it last."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def ctx():
    import networksolvers_b200 as ns
    return ns.default_context()


@pytest.mark.parametrize("n,b", [(40, 4), (130, 32), (700, 64)])
def test_experimental_bulge_chasing_kernel(ctx, n, b):
    """csrc/sbr.cu (round-2 groundwork, not used by the eigensolver yet): the persistent chasing kernel turns a symmetric band
    matrix into a tridiagonal one with the same eigenvalues; same check as tests/test_cpu_dc.py runs on the host code path."""
    rng = np.random.default_rng(n + b)
    M = rng.standard_normal((n, n))
    A = np.tril(np.triu(M + M.T, -b), b)
    ab = np.zeros((2 * b + 1, n), order="F")
    for j in range(n):
        m = min(b, n - 1 - j)
        ab[:m + 1, j] = A[j:j + m + 1, j]
    out, V2, tau2 = ctx.sbr_chase(ab, b)
    nrm = np.linalg.norm(A, 2)
    assert np.abs(out[2:, :]).max() < 50 * EPS * nrm
    d, e = out[0, :], out[1, :n - 1]
    w = np.linalg.eigvalsh(np.diag(d) + np.diag(e, 1) + np.diag(e, -1))
    assert np.abs(w - np.linalg.eigvalsh(A)).max() < 100 * EPS * nrm * max(1.0, n / 1000)
