"""GPU parity of the dense kernels under the hooks (GEMM variants, QR, truncating factorisation, range
finder) against NumPy/LAPACK, called through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import networksolvers_b200 as ns
    return ns.default_context()


def _rand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return a


def _op(a, op):
    return {"N": a, "T": a.T, "C": a.conj().T, "J": a.conj()}[op]


SHAPES = [(1, 1, 1), (5, 3, 2), (37, 53, 29), (128, 128, 16), (130, 70, 33), (200, 300, 100), (257, 129, 65),
          (512, 384, 260)]


@pytest.mark.parametrize("impl", [1, 2, 3])
@pytest.mark.parametrize("cplx", [False, True])
def test_gemm_all_ops(ctx, impl, cplx):
    rng = np.random.default_rng(7)
    ops = ["N", "T", "C", "J"] if cplx else ["N", "T"]
    worst = 0.0
    for (m, n, k) in SHAPES:
        for opa in ops:
            for opb in ops:
                A = _rand(rng, (m, k) if opa in "NJ" else (k, m), cplx)
                B = _rand(rng, (k, n) if opb in "NJ" else (n, k), cplx)
                ref = _op(A, opa) @ _op(B, opb)
                got = ctx.gemm(A, B, opa, opb, impl=impl)
                err = np.abs(got - ref).max() / (np.abs(ref).max() + 1e-300)
                tol = 1e-13 * np.sqrt(k) + 1e-15      # max rel err <= 1e-13 sqrt(K)
                assert np.isfinite(got).all(), (impl, cplx, m, n, k, opa, opb)
                assert err <= tol, (impl, cplx, m, n, k, opa, opb, err)
                worst = max(worst, err)
    print("worst rel err", worst)


@pytest.mark.parametrize("impl", [2, 3])
def test_gemm_unaligned_leading_dimension(ctx, impl):
    """Odd leading dimensions are not TMA-addressable: impl 3 must fall back to the cp.async path."""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((77, 45))
    B = rng.standard_normal((45, 31))
    assert np.allclose(ctx.gemm(A, B, impl=impl), A @ B, rtol=0, atol=1e-12)
    assert np.allclose(ctx.gemm(A.T.copy(), B, "T", "N", impl=impl), A @ B, rtol=0, atol=1e-12)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("smem", [1, 0])
@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (3, 7), (64, 64), (150, 40), (40, 150), (128, 100), (33, 1)])
def test_qr(ctx, cplx, shape, smem):
    """Thin QR: the single-launch shared-memory kernel (smem = 1: everything here fits) and the column-at-a-time / blocked
    kernels (smem = 0) it replaces at these sizes."""
    rng = np.random.default_rng(11)
    M = _rand(rng, shape, cplx)
    if shape == (128, 100):
        M[:, 7] = 0.0                   # trivial reflector
        M[:, 50] = M[:, 2]              # dependent column
    ctx.set_option("qr_smem", smem)
    try:
        Q, R = ctx.qr(M)
    finally:
        ctx.set_option("qr_smem", 1)
    k = min(shape)
    assert Q.shape == (shape[0], k) and R.shape == (k, shape[1])
    assert np.abs(Q.conj().T @ Q - np.eye(k)).max() < 1e-13
    assert np.abs(Q @ R - M).max() < 1e-12
    assert np.abs(np.tril(R, -1)).max() == 0.0


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(2048, 1024), (300, 200), (200, 300), (1024, 512), (130, 64), (777, 333)])
def test_qr_blocked_compact_wy(ctx, cplx, shape):
    """Blocked Householder QR (one launch per panel column + compact-WY GEMM updates): same contract as LAPACK geqrf/orgqr
    up to the sign convention -- Q orthonormal, R upper triangular, Q R = M -- at 512-thread (rows >= 2048) and 256-thread
    panel kernels, tall, wide and ragged panel counts; rank-deficient input with exactly zero columns."""
    rng = np.random.default_rng(23)
    M = _rand(rng, shape, cplx)
    M[:, 5] = 0.0                       # zero column: trivial reflector inside a panel
    M[:, min(70, shape[1] - 1)] = M[:, 3]     # linearly dependent column (second panel when there is one)
    ctx.set_option("qr_smem", 0)        # (130 x 64 would fit the single-launch kernel)
    try:
        Q, R = ctx.qr(M)
    finally:
        ctx.set_option("qr_smem", 1)
    k = min(shape)
    assert Q.shape == (shape[0], k) and R.shape == (k, shape[1])
    assert np.abs(Q.conj().T @ Q - np.eye(k)).max() < 5e-13
    assert np.abs(Q @ R - M).max() < 1e-11
    assert np.abs(np.tril(R, -1)).max() == 0.0


def test_qr_rank_deficient(ctx):
    rng = np.random.default_rng(5)
    M = rng.standard_normal((30, 4)) @ rng.standard_normal((4, 12))
    M[:, 3] = 0.0
    Q, R = ctx.qr(M)
    assert np.abs(Q.T @ Q - np.eye(12)).max() < 1e-12
    assert np.abs(Q @ R - M).max() < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (8, 8), (33, 20), (20, 33), (96, 96), (64, 200)])
def test_factorize_full_spectrum(ctx, cplx, shape):
    """cutoff 0, no maxdim: SVD route; spectrum = sigma^2 of LAPACK, U orthonormal, U C = M."""
    rng = np.random.default_rng(13)
    M = _rand(rng, shape, cplx)
    U, Cm, spec, info = ctx.factorize(M, cutoff=0.0)
    s = np.linalg.svd(M, compute_uv=False)
    k = min(shape)
    assert info["newdim"] == k and info["decomp"] == 1
    assert np.abs(spec - s**2).max() <= 1e-12 * s[0] ** 2
    assert np.abs(U.conj().T @ U - np.eye(k)).max() < 1e-12
    assert np.abs(U @ Cm - M).max() < 1e-12 * max(1.0, s[0])


@pytest.mark.parametrize("cutoff,maxdim", [(1e-12, None), (1e-8, None), (1e-4, None), (0.0, 10), (1e-6, 7)])
def test_factorize_truncation_rule(ctx, cutoff, maxdim):
    """Truncation rule (NDTensors truncate!) vs the oracle on a decaying spectrum."""
    from oracle.tensor import truncate_spectrum
    rng = np.random.default_rng(17)
    n = 48
    Uo, _ = np.linalg.qr(rng.standard_normal((n, n)))
    Vo, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sig = np.exp(-0.6 * np.arange(n))
    M = (Uo * sig) @ Vo.T
    U, Cm, spec, info = ctx.factorize(M, cutoff=cutoff, maxdim=maxdim)
    nk, terr = truncate_spectrum(sig**2, cutoff=cutoff, mindim=1, maxdim=maxdim)
    assert info["newdim"] == nk
    assert abs(info["truncerr"] - terr) <= 1e-8 * max(terr, 1e-30) + 1e-18
    assert info["decomp"] == (1 if cutoff <= 1e-12 else 2)
    # kept subspace reproduces the best rank-nk approximation
    best = (Uo[:, :nk] * sig[:nk]) @ Vo[:, :nk].T
    assert np.abs(U @ Cm - best).max() < 1e-10


def test_factorize_rank_deficient_drops_exact_zeros(ctx):
    rng = np.random.default_rng(19)
    M = rng.standard_normal((24, 5)) @ rng.standard_normal((5, 30))
    U, Cm, spec, info = ctx.factorize(M, cutoff=1e-14)
    assert info["newdim"] == 5
    assert np.abs(U @ Cm - M).max() < 1e-11


@pytest.mark.parametrize("cplx", [False, True])
def test_range_finder(ctx, cplx):
    """src/sketched_linear_algebra/range_finder.jl: orthonormal basis of range(A), rank <= max_rank+oversample,
    stops early when the range is exhausted."""
    rng = np.random.default_rng(23)
    A = _rand(rng, (80, 6), cplx) @ _rand(rng, (6, 50), cplx)
    Q = ctx.range_finder(A, max_rank=20, oversample=2)
    assert Q.shape[1] == 6                      # range exhausted after 6 vectors (norm < 1e-12 stop)
    assert np.abs(Q.conj().T @ Q - np.eye(6)).max() < 1e-12
    assert np.abs(Q @ (Q.conj().T @ A) - A).max() < 1e-10
    Q = ctx.range_finder(A, max_rank=3, oversample=2)
    assert Q.shape[1] == 5
    A2 = _rand(rng, (40, 40), cplx)
    assert ctx.range_finder(A2, max_rank=0).shape[1] == 0


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(40, 40), (130, 100), (100, 130), (260, 200), (200, 333)])
def test_factorize_blocked_jacobi(ctx, cplx, shape):
    """The blocked (GEMM-rich) Jacobi path, forced at small sizes, against LAPACK."""
    rng = np.random.default_rng(29)
    M = _rand(rng, shape, cplx) * np.exp(-0.05 * np.arange(shape[1]))[None, :]
    ctx.set_option("jacobi_block_min_n", 0)
    ctx.set_option("jacobi_precondition_min_n", 0)
    ctx.set_option("jacobi_cluster_max_n", 0)
    ctx.set_option("jacobi_dsmem_max_n", 0)
    try:
        U, Cm, spec, info = ctx.factorize(M, cutoff=0.0)
        U2, C2, spec2, info2 = ctx.factorize(M, cutoff=1e-10, maxdim=shape[1] // 3)
    finally:
        ctx.set_option("jacobi_block_min_n", 48)
        ctx.set_option("jacobi_precondition_min_n", 1024)
        ctx.set_option("jacobi_cluster_max_n", 112)
        ctx.set_option("jacobi_dsmem_max_n", 256)
    s = np.linalg.svd(M, compute_uv=False)
    k = min(shape)
    assert info["newdim"] == k
    assert np.abs(spec - s**2).max() <= 1e-12 * s[0] ** 2
    assert np.abs(U.conj().T @ U - np.eye(k)).max() < 1e-12
    assert np.abs(U @ Cm - M).max() < 1e-12 * max(1.0, s[0])
    from oracle.tensor import truncate_spectrum
    nk, terr = truncate_spectrum(s**2, cutoff=1e-10, mindim=1, maxdim=shape[1] // 3)
    assert info2["newdim"] == nk and abs(info2["truncerr"] - terr) <= 1e-8 * max(terr, 1e-30) + 1e-18
    Ub, sb, Vb = np.linalg.svd(M, full_matrices=False)
    best = (Ub[:, :nk] * sb[:nk]) @ Vb[:nk]
    assert np.abs(U2 @ C2 - best).max() < 1e-10 * max(1.0, s[0])


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(2, 2), (3, 3), (31, 31), (32, 64), (100, 100), (130, 100), (128, 128), (166, 166), (255, 300),
                                   (256, 256), (4, 5000)])
def test_factorize_cluster_jacobi(ctx, cplx, shape):
    """The single-launch Jacobi kernels (one CTA with the matrix in shared memory; beyond 224 KB a thread-block cluster of up to
    8 CTAs on the L2-resident matrix; register-resident column pairs up to 256 / 128 rows, streamed beyond)
    against a known graded spectrum (12 decades of sigma^2): every sigma to ~1e-14 sigma_1, U orthonormal, U C = M, and the
    truncation rule of the oracle on the same spectrum."""
    from oracle.tensor import truncate_spectrum
    rng = np.random.default_rng(31)
    rows, cols = shape
    k = min(shape)
    Uo, _ = np.linalg.qr(_rand(rng, (rows, k), cplx))
    Vo, _ = np.linalg.qr(_rand(rng, (cols, k), cplx))
    sig = 10.0 ** (-6.0 * np.arange(k) / max(k - 1, 1))
    M = (Uo * sig) @ Vo.conj().T
    ctx.set_option("jacobi_cluster_max_n", 256)     # (defaults: 112, and the tournament kernel takes 41 .. 256 first)
    ctx.set_option("jacobi_dsmem_max_n", 0)
    try:
        U, Cm, spec, info = ctx.factorize(M, cutoff=0.0)
        U2, C2, spec2, info2 = ctx.factorize(M, cutoff=1e-10, maxdim=max(k // 2, 1))
    finally:
        ctx.set_option("jacobi_cluster_max_n", 112)
        ctx.set_option("jacobi_dsmem_max_n", 256)
    assert info["newdim"] == k and info["decomp"] == 1 and 1 <= info["sweeps"] < 60
    assert (np.abs(spec - sig**2) <= 4e-13 * sig * sig[0]).all()   # |d sigma| <~ 1e-13 sigma_1 over 12 decades of sigma^2
    assert np.abs(U.conj().T @ U - np.eye(k)).max() < 1e-12
    assert np.abs(U @ Cm - M).max() < 1e-13
    nk, terr = truncate_spectrum(sig**2, cutoff=1e-10, mindim=1, maxdim=max(k // 2, 1))
    assert info2["newdim"] == nk and abs(info2["truncerr"] - terr) <= 1e-8 * max(terr, 1e-30) + 1e-18
    best = (Uo[:, :nk] * sig[:nk]) @ Vo[:, :nk].conj().T
    assert np.abs(U2 @ C2 - best).max() < 1e-10


def test_factorize_cluster_matches_round_kernels(ctx):
    """Same matrix through the cluster launch and through the launch-per-round kernels: equal spectra and kept subspaces."""
    rng = np.random.default_rng(37)
    M = rng.standard_normal((40, 40)) * np.exp(-0.3 * np.arange(40))[None, :]
    U, Cm, spec, info = ctx.factorize(M, cutoff=1e-12, maxdim=20)
    ctx.set_option("jacobi_cluster_max_n", 0)
    try:
        U0, C0, spec0, info0 = ctx.factorize(M, cutoff=1e-12, maxdim=20)
    finally:
        ctx.set_option("jacobi_cluster_max_n", 112)
    assert info["newdim"] == info0["newdim"]
    assert np.abs(spec - spec0).max() <= 1e-13 * spec0[0]
    assert np.abs(U @ Cm - U0 @ C0).max() < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(2, 2), (5, 7), (40, 40), (120, 120), (128, 128), (166, 166), (130, 200), (255, 256), (256, 256)])
def test_factorize_dsmem_tournament_jacobi(ctx, cplx, shape):
    """The cluster / distributed-shared-memory tournament Jacobi (one warp per slot, columns handed between CTAs through
    st.shared::cluster), forced at every size it accepts (columns of <= 256 real / 128 complex rows), against a known graded
    spectrum and the oracle's truncation rule; 1, 2, 4 and 8 CTAs, padded slots (odd n, n not a multiple of the cluster)."""
    from oracle.tensor import truncate_spectrum
    if cplx and max(shape) > 128:
        pytest.skip("complex columns longer than 128 rows do not fit the register depth: routed to the other kernels")
    rng = np.random.default_rng(41)
    rows, cols = shape
    k = min(shape)
    Uo, _ = np.linalg.qr(_rand(rng, (rows, k), cplx))
    Vo, _ = np.linalg.qr(_rand(rng, (cols, k), cplx))
    sig = 10.0 ** (-6.0 * np.arange(k) / max(k - 1, 1))
    M = (Uo * sig) @ Vo.conj().T
    ctx.set_option("jacobi_cluster_max_n", 0)
    ctx.set_option("jacobi_dsmem_min_n", 0)
    ctx.reset_counters()
    try:
        U, Cm, spec, info = ctx.factorize(M, cutoff=0.0)
        U2, C2, spec2, info2 = ctx.factorize(M, cutoff=1e-10, maxdim=max(k // 2, 1))
    finally:
        ctx.set_option("jacobi_cluster_max_n", 112)
        ctx.set_option("jacobi_dsmem_min_n", 41)
    assert info["newdim"] == k and info["decomp"] == 1 and 1 <= info["sweeps"] < 60
    assert (np.abs(spec - sig**2) <= 4e-13 * sig * sig[0]).all()
    assert np.abs(U.conj().T @ U - np.eye(k)).max() < 1e-12
    assert np.abs(U @ Cm - M).max() < 1e-13
    nk, terr = truncate_spectrum(sig**2, cutoff=1e-10, mindim=1, maxdim=max(k // 2, 1))
    assert info2["newdim"] == nk and abs(info2["truncerr"] - terr) <= 1e-8 * max(terr, 1e-30) + 1e-18
    best = (Uo[:, :nk] * sig[:nk]) @ Vo[:, :nk].conj().T
    assert np.abs(U2 @ C2 - best).max() < 1e-10
