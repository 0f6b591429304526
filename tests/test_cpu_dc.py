"""CPU checks of the divide & conquer logic the device eigensolver shares with the host (csrc/dc_secular.h):
deflation planning + secular root finder + Gu-Eisenstat vectors, driven through tests/dc_cpu_harness.cpp
(compiled here with g++).  The kernels / GEMMs of csrc/eigh.cu are covered by tests/test_gpu_eigh.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("dc") / "libdc_cpu.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(HERE, "dc_cpu_harness.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.dc_secular_root.restype = C.c_double
    lib.dc_leaf_bounds.restype = C.c_int64
    return lib


def dc_eigh_tridiag(lib, d, e, leaf):
    n = len(d)
    bounds = np.zeros(n + 2, dtype=np.int64)
    nb1 = lib.dc_leaf_bounds(C.c_int64(n), C.c_int64(leaf), bounds.ctypes.data_as(C.c_void_p), C.c_int64(len(bounds)))
    bounds = bounds[:nb1].copy()
    dm = d.copy()
    for x in bounds[1:-1]:
        dm[x - 1] -= abs(e[x - 1])
        dm[x] -= abs(e[x - 1])
    D = np.zeros(n)
    Z = np.zeros((n, n), order="F")
    for k in range(len(bounds) - 1):
        lo, hi = bounds[k], bounds[k + 1]
        T = np.diag(dm[lo:hi]) + np.diag(e[lo:hi - 1], 1) + np.diag(e[lo:hi - 1], -1)
        w, v = np.linalg.eigh(T)
        D[lo:hi] = w
        Z[lo:hi, lo:hi] = v
    stats = np.zeros(3, dtype=np.int64)
    e = np.ascontiguousarray(e)
    rc = lib.dc_merge_all(C.c_int64(n), e.ctypes.data_as(C.c_void_p), bounds.ctypes.data_as(C.c_void_p), C.c_int64(len(bounds) - 1),
                          D.ctypes.data_as(C.c_void_p), Z.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return D, Z, stats


def tridiag_cases(n, rng):
    from scipy.linalg import hessenberg
    def tri(A):
        H = hessenberg(A)
        return np.diag(H).copy(), np.diag(H, -1).copy()
    M = rng.standard_normal((n, n))
    yield "gauss", tri(M + M.T)
    G = rng.standard_normal((n, 2 * n))
    yield "wishart", tri(G @ G.T)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    yield "graded", tri((Q * np.exp(-40.0 * np.arange(n) / n)) @ Q.T)
    yield "lowrank", tri((Q[:, :3] * np.array([1.0, 0.5, 1e-3])) @ Q[:, :3].T)
    yield "identity+rank1", tri(np.eye(n) + 1e-3 * np.outer(Q[:, 0], Q[:, 0]))
    cl = np.repeat(np.arange(1, n // 8 + 2), 8)[:n].astype(float)
    yield "clustered", tri((Q * cl) @ Q.T)
    yield "wilkinson", (np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1))
    yield "hopping", (np.zeros(n), -np.ones(n - 1))
    yield "decoupled", (rng.standard_normal(n), np.where(np.arange(n - 1) % 7 == 3, 0.0, rng.standard_normal(n - 1)))


@pytest.mark.parametrize("n,leaf", [(37, 8), (128, 16), (301, 32)])
def test_dc_matches_eigh(lib, n, leaf):
    rng = np.random.default_rng(n)
    for name, (d, e) in tridiag_cases(n, rng):
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        D, Z, stats = dc_eigh_tridiag(lib, d, e, leaf)
        nrm = max(np.linalg.norm(T, 2), 1e-300)
        res = np.linalg.norm(T @ Z - Z * D[None, :]) / (nrm * n)
        orth = np.linalg.norm(Z.T @ Z - np.eye(n)) / n
        err = np.max(np.abs(np.sort(D) - np.linalg.eigvalsh(T))) / nrm
        assert res < 20 * EPS and orth < 20 * EPS and err < 50 * EPS, (name, n, res, orth, err, stats)
        assert stats[1] < 60, (name, stats)


def test_secular_root_relative_accuracy(lib):
    """d_j - lambda_i must be accurate to a few ulps even when lambda_i is within 1e-14 of a pole."""
    from fractions import Fraction
    rng = np.random.default_rng(5)
    K = 12
    d = np.sort(rng.standard_normal(K))
    z = rng.standard_normal(K)
    z[3] = 1e-7          # root hugging d[3]
    z /= np.linalg.norm(z)
    z2 = z * z
    rho = 0.7
    for i in range(K):
        delta = np.zeros(K)
        it = C.c_int(0)
        lam = lib.dc_secular_root(C.c_int(K), C.c_int(i), d.ctypes.data_as(C.c_void_p), z2.ctypes.data_as(C.c_void_p), C.c_double(rho),
                                  delta.ctypes.data_as(C.c_void_p), C.byref(it))
        # exact rational evaluation of the secular function at the implied lambda = d_org - delta_org
        org = int(np.argmin(np.abs(delta)))
        lam_q = Fraction(d[org]) - Fraction(delta[org])
        f = 1 + Fraction(rho) * sum(Fraction(z2[j]) / (Fraction(d[j]) - lam_q) for j in range(K))
        scale = 1 + float(Fraction(rho) * sum(abs(Fraction(z2[j]) / (Fraction(d[j]) - lam_q)) for j in range(K)))
        assert abs(float(f)) <= 64 * EPS * scale, (i, float(f), scale)
        for j in range(K):
            exact = float(Fraction(d[j]) - lam_q)
            assert abs(delta[j] - exact) <= 4 * EPS * abs(exact), (i, j)
        assert d[i] < lam and (i == K - 1 or lam < d[i + 1])


def test_symmetric_panel_unit_scheme():
    """Index scheme of trd_panel_sym_kernel (csrc/eigh.cu) restated in NumPy (tools/proto_trd_sym.py): lower-triangle
    work units, per-unit dot / row partials, fixed-order summation -- equals the dense product A_trail v for every unit
    width, odd and even column offsets, and never touches the (NaN-poisoned) upper triangle."""
    import importlib.util, os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "proto_trd_sym.py")
    spec = importlib.util.spec_from_file_location("proto_trd_sym", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = [(132, 0, 0, 7, 16), (132, 1, 1, 7, 32), (700, 64, 0, 444, 64), (700, 65, 1, 296, 0), (1100, 131, 3, 30, 32),
             (1100, 1090, 50, 444, 0), (1540, 7, 7, 296, 16)]
    assert mod.check(cases, verbose=False) < 1e-13


def test_ozaki_int8_gemm_matches_fp64():
    """tools/proto_ozaki.py (round-2 candidate for the H_eff GEMMs): an FP64 product assembled from exact 8-bit integer GEMMs
    (error-free 7-bit slicing, int32 accumulation, anti-diagonal sums) reaches FP64 accuracy with 9 slices = 45 integer GEMMs."""
    import importlib.util, os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "proto_ozaki.py")
    spec = importlib.util.spec_from_file_location("proto_ozaki", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = mod.check(m=48, k=1024, n=40, verbose=False)
    for name in ("gauss", "graded"):
        assert res[(name, 9)][0] < 2.0 * max(res[(name, "fp64")][0], 2.3e-16), (name, res[(name, 9)])
        assert res[(name, 9)][1] == 45 and res[(name, 8)][0] < 1e-13


def test_two_stage_tridiagonalisation_prototype():
    """tools/proto_sbr.py (round-2 candidate for csrc/eigh.cu): full -> band by block reflectors, band -> tridiagonal by bulge
    chasing (also in the wavefront order that lets ~n / 3b tasks run concurrently), back-transformation through both stages;
    eigenpairs agree with LAPACK for real and complex matrices, bandwidths below, at and above the matrix size."""
    import importlib.util, os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "proto_sbr.py")
    spec = importlib.util.spec_from_file_location("proto_sbr", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check(verbose=False) < 2e-14


@pytest.mark.parametrize("staged", [0, 1])
@pytest.mark.parametrize("n,b,wavefront", [(40, 4, 0), (61, 8, 1), (130, 32, 1), (97, 16, 0), (20, 32, 1)])
def test_bulge_chasing_on_band_storage(lib, n, b, wavefront, staged):
    """csrc/sbr_chase.h (round-2 groundwork, not wired into the library): one chasing task on packed band storage, the same
    code a device thread team would run.  All tasks in sweep or wavefront order turn a random symmetric band matrix into a
    tridiagonal one with the same eigenvalues; the stored reflectors give back the eigenvectors."""
    rng = np.random.default_rng(n + b)
    M = rng.standard_normal((n, n))
    A = np.tril(np.triu(M + M.T, -b), b)
    ld = 2 * b + 1
    ab = np.zeros((ld, n), order="F")
    for jj in range(n):
        m = min(b, n - 1 - jj)
        ab[:m + 1, jj] = A[jj:jj + m + 1, jj]
    nst = max(1, -(-(n - 1) // b))
    V2 = np.zeros((n, n), order="F")
    tau2 = np.zeros((nst, n), order="F")
    lib.sbr_chase_all.restype = C.c_int64
    ntask = lib.sbr_chase_all(C.c_int64(n), C.c_int(b), ab.ctypes.data_as(C.c_void_p), C.c_int64(ld), V2.ctypes.data_as(C.c_void_p),
                              tau2.ctypes.data_as(C.c_void_p), C.c_int64(nst), C.c_int(wavefront), C.c_int(staged))
    assert ntask == sum(max(0, -(-(n - jj - 1) // b)) for jj in range(n - 2))
    nrm = np.linalg.norm(A, 2)
    assert np.abs(ab[2:, :]).max() < 50 * EPS * nrm          # everything below the sub-diagonal is gone
    d, e = ab[0, :].copy(), ab[1, :n - 1].copy()
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    w, Z = np.linalg.eigh(T)
    assert np.abs(w - np.linalg.eigvalsh(A)).max() < 100 * EPS * nrm
    U = Z.copy()
    for jj in range(n - 3, -1, -1):                          # Q2 Z: reflectors in reverse order of application
        for s in range(max(0, -(-(n - jj - 1) // b)) - 1, -1, -1):
            r0 = jj + 1 + s * b
            r1 = min(r0 + b, n)
            if r1 - r0 < 2 or tau2[s, jj] == 0.0:
                continue
            v = V2[r0:r1, jj]
            U[r0:r1, :] -= tau2[s, jj] * np.outer(v, v @ U[r0:r1, :])
    assert np.linalg.norm(A @ U - U * w[None, :]) / (nrm * n) < 30 * EPS
    assert np.linalg.norm(U.T @ U - np.eye(n)) / n < 30 * EPS
