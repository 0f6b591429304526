"""NumPy statement of an Ozaki-scheme FP64 GEMM on 8-bit integer tensor cores (round-2 candidate for the H_eff GEMMs:
tcgen05 kind::i8 on sm_100a runs ~4.5 POP/s dense, the FP64 DMMA pipe 0.037 PFLOP/s).

  A (m x k) is scaled row-wise, B (k x n) column-wise, by powers of two so that every entry is < 1 in magnitude; each is cut
  into `s` slices of `bits` bits:  A = sum_i 2^(-bits (i + 1)) diag(2^ea) A_i,  A_i integer in [-2^bits, 2^bits]  (error-free:
  the slices are exact differences of truncations).  The products A_i B_j are exact in int32 as long as
  k 2^(2 bits) < 2^31; they are summed per anti-diagonal i + j = t in int32/int64 and the few anti-diagonals are combined in
  FP64, smallest first.  Keeping the pairs with i + j < s ("triangle") costs s (s + 1) / 2 integer GEMMs.

    python tools/proto_ozaki.py            # accuracy table against a long-double reference
Measured here (k = 2048 / 4096, 7-bit slices): 8 slices (36 integer GEMMs) 4e-15, 9 slices (45) 7e-17, FP64 GEMM 8e-16 --
i.e. FP64-equivalent accuracy costs ~40 int8 GEMMs, an ideal 4.5 POP/s / 40 = 112 TFLOP/s-equivalent against 37 on the DMMA pipe.
Used by tests/test_cpu_dc.py::test_ozaki_int8_gemm_matches_fp64."""
import numpy as np


def split(M, axis, slices, bits):
    """Row-wise (axis=1: one exponent per row) or column-wise (axis=0) error-free slicing into integer matrices."""
    amax = np.abs(M).max(axis=axis, keepdims=True)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)     # |M| 2^-e < 1/2
    R = M * np.exp2(-e)
    out = []
    for _ in range(slices):
        R = R * float(1 << bits)
        S = np.trunc(R)                     # |S| < 2^bits, exact
        out.append(S.astype(np.int64))
        R = R - S                           # exact remainder, |R| < 1
    return out, e


def ozaki_gemm(A, B, slices=9, bits=7, full=False):
    """C ~= A @ B from integer GEMMs only.  bits = 7 (int8 operands) keeps k < 2^17 exact in int32 accumulators."""
    As, ea = split(A, 1, slices, bits)
    Bs, eb = split(B, 0, slices, bits)
    k = A.shape[1]
    assert k * (1 << (2 * bits)) < (1 << 31), "int32 accumulator would overflow"
    ngemm = 0
    C = np.zeros((A.shape[0], B.shape[1]))
    tmax = 2 * slices - 1 if full else slices
    for t in range(tmax - 1, -1, -1):       # smallest contributions first
        acc = np.zeros((A.shape[0], B.shape[1]), dtype=np.int64)
        for i in range(slices):
            j = t - i
            if 0 <= j < slices:
                acc += As[i] @ Bs[j]        # exact integer GEMM (int8 operands, int32 accumulation on the device)
                ngemm += 1
        C += acc.astype(np.float64) * 2.0 ** (-bits * (t + 2))
    return C * np.exp2(ea) * np.exp2(eb), ngemm


def rel_err(C, Cref):
    return float(np.abs(C - Cref).max() / np.abs(Cref).max())


def check(m=96, k=2048, n=80, seed=0, verbose=True):
    rng = np.random.default_rng(seed)
    cases = {"gauss": (rng.standard_normal((m, k)), rng.standard_normal((k, n))),
             "graded": (rng.standard_normal((m, k)) * np.exp(-20.0 * rng.random((m, 1))), rng.standard_normal((k, n)) * np.exp(-20.0 * rng.random((1, n))))}
    res = {}
    for name, (A, B) in cases.items():
        ref = (A.astype(np.longdouble) @ B.astype(np.longdouble))
        e64 = rel_err((A @ B).astype(np.longdouble), ref)
        for s in (7, 8, 9):
            C, ng = ozaki_gemm(A, B, slices=s, bits=7)
            res[(name, s)] = (rel_err(C.astype(np.longdouble), ref), ng)
            if verbose:
                print(f"{name:7s} slices {s:2d} ({ng:3d} int8 GEMMs): max rel err {res[(name, s)][0]:.2e}   (FP64 GEMM: {e64:.2e})")
        res[(name, "fp64")] = (e64, 1)
    return res


if __name__ == "__main__":
    check()
