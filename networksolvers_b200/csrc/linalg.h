// Dense factorizations on device: Householder thin QR (gauge moves, K8) and the truncating
// factorization of the inserter (K9) by one-sided Jacobi (Hestenes) SVD.
#pragma once
#include <functional>

#include "common.h"

namespace nsb {

// NDTensors truncate! rule on a descending spectrum P (eigenvalues of rho = sigma^2).
// Returns number kept; truncerr = discarded weight / total weight.
int64_t truncate_spectrum(const std::vector<double>& P, double cutoff, int64_t mindim, int64_t maxdim, double* truncerr);

// Thin QR of A (rows x cols, lda; destroyed): Q (rows x k, ldq), R (k x cols, ldr), k = min(rows, cols).
template <typename T>
void qr_thin(Ctx* ctx, T* A, int64_t rows, int64_t cols, int64_t lda, T* Q, int64_t ldq, T* R, int64_t ldr);

// Batched 128 x 128 dense symmetric eigensolver (two-sided Jacobi inside one CTA); eigenvalue i <-> column i of R.
void herm_eig_batch128(Ctx* ctx, const double* S, double* R, double* evals, int batch, double abs_floor, int max_sweeps);


struct FactorInfo { int64_t newdim = 0; double truncerr = 0; int decomp = 0; int sweeps = 0; bool c_transposed = false; };

// Multi-GPU factorisation (identical input on every rank): the GEMM-shaped parts of the Gram + eigh route are split by
// column slabs, each slab computed by its owner and completed by an in-place all-gather (chunk of rank r at
// buf + r * bytes_per_rank).  The tridiagonalisation and the divide & conquer run replicated.  With allow_c_transposed the
// transposed-input case returns C^T (cols x newdim) instead of C (FactorInfo::c_transposed), which is the layout the
// caller's site tensor wants and splits over its columns.
struct FactorDist {
  int rank = 0, nranks = 1;
  std::function<void(void* buf, size_t bytes_per_rank)> allgather_inplace;
  bool allow_c_transposed = false;
};

// Truncated left-orthogonal factorization M = U C (src/inserter.jl:23 / ITensors.factorize with ortho="left"):
//   U (rows x newdim) orthonormal columns = leading left singular vectors, C = U^H M (newdim x cols).
// Spectrum (sigma^2 descending, all min(rows, cols) values) returned on the host.
// The decision rule cutoff <= 1e-12 -> "svd", else "eigen" only changes the label: both are computed by the
// same one-sided Jacobi iteration, which is at least as accurate as either LAPACK route.
// trans_in: the logical matrix is the transpose of the stored one, M(r, c) = buf[c + r * ld].
// sqrt_spectrum: truncate on sigma instead of sigma^2 (input is already a density matrix rho = S S^H, whose
//                singular values are the eigenvalues the reference's `eigen(rho; ...)` truncates on).
template <typename T>
FactorInfo factorize_left(Ctx* ctx, const T* M, int64_t rows, int64_t cols, int64_t ld, bool trans_in, double cutoff,
                          int64_t mindim, int64_t maxdim, bool sqrt_spectrum, DevBuf& U, DevBuf& C,
                          std::vector<double>& spectrum, const FactorDist* dist = nullptr);

}  // namespace nsb
