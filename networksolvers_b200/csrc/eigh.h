// Hermitian eigensolver on device for the truncating factorisation (K9/K10) at large bond dimension:
//   A = Q T Q^H   blocked Householder tridiagonalisation (panel recurrences + rank-2k GEMM update; the symmetric
//                 matrix-vector products are the HBM-bound part, n^3/3 * 8 bytes),
//   T = Z D Z^T   Cuppen divide & conquer (dc_secular.h; deflation planned on the host from O(n) data, secular
//                 roots / Gu-Eisenstat vectors in kernels, merges as GEMMs),
//   U = Q Z[:, sel]  compact-WY back-transformation of the selected eigenvectors only (GEMMs).
// Replaces the O(20 sweeps x 12 n^3) Jacobi iteration by O(10 n^3) work for n >= g_eigh_min_n.
#pragma once
#include "common.h"

namespace nsb {


template <typename T>
struct Eigh {
  Ctx* ctx = nullptr;
  int64_t n = 0;
  DevBuf Vall, taus, Z;        // reflectors (n x n, explicit unit diagonal), tau (n), eigenvectors of T (n x n, FP64)
  std::vector<double> w;       // eigenvalues; w[i] belongs to column i of Z (unsorted)
  int64_t dc_nondeflated = 0;  // sum of secular problem sizes over all merges (diagnostic)
  // A: n x n Hermitian, full storage (both triangles), column-major with leading dimension lda; destroyed.
  // lower_only_input: only the lower triangle of A holds valid data (the symmetric panel kernel reads nothing else;
  // otherwise the lower triangle is mirrored first).
  void factor(Ctx* ctx, T* A, int64_t n, int64_t lda, bool lower_only_input = false);
  // True when factor() will take the symmetric panel kernel for this problem, i.e. a caller that builds A by a GEMM
  // may skip the strictly upper output tiles (GEMM_LOWER_ONLY).
  static bool reads_lower_only(const Ctx* ctx, int64_t n, int64_t lda);
  // U (n x k, ldu) = Q Z[:, idx[0..k)]: eigenvectors of the original matrix for the chosen eigenvalues.
  void vectors(const int32_t* idx_host, int64_t k, T* U, int64_t ldu);
};

}  // namespace nsb
