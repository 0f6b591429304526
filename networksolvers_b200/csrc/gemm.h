// FP64 / complex-FP64 GEMM on sm_100a DMMA (mma.sync f64) -- interface.
#pragma once
#include "common.h"

namespace nsb {

enum GemmOp { OP_N = 0, OP_T = 1, OP_C = 2, OP_CONJ = 3 };
enum GemmImpl { GEMM_AUTO = 0, GEMM_NAIVE = 1, GEMM_DMMA = 2, GEMM_TMA = 3 };
enum GemmFlags { GEMM_LOWER_ONLY = 1 };   // tensor-core paths skip the 128-wide output tiles strictly above the diagonal

// Optional fused reduce-scatter epilogue: output column n belongs to rank n / slab_cols; the tile is written straight
// into that rank's peer-mapped staging window (NVLink P2P stores) at slot `rank`, instead of local C.
struct PeerOut {
  void* ptr[8] = {nullptr};   // peer-mapped window base of every rank (ptr[rank] is the local window)
  int nranks = 0, rank = 0;
  int64_t slab_cols = 0;      // columns per owner
  int64_t slab_elems = 0;     // M * slab_cols: one staging slot
};

// C[b] (M x N, ldc) = alpha * op(A[b]) (M x K) * op(B[b]) (K x N) + beta * C[b], column-major,
// b = 0..batch-1 with element strides strideA/B/C (0 = broadcast).
template <typename T>
void gemm(Ctx* ctx, int opa, int opb, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t lda,
          int64_t strideA, const T* B, int64_t ldb, int64_t strideB, T beta, T* C, int64_t ldc,
          int64_t strideC, int64_t batch, int impl = GEMM_AUTO, const PeerOut* peer = nullptr, int flags = 0);

const char* gemm_last_impl_name();

// Measured FP64 DMMA issue ceiling of this device (register-resident mma.sync loop), TFLOP/s.
double dmma_peak_tflops(Ctx* ctx);

}  // namespace nsb
