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

// Grouped GEMM (block-sparse sector products, K13): problem p is C_p = sum_{s in segments of p} op(A_s) op(B_s); all problems of
// one launch share opa / opb.  Offsets are element offsets into the three base buffers; tile0 is the prefix sum of output tiles
// (tiles_m x tiles_n per problem, tile sizes from gemm_tile_m / gemm_tile_n).  Tables live in device memory.
struct GroupedSegment { int64_t a_off, b_off, K, lda, ldb; int32_t alignedA, alignedB; };
struct GroupedProblem { int64_t c_off, M, N, ldc, tile0, tiles_m; int32_t seg0, nseg; };
template <typename T>
void gemm_grouped(Ctx* ctx, int opa, int opb, const GroupedProblem* probs_dev, int nprob, const GroupedSegment* segs_dev, int64_t total_tiles,
                  const T* Abase, const T* Bbase, T* Cbase, double flops);
int gemm_tile_m(bool cplx);
int gemm_tile_n(bool cplx);

const char* gemm_last_impl_name();

// Measured FP64 DMMA issue ceiling of this device (register-resident mma.sync loop), TFLOP/s.
double dmma_peak_tflops(Ctx* ctx);

}  // namespace nsb
