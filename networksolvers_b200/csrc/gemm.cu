// FP64 / complex-FP64 GEMM for sm_100a.
//
// Blackwell's tcgen05/UMMA has no FP64 kind, so FP64 tensor-core work is the warp-level DMMA path:
// `mma.sync.aligned.m16n8k8.row.col.f64` (SASS: 4 x DMMA.8x8x4).  Measured DMMA peak on B200 is
// 36.9 TFLOP/s (profiles/r01_microbench_fp64_peaks.jsonl), cuBLAS DGEMM reaches 35.5-36.0.
//
// Kernels in this file
//   gemm_naive_kernel   smem-tiled DFMA reference (tiny problems, on-device cross-check)
//   gemm_dmma_kernel    persistent, 4-stage cp.async pipeline, 128x128x16 (real) / 128x64x8 (complex)
//                       CTA tiles, 8 MMA warps, 128B-swizzled shared-memory rows
//   gemm_tma_kernel     same consumer mainloop, operands fetched by TMA (cp.async.bulk.tensor) from
//                       a dedicated producer warp through full/empty mbarriers
// Shared-memory tile layout (both DMMA kernels): rows of 128 bytes, 16-byte chunk index XOR-ed with
// (row & 7) -- exactly what CU_TENSOR_MAP_SWIZZLE_128B produces.  "K-major" operand tiles hold one
// M/N index per row (BK contiguous k values); "MN-major" tiles are split into boxes of 128B-wide
// rows (one k per row).  The k slots of the MMA fragments are permuted (thread t owns k = 4t+2h+j)
// so that K-major fragment loads are conflict-free 128-bit LDS; the permutation is applied to A and
// B alike, which leaves the product unchanged.
#include "gemm.h"

#include <cuda.h>
#include <mutex>

namespace nsb {

static thread_local const char* g_last_impl = "none";
const char* gemm_last_impl_name() { return g_last_impl; }

template <typename T>
struct GemmParams {
  const T* A; const T* B; T* C;
  int64_t M, N, K, lda, ldb, ldc, strideA, strideB, strideC, batch;
  T alpha, beta;
  int a_kmajor, b_kmajor;   // storage orientation of op(A), op(B)
  double sa, sb;            // -1 to conjugate (complex only)
  int alignedA, alignedB;   // 16-byte alignment of every chunk source
  int bcoordA, bcoordB;     // 0 when the operand is broadcast over the batch (stride 0)
  int64_t tiles_m, tiles_n;
  int lower_only;           // skip output tiles strictly above the diagonal (symmetric rank-k updates that only need the lower triangle)
  PeerOut peer;             // nranks > 1: fused reduce-scatter epilogue over peer memory
};

// ------------------------------------------------------------------------------------------------
// naive kernel: 16x16 tiles, generic element accessors
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void gemm_naive_kernel(GemmParams<T> p) {
  __shared__ T sA[16][17];
  __shared__ T sB[16][17];
  int64_t b = blockIdx.z;
  const T* A = p.A + b * p.strideA;
  const T* B = p.B + b * p.strideB;
  T* C = p.C + b * p.strideC;
  int tx = threadIdx.x, ty = threadIdx.y;
  int64_t m = (int64_t)blockIdx.x * 16 + tx;
  int64_t n = (int64_t)blockIdx.y * 16 + ty;
  T acc = zero_<T>();
  for (int64_t k0 = 0; k0 < p.K; k0 += 16) {
    {  // A tile element (m = tile_m + tx, k = k0 + ty)
      int64_t k = k0 + ty;
      T v = zero_<T>();
      if (m < p.M && k < p.K) {
        v = p.a_kmajor ? A[k + m * p.lda] : A[m + k * p.lda];
        if (p.sa < 0) v = conj_(v);
      }
      sA[ty][tx] = v;
    }
    {  // B tile element (k = k0 + tx, n = tile_n + ty)
      int64_t k = k0 + tx;
      T v = zero_<T>();
      if (n < p.N && k < p.K) {
        v = p.b_kmajor ? B[k + n * p.ldb] : B[n + k * p.ldb];
        if (p.sb < 0) v = conj_(v);
      }
      sB[ty][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) fma_(acc, sA[kk][tx], sB[ty][kk]);
    __syncthreads();
  }
  if (m < p.M && n < p.N) {
    T r = mul_(p.alpha, acc);
    if (re(p.beta) != 0.0 || im(p.beta) != 0.0) r = add_(r, mul_(p.beta, C[m + n * p.ldc]));
    C[m + n * p.ldc] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// DMMA building blocks
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16x8x8(double (&d)[4], double a0, double a1, double a2, double a3,
                                           double b0, double b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <typename T> struct TileCfg;
template <> struct TileCfg<double> {
  static constexpr int BM = 128, BN = 128, BK = 16, EPR = 16, EPC = 2;
  static constexpr int WM = 64, WN = 32, MI = 4, NI = 4, WARPS_M = 2;
};
template <> struct TileCfg<cdouble> {
  static constexpr int BM = 128, BN = 64, BK = 8, EPR = 8, EPC = 1;
  static constexpr int WM = 32, WN = 32, MI = 2, NI = 4, WARPS_M = 4;
};
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_STAGES = 4;

// Load one operand tile (ROWS m- or n-indices x BK k-values) with cp.async into the swizzled layout.
template <typename T, int ROWS>
__device__ __forceinline__ void load_tile_cpasync(char* sT, const T* __restrict__ G, int64_t ld, int64_t r0,
                                                  int64_t k0, int64_t R, int64_t K, bool kmajor, bool al16,
                                                  int tid) {
  constexpr int EPC = TileCfg<T>::EPC;
  constexpr int EPR = TileCfg<T>::EPR;
  constexpr int NCHUNK = ROWS * 8;
  const uint32_t sbase = smem_u32(sT);
#pragma unroll
  for (int it = 0; it < NCHUNK / GEMM_THREADS; ++it) {
    int cid = tid + it * GEMM_THREADS;
    int64_t goff;
    uint32_t soff;
    int valid;
    if (kmajor) {
      int row = cid >> 3, c = cid & 7;
      int64_t r = r0 + row, k = k0 + (int64_t)c * EPC;
      int64_t rem = K - k;
      valid = (r < R && rem > 0) ? (rem >= EPC ? EPC : (int)rem) : 0;
      goff = r * ld + k;
      soff = row * 128 + ((c ^ (row & 7)) << 4);
    } else {
      int box = cid / (EPR * 8), rem_c = cid % (EPR * 8);
      int krow = rem_c >> 3, c = rem_c & 7;
      int64_t r = r0 + (int64_t)box * EPR + (int64_t)c * EPC, k = k0 + krow;
      int64_t rem = R - r;
      valid = (k < K && rem > 0) ? (rem >= EPC ? EPC : (int)rem) : 0;
      goff = k * ld + r;
      soff = box * (EPR * 128) + krow * 128 + ((c ^ (krow & 7)) << 4);
    }
    const T* src = valid ? (G + goff) : G;
    if (al16) {
      cp_async16(sbase + soff, src, valid * (int)sizeof(T));
    } else {  // real only: two 8-byte copies
      cp_async8(sbase + soff, src, valid >= 1 ? 8 : 0);
      cp_async8(sbase + soff + 8, valid >= 2 ? (const void*)(src + 1) : (const void*)G, valid >= 2 ? 8 : 0);
    }
  }
}

// ---- consumer: real -----------------------------------------------------------------------------
struct AccReal { double v[4][4][4]; };

template <bool AK, bool BKM>
__device__ __forceinline__ void compute_stage(const char* sA, const char* sB, AccReal& acc, int wm, int wn,
                                              int g, int t, double, double) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double b0[4], b1[4];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      if (BKM) {
        int row = wn * 32 + ni * 8 + g;
        int chunk = (2 * t + h) ^ (row & 7);
        double2 v = *reinterpret_cast<const double2*>(sB + row * 128 + chunk * 16);
        b0[ni] = v.x; b1[ni] = v.y;
      } else {
        int n = wn * 32 + ni * 8 + g;
        int box = n >> 4, nn = n & 15;
        int k = 4 * t + 2 * h;
        b0[ni] = *reinterpret_cast<const double*>(sB + box * 2048 + k * 128 + (((nn >> 1) ^ (k & 7)) << 4) + (nn & 1) * 8);
        b1[ni] = *reinterpret_cast<const double*>(sB + box * 2048 + (k + 1) * 128 + (((nn >> 1) ^ ((k + 1) & 7)) << 4) + (nn & 1) * 8);
      }
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      double a0, a1, a2, a3;
      if (AK) {
        int r0 = wm * 64 + mi * 16 + g;
        int chunk = (2 * t + h) ^ (r0 & 7);
        double2 v0 = *reinterpret_cast<const double2*>(sA + r0 * 128 + chunk * 16);
        double2 v1 = *reinterpret_cast<const double2*>(sA + (r0 + 8) * 128 + chunk * 16);
        a0 = v0.x; a1 = v1.x; a2 = v0.y; a3 = v1.y;
      } else {
        int box = wm * 4 + mi;
        int k = 4 * t + 2 * h;
        const char* bp = sA + box * 2048;
        int c0 = (g >> 1), c1 = ((g + 8) >> 1), lo = (g & 1) * 8;
        a0 = *reinterpret_cast<const double*>(bp + k * 128 + ((c0 ^ (k & 7)) << 4) + lo);
        a1 = *reinterpret_cast<const double*>(bp + k * 128 + ((c1 ^ (k & 7)) << 4) + lo);
        a2 = *reinterpret_cast<const double*>(bp + (k + 1) * 128 + ((c0 ^ ((k + 1) & 7)) << 4) + lo);
        a3 = *reinterpret_cast<const double*>(bp + (k + 1) * 128 + ((c1 ^ ((k + 1) & 7)) << 4) + lo);
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) mma_16x8x8(acc.v[mi][ni], a0, a1, a2, a3, b0[ni], b1[ni]);
    }
  }
}

__device__ __forceinline__ void acc_zero(AccReal& a) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) a.v[i][j][e] = 0.0;
}

__device__ __forceinline__ void store_tile(const AccReal& acc, const GemmParams<double>& p, double* C, int64_t m0,
                                           int64_t n0, int wm, int wn, int g, int t) {
  const bool has_beta = (p.beta != 0.0);
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int64_t m = m0 + wm * 64 + mi * 16 + g + ((e >> 1) ? 8 : 0);
        int64_t n = n0 + wn * 32 + ni * 8 + 2 * t + (e & 1);
        if (m < p.M && n < p.N) {
          double r = p.alpha * acc.v[mi][ni][e];
          if (p.peer.nranks > 1) {
            int64_t owner = n / p.peer.slab_cols;
            double* dst = reinterpret_cast<double*>(p.peer.ptr[owner]) + p.peer.rank * p.peer.slab_elems;
            dst[m + (n - owner * p.peer.slab_cols) * p.M] = r;
          } else {
            if (has_beta) r += p.beta * C[m + n * p.ldc];
            C[m + n * p.ldc] = r;
          }
        }
      }
}

// ---- consumer: complex --------------------------------------------------------------------------
struct AccCplx { double re[2][4][4]; double im[2][4][4]; };

template <bool AK, bool BKM>
__device__ __forceinline__ void compute_stage(const char* sA, const char* sB, AccCplx& acc, int wm, int wn,
                                              int g, int t, double sa, double sb) {
  double br0[4], br1[4], bi0[4], bi1[4];
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) {
    double2 v0, v1;
    if (BKM) {
      int row = wn * 32 + ni * 8 + g;
      v0 = *reinterpret_cast<const double2*>(sB + row * 128 + (((2 * t) ^ (row & 7)) << 4));
      v1 = *reinterpret_cast<const double2*>(sB + row * 128 + (((2 * t + 1) ^ (row & 7)) << 4));
    } else {
      int box = wn * 4 + ni;
      int k = 2 * t;
      v0 = *reinterpret_cast<const double2*>(sB + box * 1024 + k * 128 + ((g ^ (k & 7)) << 4));
      v1 = *reinterpret_cast<const double2*>(sB + box * 1024 + (k + 1) * 128 + ((g ^ ((k + 1) & 7)) << 4));
    }
    br0[ni] = v0.x; bi0[ni] = v0.y * sb; br1[ni] = v1.x; bi1[ni] = v1.y * sb;
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    double2 v00, v01, v10, v11;  // v{rowhalf}{kslot}
    if (AK) {
      int r0 = wm * 32 + mi * 16 + g;
      int x = r0 & 7;
      v00 = *reinterpret_cast<const double2*>(sA + r0 * 128 + (((2 * t) ^ x) << 4));
      v01 = *reinterpret_cast<const double2*>(sA + r0 * 128 + (((2 * t + 1) ^ x) << 4));
      v10 = *reinterpret_cast<const double2*>(sA + (r0 + 8) * 128 + (((2 * t) ^ x) << 4));
      v11 = *reinterpret_cast<const double2*>(sA + (r0 + 8) * 128 + (((2 * t + 1) ^ x) << 4));
    } else {
      int box = wm * 4 + mi * 2;
      int k = 2 * t;
      v00 = *reinterpret_cast<const double2*>(sA + box * 1024 + k * 128 + ((g ^ (k & 7)) << 4));
      v01 = *reinterpret_cast<const double2*>(sA + box * 1024 + (k + 1) * 128 + ((g ^ ((k + 1) & 7)) << 4));
      v10 = *reinterpret_cast<const double2*>(sA + (box + 1) * 1024 + k * 128 + ((g ^ (k & 7)) << 4));
      v11 = *reinterpret_cast<const double2*>(sA + (box + 1) * 1024 + (k + 1) * 128 + ((g ^ ((k + 1) & 7)) << 4));
    }
    double ar0 = v00.x, ar1 = v10.x, ar2 = v01.x, ar3 = v11.x;
    double ai0 = v00.y * sa, ai1 = v10.y * sa, ai2 = v01.y * sa, ai3 = v11.y * sa;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      mma_16x8x8(acc.re[mi][ni], ar0, ar1, ar2, ar3, br0[ni], br1[ni]);
      mma_16x8x8(acc.re[mi][ni], -ai0, -ai1, -ai2, -ai3, bi0[ni], bi1[ni]);
      mma_16x8x8(acc.im[mi][ni], ar0, ar1, ar2, ar3, bi0[ni], bi1[ni]);
      mma_16x8x8(acc.im[mi][ni], ai0, ai1, ai2, ai3, br0[ni], br1[ni]);
    }
  }
}

__device__ __forceinline__ void acc_zero(AccCplx& a) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { a.re[i][j][e] = 0.0; a.im[i][j][e] = 0.0; }
}

__device__ __forceinline__ void store_tile(const AccCplx& acc, const GemmParams<cdouble>& p, cdouble* C,
                                           int64_t m0, int64_t n0, int wm, int wn, int g, int t) {
  const bool has_beta = (p.beta.x != 0.0 || p.beta.y != 0.0);
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int64_t m = m0 + wm * 32 + mi * 16 + g + ((e >> 1) ? 8 : 0);
        int64_t n = n0 + wn * 32 + ni * 8 + 2 * t + (e & 1);
        if (m < p.M && n < p.N) {
          cdouble r = mul_(p.alpha, make_cuDoubleComplex(acc.re[mi][ni][e], acc.im[mi][ni][e]));
          if (p.peer.nranks > 1) {
            int64_t owner = n / p.peer.slab_cols;
            cdouble* dst = reinterpret_cast<cdouble*>(p.peer.ptr[owner]) + p.peer.rank * p.peer.slab_elems;
            dst[m + (n - owner * p.peer.slab_cols) * p.M] = r;
          } else {
            if (has_beta) r = add_(r, mul_(p.beta, C[m + n * p.ldc]));
            C[m + n * p.ldc] = r;
          }
        }
      }
}

template <typename T> struct AccOf;
template <> struct AccOf<double> { typedef AccReal type; };
template <> struct AccOf<cdouble> { typedef AccCplx type; };

__device__ __forceinline__ void tile_coords(int64_t tile, int64_t tiles_m, int64_t tiles_n, int64_t& tm,
                                            int64_t& tn) {
  // grouped rasterisation: bands of GROUP_M row-tiles, m fastest inside a band, so one wave of 148 CTAs
  // touches ~12 A panels x ~12 B panels (L2-resident)
  const int64_t GROUP_M = 12;
  int64_t group_size = GROUP_M * tiles_n;
  int64_t group = tile / group_size;
  int64_t first_m = group * GROUP_M;
  int64_t gm = tiles_m - first_m < GROUP_M ? tiles_m - first_m : GROUP_M;
  int64_t r = tile % group_size;
  tm = first_m + r % gm;
  tn = r / gm;
}

// ------------------------------------------------------------------------------------------------
// gemm_dmma_kernel: cp.async multistage pipeline
// ------------------------------------------------------------------------------------------------
template <typename T, bool AK, bool BKM>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_dmma_kernel(const GemmParams<T> p) {
  typedef TileCfg<T> Cfg;
  typedef typename AccOf<T>::type Acc;
  extern __shared__ __align__(1024) char smem[];
  constexpr int A_BYTES = Cfg::BM * 128, B_BYTES = Cfg::BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp % Cfg::WARPS_M, wn = warp / Cfg::WARPS_M;
  const int64_t tiles_per_batch = p.tiles_m * p.tiles_n;
  const int64_t total = tiles_per_batch * p.batch;
  const int64_t nk = (p.K + Cfg::BK - 1) / Cfg::BK;
  const bool alA = p.alignedA, alB = p.alignedB;

  for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int64_t b = tile / tiles_per_batch;
    int64_t tm, tn;
    tile_coords(tile % tiles_per_batch, p.tiles_m, p.tiles_n, tm, tn);
    const int64_t m0 = tm * Cfg::BM, n0 = tn * Cfg::BN;
    if (p.lower_only && n0 >= m0 + Cfg::BM) continue;
    const T* A = p.A + b * p.strideA;
    const T* B = p.B + b * p.strideB;
    Acc acc;
    acc_zero(acc);

#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; ++s) {
      if (s < nk) {
        char* st = smem + s * STAGE_BYTES;
        load_tile_cpasync<T, Cfg::BM>(st, A, p.lda, m0, (int64_t)s * Cfg::BK, p.M, p.K, AK, alA, tid);
        load_tile_cpasync<T, Cfg::BN>(st + A_BYTES, B, p.ldb, n0, (int64_t)s * Cfg::BK, p.N, p.K, BKM, alB, tid);
      }
      cp_async_commit();
    }
    for (int64_t kt = 0; kt < nk; ++kt) {
      cp_async_wait<GEMM_STAGES - 2>();
      __syncthreads();
      int64_t kn = kt + GEMM_STAGES - 1;
      if (kn < nk) {
        char* st = smem + (kn % GEMM_STAGES) * STAGE_BYTES;
        load_tile_cpasync<T, Cfg::BM>(st, A, p.lda, m0, kn * Cfg::BK, p.M, p.K, AK, alA, tid);
        load_tile_cpasync<T, Cfg::BN>(st + A_BYTES, B, p.ldb, n0, kn * Cfg::BK, p.N, p.K, BKM, alB, tid);
      }
      cp_async_commit();
      const char* cs = smem + (kt % GEMM_STAGES) * STAGE_BYTES;
      compute_stage<AK, BKM>(cs, cs + A_BYTES, acc, wm, wn, g, t, p.sa, p.sb);
    }
    cp_async_wait<0>();
    __syncthreads();  // all warps done with smem before the next tile's prologue overwrites it
    store_tile(acc, p, p.C + b * p.strideC, m0, n0, wm, wn, g, t);
  }
}

// ------------------------------------------------------------------------------------------------
// gemm_grouped_kernel: one persistent launch for a whole list of independent GEMMs of unequal size -- the symmetry-sector
// block products of a QN-conserving contraction (K13).  Problem p is C_p (M_p x N_p) = sum over its segments s of
// op(A_s) (M_p x K_s) op(B_s) (K_s x N_p): the segments are the sector pairs that contribute to one output block, summed in
// the accumulator registers, so every output tile is written exactly once (no beta passes, no atomics).  Operands are
// offsets into three base buffers (block-sparse flat storage).  Same DMMA consumer and cp.async pipeline as
// gemm_dmma_kernel; output tiles of all problems form one work list dealt round-robin to the CTAs.
// ------------------------------------------------------------------------------------------------
template <typename T, bool AK, bool BKM>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_grouped_kernel(const GroupedProblem* __restrict__ probs, int nprob,
                                                                      const GroupedSegment* __restrict__ segs, int64_t total_tiles,
                                                                      const T* __restrict__ Abase, const T* __restrict__ Bbase,
                                                                      T* __restrict__ Cbase, GemmParams<T> proto) {
  typedef TileCfg<T> Cfg;
  typedef typename AccOf<T>::type Acc;
  extern __shared__ __align__(1024) char smem[];
  constexpr int A_BYTES = Cfg::BM * 128, B_BYTES = Cfg::BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp % Cfg::WARPS_M, wn = warp / Cfg::WARPS_M;
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    int lo = 0, hi = nprob - 1;       // problem of this tile: last p with tile0 <= tile
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (probs[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
    }
    const GroupedProblem pr = probs[lo];
    const int64_t lt = tile - pr.tile0;
    const int64_t tm = lt % pr.tiles_m, tn = lt / pr.tiles_m;
    const int64_t m0 = tm * Cfg::BM, n0 = tn * Cfg::BN;
    Acc acc;
    acc_zero(acc);
    for (int sidx = pr.seg0; sidx < pr.seg0 + pr.nseg; ++sidx) {
      const GroupedSegment sg = segs[sidx];
      const T* A = Abase + sg.a_off;
      const T* B = Bbase + sg.b_off;
      const int64_t nk = (sg.K + Cfg::BK - 1) / Cfg::BK;
      const bool alA = sg.alignedA, alB = sg.alignedB;
#pragma unroll
      for (int s = 0; s < GEMM_STAGES - 1; ++s) {
        if (s < nk) {
          char* st = smem + s * STAGE_BYTES;
          load_tile_cpasync<T, Cfg::BM>(st, A, sg.lda, m0, (int64_t)s * Cfg::BK, pr.M, sg.K, AK, alA, tid);
          load_tile_cpasync<T, Cfg::BN>(st + A_BYTES, B, sg.ldb, n0, (int64_t)s * Cfg::BK, pr.N, sg.K, BKM, alB, tid);
        }
        cp_async_commit();
      }
      for (int64_t kt = 0; kt < nk; ++kt) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        const int64_t kn = kt + GEMM_STAGES - 1;
        if (kn < nk) {
          char* st = smem + (kn % GEMM_STAGES) * STAGE_BYTES;
          load_tile_cpasync<T, Cfg::BM>(st, A, sg.lda, m0, kn * Cfg::BK, pr.M, sg.K, AK, alA, tid);
          load_tile_cpasync<T, Cfg::BN>(st + A_BYTES, B, sg.ldb, n0, kn * Cfg::BK, pr.N, sg.K, BKM, alB, tid);
        }
        cp_async_commit();
        const char* cs = smem + (kt % GEMM_STAGES) * STAGE_BYTES;
        compute_stage<AK, BKM>(cs, cs + A_BYTES, acc, wm, wn, g, t, proto.sa, proto.sb);
      }
      cp_async_wait<0>();
      __syncthreads();  // all warps done with smem before the next segment's prologue overwrites it
    }
    GemmParams<T> p = proto;
    p.M = pr.M; p.N = pr.N; p.ldc = pr.ldc;
    store_tile(acc, p, Cbase + pr.c_off, m0, n0, wm, wn, g, t);
  }
}

// ------------------------------------------------------------------------------------------------
// gemm_tma_kernel: TMA producer warp + 8 DMMA consumer warps, full/empty mbarrier ring
// ------------------------------------------------------------------------------------------------
struct TmaMaps { CUtensorMap a; CUtensorMap b; };

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
          "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

constexpr int TMA_THREADS = GEMM_THREADS + 128;  // 2 consumer warpgroups + 1 producer warpgroup (setmaxnreg split)

template <typename T, bool AK, bool BKM>
__global__ void __launch_bounds__(TMA_THREADS, 1)
gemm_tma_kernel(const GemmParams<T> p, const __grid_constant__ TmaMaps maps) {
  typedef TileCfg<T> Cfg;
  typedef typename AccOf<T>::type Acc;
  extern __shared__ __align__(1024) char smem[];
  constexpr int A_BYTES = Cfg::BM * 128, B_BYTES = Cfg::BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * STAGE_BYTES);
  uint64_t* empty = full + GEMM_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tiles_per_batch = p.tiles_m * p.tiles_n;
  const int64_t total = tiles_per_batch * p.batch;
  const int64_t nk = (p.K + Cfg::BK - 1) / Cfg::BK;

  if (tid == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GEMM_THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= GEMM_THREADS / 32) {
    // ---------------- producer warpgroup (one elected lane issues TMA) ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (warp == GEMM_THREADS / 32 && lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int64_t b = tile / tiles_per_batch;
        int64_t tm, tn;
        tile_coords(tile % tiles_per_batch, p.tiles_m, p.tiles_n, tm, tn);
        const int m0 = (int)(tm * Cfg::BM), n0 = (int)(tn * Cfg::BN);
        if (p.lower_only && n0 >= m0 + Cfg::BM) continue;
        for (int64_t kt = 0; kt < nk; ++kt, ++it) {
          int s = it % GEMM_STAGES;
          uint32_t ph = (it / GEMM_STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          char* st = smem + s * STAGE_BYTES;
          const int k0 = (int)(kt * Cfg::BK);
          constexpr int F = (int)(sizeof(T) / 8);  // FP64 words per element (inner coordinate unit)
          const int ba = (int)b * p.bcoordA, bb = (int)b * p.bcoordB;
          if (AK) {
            tma_load_3d(st, &maps.a, &full[s], k0 * F, m0, ba);
          } else {
#pragma unroll
            for (int bx = 0; bx < Cfg::BM / Cfg::EPR; ++bx)
              tma_load_3d(st + bx * (Cfg::EPR * 128), &maps.a, &full[s], (m0 + bx * Cfg::EPR) * F, k0, ba);
          }
          if (BKM) {
            tma_load_3d(st + A_BYTES, &maps.b, &full[s], k0 * F, n0, bb);
          } else {
#pragma unroll
            for (int bx = 0; bx < Cfg::BN / Cfg::EPR; ++bx)
              tma_load_3d(st + A_BYTES + bx * (Cfg::EPR * 128), &maps.b, &full[s], (n0 + bx * Cfg::EPR) * F, k0, bb);
          }
        }
      }
    }
    return;
  }
  // ---------------- consumer warpgroups ----------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp % Cfg::WARPS_M, wn = warp / Cfg::WARPS_M;
  uint32_t it = 0;
  for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int64_t b = tile / tiles_per_batch;
    int64_t tm, tn;
    tile_coords(tile % tiles_per_batch, p.tiles_m, p.tiles_n, tm, tn);
    const int64_t m0 = tm * Cfg::BM, n0 = tn * Cfg::BN;
    if (p.lower_only && n0 >= m0 + Cfg::BM) continue;
    Acc acc;
    acc_zero(acc);
    for (int64_t kt = 0; kt < nk; ++kt, ++it) {
      int s = it % GEMM_STAGES;
      uint32_t ph = (it / GEMM_STAGES) & 1;
      mbar_wait(&full[s], ph);
      const char* cs = smem + s * STAGE_BYTES;
      compute_stage<AK, BKM>(cs, cs + A_BYTES, acc, wm, wn, g, t, p.sa, p.sb);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    store_tile(acc, p, p.C + b * p.strideC, m0, n0, wm, wn, g, t);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// Build a 3-D tensor map {inner, outer, batch} over `ptr` with 128B swizzle; elements are FP64 (a complex
// element is two consecutive FP64 values, so inner extents are doubled for cdouble).
template <typename T>
static bool make_map(CUtensorMap* map, const T* ptr, int64_t inner, int64_t outer, int64_t ld, int64_t batch,
                     int64_t bstride, int box_inner, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const int f = sizeof(T) / 8;
  if (bstride == 0) batch = 1;  // broadcast operand: single slice, kernel passes batch coordinate 0
  cuuint64_t dims[3] = {(cuuint64_t)(inner * f), (cuuint64_t)outer, (cuuint64_t)(batch > 0 ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)(ld * sizeof(T)), (cuuint64_t)((bstride > 0 ? bstride : ld * outer) * sizeof(T))};
  if (batch <= 1) strides[1] = (cuuint64_t)(ld * (outer > 0 ? outer : 1)) * sizeof(T);
  if (strides[0] % 16 || strides[1] % 16 || ((uintptr_t)ptr % 16)) return false;
  if (strides[0] >= (1ull << 40) || strides[1] >= (1ull << 40)) return false;
  cuuint32_t box[3] = {(cuuint32_t)(box_inner * f), (cuuint32_t)box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <typename T, bool AK, bool BKM>
static void launch_dmma(Ctx* ctx, const GemmParams<T>& p) {
  typedef TileCfg<T> Cfg;
  size_t smem = (size_t)GEMM_STAGES * (Cfg::BM + Cfg::BN) * 128;
  auto kern = gemm_dmma_kernel<T, AK, BKM>;
  static bool configured_dev[64] = {false};
  bool& configured = configured_dev[ctx->device & 63];   // function attributes are per device
  if (!configured) {
    NSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int64_t total = p.tiles_m * p.tiles_n * p.batch;
  int grid = (int)std::min<int64_t>(total, ctx->num_sms);
  kern<<<grid, GEMM_THREADS, smem, ctx->stream>>>(p);
  NSB_CUDA(cudaGetLastError());
}

template <typename T, bool AK, bool BKM>
static bool launch_tma(Ctx* ctx, const GemmParams<T>& p) {
  typedef TileCfg<T> Cfg;
  TmaMaps maps;
  bool ok;
  if (AK) ok = make_map<T>(&maps.a, p.A, p.K, p.M, p.lda, p.batch, p.strideA, Cfg::BK, Cfg::BM);
  else    ok = make_map<T>(&maps.a, p.A, p.M, p.K, p.lda, p.batch, p.strideA, Cfg::EPR, Cfg::BK);
  if (!ok) return false;
  if (BKM) ok = make_map<T>(&maps.b, p.B, p.K, p.N, p.ldb, p.batch, p.strideB, Cfg::BK, Cfg::BN);
  else     ok = make_map<T>(&maps.b, p.B, p.N, p.K, p.ldb, p.batch, p.strideB, Cfg::EPR, Cfg::BK);
  if (!ok) return false;
  size_t smem = (size_t)GEMM_STAGES * (Cfg::BM + Cfg::BN) * 128 + 2 * GEMM_STAGES * sizeof(uint64_t);
  auto kern = gemm_tma_kernel<T, AK, BKM>;
  static bool configured_dev[64] = {false};
  bool& configured = configured_dev[ctx->device & 63];   // function attributes are per device
  if (!configured) {
    NSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int64_t total = p.tiles_m * p.tiles_n * p.batch;
  int grid = (int)std::min<int64_t>(total, ctx->num_sms);
  kern<<<grid, TMA_THREADS, smem, ctx->stream>>>(p, maps);
  NSB_CUDA(cudaGetLastError());
  return true;
}

template <typename T>
void gemm(Ctx* ctx, int opa, int opb, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t lda,
          int64_t strideA, const T* B, int64_t ldb, int64_t strideB, T beta, T* C, int64_t ldc,
          int64_t strideC, int64_t batch, int impl, const PeerOut* peer, int flags) {
  if (M <= 0 || N <= 0 || batch <= 0) return;
  NSB_REQUIRE(K >= 0, NSB_EINVAL, "gemm: negative K");
  typedef TileCfg<T> Cfg;
  GemmParams<T> p;
  p.A = A; p.B = B; p.C = C;
  p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.strideA = strideA; p.strideB = strideB; p.strideC = strideC; p.batch = batch;
  p.alpha = alpha; p.beta = beta;
  p.lower_only = (flags & GEMM_LOWER_ONLY) ? 1 : 0;
  if (peer) { NSB_REQUIRE(batch == 1 && peer->nranks <= 8, NSB_EINVAL, "gemm: peer epilogue needs batch == 1"); p.peer = *peer; }
  p.a_kmajor = (opa == OP_T || opa == OP_C);
  p.b_kmajor = (opb == OP_N || opb == OP_CONJ);
  p.sa = (ScalarTraits<T>::is_complex && (opa == OP_C || opa == OP_CONJ)) ? -1.0 : 1.0;
  p.sb = (ScalarTraits<T>::is_complex && (opb == OP_C || opb == OP_CONJ)) ? -1.0 : 1.0;
  p.tiles_m = (M + Cfg::BM - 1) / Cfg::BM;
  p.tiles_n = (N + Cfg::BN - 1) / Cfg::BN;
  auto aligned = [](const void* ptr, int64_t ld, int64_t stride) {
    return ((uintptr_t)ptr % 16 == 0) && ((ld * sizeof(T)) % 16 == 0) && ((stride * sizeof(T)) % 16 == 0);
  };
  p.alignedA = aligned(A, lda, strideA);
  p.alignedB = aligned(B, ldb, strideB);
  p.bcoordA = (strideA != 0 && batch > 1) ? 1 : 0;
  p.bcoordB = (strideB != 0 && batch > 1) ? 1 : 0;

  if (impl == GEMM_AUTO) impl = ctx->gemm_impl;
  if (peer && (impl == GEMM_AUTO || impl == GEMM_NAIVE)) impl = GEMM_TMA;
  if (impl == GEMM_AUTO) {
    // measured on B200 in a stream of back-to-back calls (profiles/r02_perf_small_gemm.log): the persistent DMMA kernels have
    // a floor of 8 us plus ~1 us per k-step of a single tile and need their tensor maps encoded on the host per call; the plain
    // one-thread-per-element kernel is faster up to m n k ~ 4e7 (166 x 664 x 166: 17 us against 31 us; 49 x 196 x 196: 11
    // against 33 .. 43 us) -- the launch-bound small-chi regime runs on it
    double work = (double)M * (double)N * (double)(K > 0 ? K : 1) * (double)batch;
    const bool naive_ok = batch <= 65535 && (N + 15) / 16 <= 65535;   // (grid limits of the plain kernel)
    impl = (naive_ok && work < (double)ctx->opt.gemm_naive_max_work) ? GEMM_NAIVE : GEMM_TMA;
  }
  ctx->cnt.gemm_calls++;
  ctx->cnt.kernel_launches++;
  double frac = 1.0;
  if (p.lower_only) {   // executed flops: only the tiles that are not skipped
    int64_t kept = 0;
    for (int64_t tm = 0; tm < p.tiles_m; ++tm) kept += std::min<int64_t>(p.tiles_n, (tm * Cfg::BM + Cfg::BM - 1) / Cfg::BN + 1);
    frac = (double)kept / (double)(p.tiles_m * p.tiles_n);
  }
  const double flops_issued = frac * (ScalarTraits<T>::is_complex ? 8.0 : 2.0) * (double)M * (double)N * (double)K * (double)batch;
  ctx->cnt.gemm_flops += flops_issued;
  struct ProfScope {   // events around this launch (all return paths)
    Ctx* c; size_t idx; bool on;
    ProfScope(Ctx* c_, double fl, int64_t M_, int64_t N_, int64_t K_, int64_t b_) : c(c_), idx(0), on(c_->gemm_profile && c_->gemm_prof.size() < 65536) {
      if (!on) return;
      Ctx::GemmProf g; g.flops = fl; g.M = M_; g.N = N_; g.K = K_; g.batch = b_;
      cudaEventCreate(&g.e0); cudaEventCreate(&g.e1);
      cudaEventRecord(g.e0, c->stream);
      idx = c->gemm_prof.size();
      c->gemm_prof.push_back(g);
    }
    ~ProfScope() { if (on) cudaEventRecord(c->gemm_prof[idx].e1, c->stream); }
  } prof_scope(ctx, flops_issued, M, N, K, batch);

  if (impl == GEMM_NAIVE) {
    NSB_REQUIRE(batch <= 65535 && (N + 15) / 16 <= 65535, NSB_EINVAL, "gemm naive: grid too large");
    dim3 grid((unsigned)((M + 15) / 16), (unsigned)((N + 15) / 16), (unsigned)batch), block(16, 16);
    gemm_naive_kernel<T><<<grid, block, 0, ctx->stream>>>(p);
    NSB_CUDA(cudaGetLastError());
    g_last_impl = "naive";
    return;
  }
  if (impl == GEMM_TMA) {
    bool ok = false;
    if (p.alignedA && p.alignedB && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31)) {
      if (p.a_kmajor && p.b_kmajor) ok = launch_tma<T, true, true>(ctx, p);
      else if (p.a_kmajor && !p.b_kmajor) ok = launch_tma<T, true, false>(ctx, p);
      else if (!p.a_kmajor && p.b_kmajor) ok = launch_tma<T, false, true>(ctx, p);
      else ok = launch_tma<T, false, false>(ctx, p);
    }
    if (ok) { g_last_impl = "dmma_tma"; return; }
    impl = GEMM_DMMA;  // operands not TMA-addressable (unaligned strides): cp.async path
  }
  if (p.a_kmajor && p.b_kmajor) launch_dmma<T, true, true>(ctx, p);
  else if (p.a_kmajor && !p.b_kmajor) launch_dmma<T, true, false>(ctx, p);
  else if (!p.a_kmajor && p.b_kmajor) launch_dmma<T, false, true>(ctx, p);
  else launch_dmma<T, false, false>(ctx, p);
  g_last_impl = "dmma_cpasync";
}

template <typename T, bool AK, bool BKM>
static void launch_grouped(Ctx* ctx, const GroupedProblem* probs, int nprob, const GroupedSegment* segs, int64_t total_tiles, const T* A,
                           const T* B, T* C, const GemmParams<T>& proto) {
  typedef TileCfg<T> Cfg;
  size_t smem = (size_t)GEMM_STAGES * (Cfg::BM + Cfg::BN) * 128;
  auto kern = gemm_grouped_kernel<T, AK, BKM>;
  static bool configured_dev[64] = {false};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int grid = (int)std::min<int64_t>(total_tiles, ctx->num_sms);
  kern<<<grid, GEMM_THREADS, smem, ctx->stream>>>(probs, nprob, segs, total_tiles, A, B, C, proto);
  NSB_CUDA(cudaGetLastError());
}

template <typename T>
void gemm_grouped(Ctx* ctx, int opa, int opb, const GroupedProblem* probs_dev, int nprob, const GroupedSegment* segs_dev, int64_t total_tiles,
                  const T* Abase, const T* Bbase, T* Cbase, double flops) {
  if (nprob <= 0 || total_tiles <= 0) return;
  GemmParams<T> p{};
  p.alpha = from_complex<T>(1.0, 0.0); p.beta = zero_<T>();
  p.a_kmajor = (opa == OP_T || opa == OP_C);
  p.b_kmajor = (opb == OP_N || opb == OP_CONJ);
  p.sa = (ScalarTraits<T>::is_complex && (opa == OP_C || opa == OP_CONJ)) ? -1.0 : 1.0;
  p.sb = (ScalarTraits<T>::is_complex && (opb == OP_C || opb == OP_CONJ)) ? -1.0 : 1.0;
  p.lower_only = 0;
  ctx->cnt.gemm_calls++;
  ctx->cnt.kernel_launches++;
  ctx->cnt.gemm_flops += flops;
  const bool prof = ctx->gemm_profile && ctx->gemm_prof.size() < 65536;
  size_t pidx = 0;
  if (prof) {
    Ctx::GemmProf g; g.flops = flops; g.M = -1; g.N = nprob; g.K = total_tiles; g.batch = 1;     // M = -1 marks a grouped launch
    cudaEventCreate(&g.e0); cudaEventCreate(&g.e1);
    cudaEventRecord(g.e0, ctx->stream);
    pidx = ctx->gemm_prof.size();
    ctx->gemm_prof.push_back(g);
  }
  struct Closer { Ctx* c; size_t i; bool on; ~Closer() { if (on) cudaEventRecord(c->gemm_prof[i].e1, c->stream); } } closer{ctx, pidx, prof};
  if (p.a_kmajor && p.b_kmajor) launch_grouped<T, true, true>(ctx, probs_dev, nprob, segs_dev, total_tiles, Abase, Bbase, Cbase, p);
  else if (p.a_kmajor && !p.b_kmajor) launch_grouped<T, true, false>(ctx, probs_dev, nprob, segs_dev, total_tiles, Abase, Bbase, Cbase, p);
  else if (!p.a_kmajor && p.b_kmajor) launch_grouped<T, false, true>(ctx, probs_dev, nprob, segs_dev, total_tiles, Abase, Bbase, Cbase, p);
  else launch_grouped<T, false, false>(ctx, probs_dev, nprob, segs_dev, total_tiles, Abase, Bbase, Cbase, p);
  g_last_impl = "dmma_grouped";
}
int gemm_tile_m(bool cplx) { return cplx ? TileCfg<cdouble>::BM : TileCfg<double>::BM; }
int gemm_tile_n(bool cplx) { return cplx ? TileCfg<cdouble>::BN : TileCfg<double>::BN; }

// ------------------------------------------------------------------------------------------------
// FP64 tensor-pipe ceiling probe: register-resident DMMA issue loop (no memory traffic)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) dmma_peak_kernel(double* out, int iters) {
  double a[4], b[2], c[8][4];
  for (int i = 0; i < 4; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < 2; ++i) b[i] = 1e-9 * (threadIdx.x + i);
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) mma_16x8x8(c[j], a[0], a[1], a[2], a[3], b[0], b[1]);
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  if (s == 123.456) out[0] = s;
}

double dmma_peak_tflops(Ctx* ctx) {
  const int iters = 8192, warps = 16;
  cudaEvent_t e0, e1;
  NSB_CUDA(cudaEventCreate(&e0)); NSB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    NSB_CUDA(cudaEventRecord(e0, ctx->stream));
    dmma_peak_kernel<<<ctx->num_sms, warps * 32, 0, ctx->stream>>>(ctx->d_scratch, iters);
    NSB_CUDA(cudaEventRecord(e1, ctx->stream));
    NSB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    NSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * ctx->num_sms;
    if (rep > 0) best = std::max(best, fl / ms * 1e-9);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return best;
}

template void gemm_grouped<double>(Ctx*, int, int, const GroupedProblem*, int, const GroupedSegment*, int64_t, const double*, const double*, double*, double);
template void gemm_grouped<cdouble>(Ctx*, int, int, const GroupedProblem*, int, const GroupedSegment*, int64_t, const cdouble*, const cdouble*, cdouble*, double);
template void gemm<double>(Ctx*, int, int, int64_t, int64_t, int64_t, double, const double*, int64_t, int64_t,
                           const double*, int64_t, int64_t, double, double*, int64_t, int64_t, int64_t, int, const PeerOut*, int);
template void gemm<cdouble>(Ctx*, int, int, int64_t, int64_t, int64_t, cdouble, const cdouble*, int64_t, int64_t,
                            const cdouble*, int64_t, int64_t, cdouble, cdouble*, int64_t, int64_t, int64_t, int, const PeerOut*, int);

}  // namespace nsb
