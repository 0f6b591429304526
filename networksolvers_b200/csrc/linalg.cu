// Householder QR and one-sided Jacobi SVD on device (see linalg.h).
#include "linalg.h"
#include "ops.h"
#include "gemm.h"
#include "eigh.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <numeric>

namespace nsb {

#define LAUNCH_CHECK(ctx) do { (ctx)->cnt.kernel_launches++; NSB_CUDA(cudaGetLastError()); } while (0)

int64_t truncate_spectrum(const std::vector<double>& Pin, double cutoff, int64_t mindim, int64_t maxdim, double* truncerr) {
  std::vector<double> P(Pin);
  int64_t origm = (int64_t)P.size();
  if (truncerr) *truncerr = 0.0;
  if (origm == 0) return 0;
  for (int64_t n = origm - 1; n >= 0; --n) {
    if (P[n] >= 0.0) break;
    P[n] = 0.0;
  }
  if (origm == 1) return 1;
  int64_t n = origm;
  double terr = 0.0;
  while (n > maxdim) { terr += P[n - 1]; --n; }
  double scale = 0.0;
  for (double x : P) scale += x;
  if (scale == 0.0) scale = 1.0;
  while (n > mindim && (terr + P[n - 1] <= cutoff * scale)) { terr += P[n - 1]; --n; }
  terr /= scale;
  if (n < 1) n = 1;
  if (truncerr) *truncerr = terr;
  return n;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV doubles per thread; result valid in all threads
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sh /* >= NV*8 + NV */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum_d(v[i]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[i * 8 + w] = v[i];
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = 0.0;
      for (int j = 0; j < nw; ++j) s += sh[i * 8 + j];
      sh[NV * 8 + i] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = sh[NV * 8 + i];
}

// ------------------------------------------------------------------------------------------------
// Householder QR
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) house_gen_kernel(T* __restrict__ A, int64_t rows, int64_t lda, int64_t j, T* __restrict__ tau) {
  __shared__ double sh[32];
  T* col = A + j * lda;
  double s[1] = {0.0};
  for (int64_t i = j + 1 + threadIdx.x; i < rows; i += blockDim.x) s[0] += abs2_(col[i]);
  block_sum<1>(s, sh);
  T alpha = col[j];
  double sigma = s[0];
  if (sigma == 0.0 && im(alpha) == 0.0) {
    if (threadIdx.x == 0) tau[j] = zero_<T>();
    return;
  }
  double ar = re(alpha), ai = im(alpha);
  double beta = -copysign(sqrt(ar * ar + ai * ai + sigma), ar);
  T t = from_complex<T>((beta - ar) / beta, -ai / beta);
  // scale = 1 / (alpha - beta)
  double dr = ar - beta, di = ai, den = dr * dr + di * di;
  T scale = from_complex<T>(dr / den, -di / den);
  for (int64_t i = j + 1 + threadIdx.x; i < rows; i += blockDim.x) col[i] = mul_(scale, col[i]);
  __syncthreads();
  if (threadIdx.x == 0) { tau[j] = t; col[j] = from_complex<T>(beta, 0.0); }
}

// Apply H_j = I - tau v v^H (or its adjoint) to columns [c0, c1) of B (rows x *, ldb); v from column j of Afac.
template <typename T>
__global__ void __launch_bounds__(256) house_apply_kernel(const T* __restrict__ Afac, int64_t lda, int64_t rows, int64_t j,
                                                          const T* __restrict__ tau, int adjoint, T* __restrict__ B,
                                                          int64_t ldb, int64_t c0, int64_t c1) {
  __shared__ double sh[32];
  T tj = tau[j];
  if (re(tj) == 0.0 && im(tj) == 0.0) return;
  if (adjoint) tj = conj_(tj);
  const T* v = Afac + j * lda;
  for (int64_t c = c0 + blockIdx.x; c < c1; c += gridDim.x) {
    T* col = B + c * ldb;
    double s[2] = {0.0, 0.0};
    for (int64_t i = j + threadIdx.x; i < rows; i += blockDim.x) {
      T vi = (i == j) ? from_complex<T>(1.0, 0.0) : v[i];
      T x = col[i];
      s[0] += re(vi) * re(x) + im(vi) * im(x);
      s[1] += re(vi) * im(x) - im(vi) * re(x);
    }
    block_sum<2>(s, sh);
    T w = mul_(tj, from_complex<T>(s[0], s[1]));
    T mw = from_complex<T>(-re(w), -im(w));
    for (int64_t i = j + threadIdx.x; i < rows; i += blockDim.x) {
      T vi = (i == j) ? from_complex<T>(1.0, 0.0) : v[i];
      T x = col[i];
      fma_(x, mw, vi);
      col[i] = x;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void extract_r_kernel(const T* __restrict__ A, int64_t lda, int64_t k, int64_t cols, T* __restrict__ R, int64_t ldr) {
  int64_t total = k * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i % k, c = i / k;
    R[r + c * ldr] = (r <= c) ? A[r + c * lda] : zero_<T>();
  }
}

// ------------------------------------------------------------------------------------------------
// Blocked Householder QR in compact-WY form (K8: gauge moves, 1-site TDVP split, tall factorisations).
//   panel (QR_NB columns): ONE launch per column -- CTA b of the launch re-derives the reflector scalars from the raw
//     column j (kept unmodified in A: the strictly lower part of a factored column is dead storage, R's diagonal goes to
//     `rdiag`), then
//       b <  i : Gram entry S[b, i] = v_b^H v_i for the T factor,
//       b == i : writes v_i (explicit unit diagonal, zeros above) into V and tau_i, rdiag_j,
//       b >  i : applies H_i^H to panel column p + b;
//   T factor from S by one CTA (larft recurrence); trailing update A2 <- (I - V T^H V^H) A2 and the formation of
//   Q = H_1 .. H_k [I; 0] (blocks descending) as DMMA GEMMs, with Y = V^H C as a split-K batch whose slabs are summed by the
//   T-factor GEMM ([T T .. T] x stack), exactly as the back-transformation of eigh.cu.
// ------------------------------------------------------------------------------------------------
constexpr int QR_NB = 64;

// block-wide sum for up to 32 warps (block_sum above assumes <= 8); result valid in all threads
template <int NV>
__device__ __forceinline__ void block_sum32(double (&v)[NV], double* sh /* >= NV * 33 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum_d(v[i]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[i * 33 + w] = v[i];
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = lane < nw ? sh[i * 33 + lane] : 0.0;
      x = warp_sum_d(x);
      if (lane == 0) sh[i * 33 + 32] = x;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = sh[i * 33 + 32];
}

// ITEMS: rows per thread held in registers (column j and the CTA's own column are read once; rows beyond ITEMS * blockDim are
// streamed a second time).
template <typename T, int ITEMS>
__global__ void __launch_bounds__(512) qr_panel_col_kernel(T* __restrict__ A, int64_t lda, int64_t rows, int64_t cols, int64_t p, int i,
                                                           int w, T* __restrict__ V, int64_t ldv, T* __restrict__ tau,
                                                           T* __restrict__ rdiag, T* __restrict__ S /* QR_NB x QR_NB */) {
  __shared__ double sh[3 * 33];
  const int b = blockIdx.x;
  const int64_t j = p + i;
  const T* cj = A + j * lda;
  const int64_t c = p + b;
  // second operand of this CTA: b < i: the finished reflector v_b; b > i: panel column p + b; b == i: none
  const T* other = (b < i) ? (V + (p + b) * ldv) : ((b > i && c < cols) ? (A + c * lda) : nullptr);
  const int64_t r0 = j + 1 + threadIdx.x, stride = blockDim.x;
  T a[ITEMS], x[ITEMS];
  double s[3] = {0.0, 0.0, 0.0};    // sum |a|^2, sum conj(a) x (re, im)
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int64_t r = r0 + (int64_t)k * stride;
    a[k] = (r < rows) ? cj[r] : zero_<T>();
    x[k] = (r < rows && other) ? other[r] : zero_<T>();
  }
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    s[0] += abs2_(a[k]);
    s[1] += re(a[k]) * re(x[k]) + im(a[k]) * im(x[k]);
    s[2] += re(a[k]) * im(x[k]) - im(a[k]) * re(x[k]);
  }
  for (int64_t r = r0 + (int64_t)ITEMS * stride; r < rows; r += stride) {   // tall matrices: streamed remainder
    const T ar_ = cj[r];
    const T xr_ = other ? other[r] : zero_<T>();
    s[0] += abs2_(ar_);
    s[1] += re(ar_) * re(xr_) + im(ar_) * im(xr_);
    s[2] += re(ar_) * im(xr_) - im(ar_) * re(xr_);
  }
  // row j of this CTA's own column is read by every thread BEFORE the barriers of the block sum and rewritten by thread 0
  // at the very end: reading it after the barriers would race with that store (a warp running ahead of a stalled one)
  const T xj = (b > i && other) ? other[j] : zero_<T>();
  block_sum32<3>(s, sh);
  const T alpha = cj[j];
  const double sigma = s[0], ar = re(alpha), ai = im(alpha);
  T t = zero_<T>(), scale = zero_<T>();
  double beta = ar;
  const bool trivial = (sigma == 0.0 && ai == 0.0);
  if (!trivial) {
    beta = -copysign(sqrt(ar * ar + ai * ai + sigma), ar);
    t = from_complex<T>((beta - ar) / beta, -ai / beta);
    const double dr = ar - beta, di = ai, den = dr * dr + di * di;
    scale = from_complex<T>(dr / den, -di / den);
  }
  if (b == i) {
    T* vj = V + j * ldv;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int64_t r = r0 + (int64_t)k * stride;
      if (r < rows) vj[r] = mul_(scale, a[k]);
    }
    for (int64_t r = r0 + (int64_t)ITEMS * stride; r < rows; r += stride) vj[r] = mul_(scale, cj[r]);
    if (threadIdx.x == 0) {
      vj[j] = from_complex<T>(1.0, 0.0);
      tau[j] = t;
      rdiag[j] = trivial ? alpha : from_complex<T>(beta, 0.0);
      S[i + i * QR_NB] = from_complex<T>(1.0, 0.0);
    }
    return;
  }
  // sum conj(v) x over the rows below j with v = scale * a:  conj(scale) * (s[1] + i s[2])
  const T cs = mul_(conj_(scale), from_complex<T>(s[1], s[2]));
  if (b < i) {   // S[b, i] = v_b^H v_i = conj(v_b[j]) + sum_r conj(v_b[r]) v_i[r] = conj(v_b[j]) + conj(sum_r conj(v_i[r]) v_b[r])
    if (threadIdx.x == 0) {
      const T y = other[j];
      S[b + i * QR_NB] = add_(conj_(y), conj_(cs));
    }
    return;
  }
  if (!other || trivial) return;
  T* cc = A + c * lda;
  const T dot = add_(xj, cs);                                            // v^H x  (v_j = 1)
  const T f = mul_(conj_(t), dot);                                       // H^H x = x - conj(tau) (v^H x) v
  const T mf = from_complex<T>(-re(f), -im(f));
  const T mfs = mul_(mf, scale);
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int64_t r = r0 + (int64_t)k * stride;
    if (r < rows) { T xv = x[k]; fma_(xv, mfs, a[k]); cc[r] = xv; }
  }
  for (int64_t r = r0 + (int64_t)ITEMS * stride; r < rows; r += stride) { T xv = cc[r]; fma_(xv, mfs, cj[r]); cc[r] = xv; }
  if (threadIdx.x == 0) cc[j] = add_(xj, mf);
}

template <typename T> struct QrItems;
template <> struct QrItems<double> { static constexpr int N = 16; };
template <> struct QrItems<cdouble> { static constexpr int N = 8; };

// T factor of one panel from its Gram matrix (upper triangle of S) and tau; written `reps` times side by side as T (Tst) and
// as T^H (THst), both with leading dimension QR_NB.
template <typename T>
__global__ void __launch_bounds__(256) qr_larft_kernel(const T* __restrict__ S, const T* __restrict__ tau, int w, T* __restrict__ Tst,
                                                       T* __restrict__ THst, int reps) {
  extern __shared__ __align__(16) char qr_larft_sm[];
  T* Ts = reinterpret_cast<T*>(qr_larft_sm);
  T* sc = Ts + QR_NB * QR_NB;
  for (int e = threadIdx.x; e < w * w; e += blockDim.x) Ts[e] = zero_<T>();
  __syncthreads();
  for (int i = 0; i < w; ++i) {
    for (int k = threadIdx.x; k < i; k += blockDim.x) sc[k] = S[k + (size_t)i * QR_NB];
    __syncthreads();
    const T ti = tau[i];
    const T mti = from_complex<T>(-re(ti), -im(ti));
    for (int r = threadIdx.x; r <= i; r += blockDim.x) {
      if (r < i) {
        T acc = zero_<T>();
        for (int k = r; k < i; ++k) fma_(acc, Ts[r + k * w], sc[k]);
        Ts[r + i * w] = mul_(mti, acc);
      } else {
        Ts[i + i * w] = ti;
      }
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < w * w * reps; e += blockDim.x) {
    const int sidx = e / (w * w), rc = e - sidx * (w * w), r = rc % w, c = rc / w;
    Tst[r + ((size_t)sidx * w + c) * QR_NB] = Ts[rc];
    THst[c + ((size_t)sidx * w + r) * QR_NB] = conj_(Ts[rc]);
  }
}

template <typename T>
__global__ void extract_r_diag_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ rdiag, int64_t k, int64_t cols,
                                      T* __restrict__ R, int64_t ldr) {
  const int64_t total = k * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i % k, c = i / k;
    R[r + c * ldr] = (r < c) ? A[r + c * lda] : (r == c ? rdiag[r] : zero_<T>());
  }
}

// C (mp x nc, ldc) <- (I - V Tm V^H) C with Tm given as `split` side-by-side copies (T for Q, T^H for Q^H); V: mp x w (ldv)
template <typename T>
static void apply_block_reflector(Ctx* ctx, const T* V, int64_t ldv, int64_t mp, int w, const T* Tstack, int split, T* Cm, int64_t ldc,
                                  int64_t nc, T* Y, T* Y2) {
  if (nc <= 0 || mp <= 0) return;
  const T one = from_complex<T>(1.0, 0.0), mone = from_complex<T>(-1.0, 0.0), zero = zero_<T>();
  int64_t sb = std::min<int64_t>(split, std::max<int64_t>(1, mp / 256));   // slabs of at least 256 rows
  int64_t kc = (((mp + sb - 1) / sb) + 15) / 16 * 16;
  const int64_t nf = mp / kc, rem = mp - nf * kc, st = nf + (rem > 0 ? 1 : 0);
  const int64_t ldy = st * w;
  if (nf > 0) gemm<T>(ctx, OP_C, OP_N, w, nc, kc, one, V, ldv, kc, Cm, ldc, kc, zero, Y, ldy, w, nf);
  if (rem > 0) gemm<T>(ctx, OP_C, OP_N, w, nc, rem, one, V + nf * kc, ldv, 0, Cm + nf * kc, ldc, 0, zero, Y + nf * w, ldy, 0, 1);
  gemm<T>(ctx, OP_N, OP_N, w, nc, st * w, one, Tstack, QR_NB, 0, Y, ldy, 0, zero, Y2, w, 0, 1);
  gemm<T>(ctx, OP_N, OP_N, mp, nc, w, mone, V, ldv, 0, Y2, w, 0, one, Cm, ldc, 0, 1);
}

template <typename T>
static void qr_thin_blocked(Ctx* ctx, T* A, int64_t rows, int64_t cols, int64_t lda, T* Q, int64_t ldq, T* R, int64_t ldr) {
  const int64_t k = std::min(rows, cols);
  const int64_t npanels = (k + QR_NB - 1) / QR_NB;
  const int split = 8;
  DevBuf Vb(ctx, sizeof(T) * (size_t)rows * k), taub(ctx, sizeof(T) * k), rdb(ctx, sizeof(T) * k);
  DevBuf Sb(ctx, sizeof(T) * QR_NB * QR_NB), Tall(ctx, sizeof(T) * (size_t)QR_NB * QR_NB * split * npanels),
      THb(ctx, sizeof(T) * (size_t)QR_NB * QR_NB * split);
  const int64_t ncmax = std::max(cols, k);
  DevBuf Y(ctx, sizeof(T) * (size_t)QR_NB * split * ncmax), Y2(ctx, sizeof(T) * (size_t)QR_NB * ncmax);
  T* V = (T*)Vb.ptr;
  T* tau = (T*)taub.ptr;
  NSB_CUDA(cudaMemsetAsync(V, 0, sizeof(T) * (size_t)rows * k, ctx->stream));
  const int threads = rows >= 2048 ? 512 : 256;
  const size_t larft_smem = sizeof(T) * (QR_NB * QR_NB + QR_NB);
  {
    static bool configured[2][64] = {{false}};
    bool& c = configured[ScalarTraits<T>::is_complex ? 1 : 0][ctx->device & 63];
    if (!c) { NSB_CUDA(cudaFuncSetAttribute(qr_larft_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)larft_smem)); c = true; }
  }
  for (int64_t pi = 0; pi < npanels; ++pi) {
    const int64_t p = pi * QR_NB;
    const int w = (int)std::min<int64_t>(QR_NB, k - p);
    NSB_CUDA(cudaMemsetAsync(Sb.ptr, 0, sizeof(T) * QR_NB * QR_NB, ctx->stream));
    for (int i = 0; i < w; ++i) {
      qr_panel_col_kernel<T, QrItems<T>::N><<<w, threads, 0, ctx->stream>>>(A, lda, rows, cols, p, i, w, V, rows, tau, (T*)rdb.ptr, (T*)Sb.ptr);
      LAUNCH_CHECK(ctx);
    }
    T* Tp = (T*)Tall.ptr + (size_t)pi * QR_NB * QR_NB * split;
    qr_larft_kernel<T><<<1, 256, larft_smem, ctx->stream>>>((const T*)Sb.ptr, tau + p, w, Tp, (T*)THb.ptr, split);
    LAUNCH_CHECK(ctx);
    const int64_t nc = cols - (p + w);
    if (nc > 0)   // A[p:, p+w:] <- Q_p^H A[p:, p+w:]
      apply_block_reflector<T>(ctx, V + p + p * rows, rows, rows - p, w, (const T*)THb.ptr, split, A + p + (p + w) * lda, lda, nc,
                               (T*)Y.ptr, (T*)Y2.ptr);
  }
  {
    const int64_t total = k * cols;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 8);
    extract_r_diag_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, lda, (const T*)rdb.ptr, k, cols, R, ldr);
    LAUNCH_CHECK(ctx);
  }
  set_identity<T>(ctx, Q, rows, k, ldq);
  for (int64_t pi = npanels - 1; pi >= 0; --pi) {
    const int64_t p = pi * QR_NB;
    const int w = (int)std::min<int64_t>(QR_NB, k - p);
    const T* Tp = (const T*)Tall.ptr + (size_t)pi * QR_NB * QR_NB * split;
    apply_block_reflector<T>(ctx, V + p + p * rows, rows, rows - p, w, Tp, split, Q + p + p * ldq, ldq, k - p, (T*)Y.ptr, (T*)Y2.ptr);
  }
}

// ------------------------------------------------------------------------------------------------
// Small thin QR as ONE launch (the latency-bound regime: gauge moves and tall factorisations at chi <~ 100, where the
// column-at-a-time kernels above cost three launches per column).  One CTA, A and Q resident in shared memory; per column one
// block reduction (norm), then one warp per trailing column (reflector application: shuffle-reduced dot + axpy); Q is formed
// by applying the reflectors to [I; 0] in reverse order.  Same reflector conventions as house_gen_kernel / house_apply_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int QS_THREADS = 512;
constexpr size_t QS_SMEM_MAX = 224 * 1024;

template <typename T>
__device__ __forceinline__ T warp_sum_T(T v);
template <> __device__ __forceinline__ double warp_sum_T<double>(double v) { return warp_sum_d(v); }
template <> __device__ __forceinline__ cdouble warp_sum_T<cdouble>(cdouble v) {
  return make_cuDoubleComplex(warp_sum_d(v.x), warp_sum_d(v.y));
}

// x <- (I - t v v^H) x for the reflector stored in column `vcol` below row j (v_j = 1), executed by one warp
template <typename T>
__device__ __forceinline__ void qs_apply(const T* vcol, T* x, int j, int rows, T t, int lane) {
  T dot = zero_<T>();
  for (int i = j + 1 + lane; i < rows; i += 32) fma_(dot, conj_(vcol[i]), x[i]);
  dot = warp_sum_T<T>(dot);
  dot = add_(dot, x[j]);                 // (every lane has read x[j] before lane 0 rewrites it: the shuffles above order them)
  const T w = mul_(t, dot);
  const T mw = from_complex<T>(-re(w), -im(w));
  __syncwarp();
  if (lane == 0) x[j] = add_(x[j], mw);
  for (int i = j + 1 + lane; i < rows; i += 32) { T xi = x[i]; fma_(xi, mw, vcol[i]); x[i] = xi; }
}

template <typename T>
__global__ void __launch_bounds__(QS_THREADS) qr_smem_kernel(const T* __restrict__ A, int64_t lda, int rows, int cols,
                                                             T* __restrict__ Q, int64_t ldq, T* __restrict__ R, int64_t ldr) {
  extern __shared__ __align__(16) unsigned char qs_raw[];
  __shared__ double red[33];
  const int k = min(rows, cols);
  T* As = reinterpret_cast<T*>(qs_raw);          // rows x cols
  T* Qs = As + (size_t)rows * cols;              // rows x k
  T* taus = Qs + (size_t)rows * k;               // k
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int e = tid; e < rows * cols; e += nt) As[e] = A[(e % rows) + (int64_t)(e / rows) * lda];
  for (int e = tid; e < rows * k; e += nt) Qs[e] = ((e % rows) == (e / rows)) ? from_complex<T>(1.0, 0.0) : zero_<T>();
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    T* col = As + (size_t)j * rows;
    double s[1] = {0.0};
    for (int i = j + 1 + tid; i < rows; i += nt) s[0] += abs2_(col[i]);
    block_sum32<1>(s, red);
    const T alpha = col[j];
    const double sigma = s[0], ar = re(alpha), ai = im(alpha);
    T t = zero_<T>();
    if (!(sigma == 0.0 && ai == 0.0)) {
      const double beta = -copysign(sqrt(ar * ar + ai * ai + sigma), ar);
      t = from_complex<T>((beta - ar) / beta, -ai / beta);
      const double dr = ar - beta, di = ai, den = dr * dr + di * di;
      const T scale = from_complex<T>(dr / den, -di / den);
      __syncthreads();                   // every thread holds alpha before column j is overwritten
      for (int i = j + 1 + tid; i < rows; i += nt) col[i] = mul_(scale, col[i]);
      if (tid == 0) col[j] = from_complex<T>(beta, 0.0);
    }
    if (tid == 0) taus[j] = t;
    __syncthreads();
    if (re(t) != 0.0 || im(t) != 0.0) {
      const T tc = conj_(t);             // factorisation applies H^H
      for (int c = j + 1 + warp; c < cols; c += nw) qs_apply<T>(col, As + (size_t)c * rows, j, rows, tc, lane);
    }
    __syncthreads();
  }
  for (int e = tid; e < k * cols; e += nt) {
    const int r = e % k, c = e / k;
    R[r + (int64_t)c * ldr] = (r <= c) ? As[r + (size_t)c * rows] : zero_<T>();
  }
  for (int j = k - 1; j >= 0; --j) {     // Q = H_0 ... H_{k-1} [I; 0]
    const T t = taus[j];
    if (re(t) != 0.0 || im(t) != 0.0)
      for (int c = j + warp; c < k; c += nw) qs_apply<T>(As + (size_t)j * rows, Qs + (size_t)c * rows, j, rows, t, lane);
    __syncthreads();
  }
  for (int e = tid; e < rows * k; e += nt) Q[(e % rows) + (int64_t)(e / rows) * ldq] = Qs[e];
}

template <typename T>
static bool qr_thin_smem(Ctx* ctx, const T* A, int64_t rows, int64_t cols, int64_t lda, T* Q, int64_t ldq, T* R, int64_t ldr) {
  const int64_t k = std::min(rows, cols);
  const size_t smem = sizeof(T) * ((size_t)rows * cols + (size_t)rows * k + (size_t)k);
  if (smem > QS_SMEM_MAX) return false;
  static bool configured[2][64] = {{false}};
  bool& c = configured[ScalarTraits<T>::is_complex ? 1 : 0][ctx->device & 63];
  if (!c) { NSB_CUDA(cudaFuncSetAttribute(qr_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QS_SMEM_MAX)); c = true; }
  int threads = (int)std::min<int64_t>(QS_THREADS, std::max<int64_t>(64, 32 * std::max<int64_t>(cols, (rows + 31) / 32)));
  threads = (threads + 31) / 32 * 32;
  qr_smem_kernel<T><<<1, threads, smem, ctx->stream>>>(A, lda, (int)rows, (int)cols, Q, ldq, R, ldr);
  LAUNCH_CHECK(ctx);
  return true;
}

template <typename T>
void qr_thin(Ctx* ctx, T* A, int64_t rows, int64_t cols, int64_t lda, T* Q, int64_t ldq, T* R, int64_t ldr) {
  int64_t k = std::min(rows, cols);
  if (k == 0) return;
  ctx->cnt.qr_calls++;
  if (ctx->opt.qr_smem && qr_thin_smem<T>(ctx, A, rows, cols, lda, Q, ldq, R, ldr)) return;
  if (ctx->opt.qr_block_min > 0 && k >= ctx->opt.qr_block_min && rows >= 2 * QR_NB) {
    qr_thin_blocked<T>(ctx, A, rows, cols, lda, Q, ldq, R, ldr);
    return;
  }
  DevBuf tau(ctx, sizeof(T) * k);
  T* dtau = (T*)tau.ptr;
  int maxgrid = ctx->num_sms * 4;
  for (int64_t j = 0; j < k; ++j) {
    house_gen_kernel<T><<<1, 256, 0, ctx->stream>>>(A, rows, lda, j, dtau);
    LAUNCH_CHECK(ctx);
    if (j + 1 < cols) {
      int grid = (int)std::min<int64_t>(cols - j - 1, maxgrid);
      house_apply_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, lda, rows, j, dtau, 1, A, lda, j + 1, cols);
      LAUNCH_CHECK(ctx);
    }
  }
  {
    int64_t total = k * cols;
    int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 8);
    extract_r_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, lda, k, cols, R, ldr);
    LAUNCH_CHECK(ctx);
  }
  set_identity<T>(ctx, Q, rows, k, ldq);
  for (int64_t j = k - 1; j >= 0; --j) {
    int grid = (int)std::min<int64_t>(k - j, maxgrid);
    house_apply_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, lda, rows, j, dtau, 0, Q, ldq, j, k);
    LAUNCH_CHECK(ctx);
  }
}

// Column-pivoted Householder QR (rank revealing; used as the preconditioner of the blocked Jacobi SVD):
//   A P = Q R,  |R_jj| non-increasing.  perm_out[j] = original column placed at position j.
__global__ void __launch_bounds__(256) pivot_select_kernel(double* __restrict__ cn, int64_t j, int64_t cols, int32_t* __restrict__ piv) {
  __shared__ double sv[256];
  __shared__ int si[256];
  double best = -1.0; int bi = (int)j;
  for (int64_t c = j + threadIdx.x; c < cols; c += blockDim.x) if (cn[c] > best) { best = cn[c]; bi = (int)c; }
  sv[threadIdx.x] = best; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      if (sv[threadIdx.x + o] > sv[threadIdx.x] || (sv[threadIdx.x + o] == sv[threadIdx.x] && si[threadIdx.x + o] < si[threadIdx.x])) {
        sv[threadIdx.x] = sv[threadIdx.x + o]; si[threadIdx.x] = si[threadIdx.x + o];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int p = si[0];
    piv[j] = p;
    double t = cn[j]; cn[j] = cn[p]; cn[p] = t;
  }
}
template <typename T>
__global__ void swap_cols_kernel(T* __restrict__ A, int64_t lda, int64_t rows, int64_t j, const int32_t* __restrict__ piv) {
  int64_t p = piv[j];
  if (p == j) return;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    T a = A[r + j * lda], b = A[r + p * lda];
    A[r + j * lda] = b; A[r + p * lda] = a;
  }
}
template <typename T>
__global__ void norm_downdate_kernel(const T* __restrict__ A, int64_t lda, int64_t j, int64_t cols, double* __restrict__ cn) {
  for (int64_t c = j + 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += (int64_t)gridDim.x * blockDim.x) {
    double v = cn[c] - abs2_(A[j + c * lda]);
    cn[c] = v > 0.0 ? v : 0.0;
  }
}

template <typename T>
void qr_pivoted_thin(Ctx* ctx, T* A, int64_t rows, int64_t cols, int64_t lda, T* Q, int64_t ldq, T* R, int64_t ldr,
                     std::vector<int32_t>& perm_out) {
  int64_t k = std::min(rows, cols);
  ctx->cnt.qr_calls++;
  DevBuf tau(ctx, sizeof(T) * k), cn(ctx, sizeof(double) * cols), piv(ctx, sizeof(int32_t) * k);
  T* dtau = (T*)tau.ptr;
  col_norms2<T>(ctx, A, rows, cols, lda, (double*)cn.ptr);
  int maxgrid = ctx->num_sms * 4;
  for (int64_t j = 0; j < k; ++j) {
    if (j > 0 && j % 64 == 0 && j < cols)   // refresh the trailing norms (the downdate loses accuracy by cancellation)
      col_norms2<T>(ctx, A + j + j * lda, rows - j, cols - j, lda, (double*)cn.ptr + j);
    pivot_select_kernel<<<1, 256, 0, ctx->stream>>>((double*)cn.ptr, j, cols, (int32_t*)piv.ptr);
    LAUNCH_CHECK(ctx);
    swap_cols_kernel<T><<<std::max(1, (int)std::min<int64_t>((rows + 255) / 256, maxgrid)), 256, 0, ctx->stream>>>(A, lda, rows, j, (const int32_t*)piv.ptr);
    LAUNCH_CHECK(ctx);
    house_gen_kernel<T><<<1, 256, 0, ctx->stream>>>(A, rows, lda, j, dtau);
    LAUNCH_CHECK(ctx);
    if (j + 1 < cols) {
      int grid = (int)std::min<int64_t>(cols - j - 1, maxgrid);
      house_apply_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, lda, rows, j, dtau, 1, A, lda, j + 1, cols);
      LAUNCH_CHECK(ctx);
      norm_downdate_kernel<T><<<std::max(1, (int)std::min<int64_t>((cols - j + 255) / 256, maxgrid)), 256, 0, ctx->stream>>>(A, lda, j, cols, (double*)cn.ptr);
      LAUNCH_CHECK(ctx);
    }
  }
  {
    int64_t total = k * cols;
    int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 8);
    extract_r_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, lda, k, cols, R, ldr);
    LAUNCH_CHECK(ctx);
  }
  set_identity<T>(ctx, Q, rows, k, ldq);
  for (int64_t j = k - 1; j >= 0; --j) {
    int grid = (int)std::min<int64_t>(k - j, maxgrid);
    house_apply_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, lda, rows, j, dtau, 0, Q, ldq, j, k);
    LAUNCH_CHECK(ctx);
  }
  std::vector<int32_t> hp(k);
  NSB_CUDA(cudaMemcpyAsync(hp.data(), piv.ptr, sizeof(int32_t) * k, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  perm_out.resize(cols);
  std::iota(perm_out.begin(), perm_out.end(), 0);
  for (int64_t j = 0; j < k; ++j) std::swap(perm_out[j], perm_out[hp[j]]);
}

// ------------------------------------------------------------------------------------------------
// one-sided Jacobi: orthogonalise the n columns of G (m x n), accumulating the rotations in V (nv x n)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) jacobi_round_kernel(T* __restrict__ G, int64_t ldg, int64_t m, T* __restrict__ V,
                                                           int64_t ldv, int64_t nv, int64_t n, int64_t npad, int64_t round,
                                                           double tol, unsigned long long* __restrict__ maxoff) {
  __shared__ double sh[64];
  const int64_t npairs = npad / 2, ring = npad - 1;
  for (int64_t i = blockIdx.x; i < npairs; i += gridDim.x) {
    int64_t p, q;
    if (i == 0) { p = ring; q = round % ring; }
    else { p = (round + i) % ring; q = (round + ring - i) % ring; }
    if (p > q) { int64_t tmp = p; p = q; q = tmp; }
    if (q >= n) continue;   // padded column (uniform per block)
    T* gp = G + p * ldg;
    T* gq = G + q * ldg;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t r = threadIdx.x; r < m; r += blockDim.x) {
      T a = gp[r], b = gq[r];
      s[0] += abs2_(a);
      s[1] += abs2_(b);
      s[2] += re(a) * re(b) + im(a) * im(b);   // conj(a) * b
      s[3] += re(a) * im(b) - im(a) * re(b);
    }
    block_sum<4>(s, sh);
    double alpha = s[0], beta = s[1], gr = s[2], gi = s[3];
    double gabs = sqrt(gr * gr + gi * gi);
    double denom = sqrt(alpha * beta);
    if (gabs == 0.0 || gabs <= tol * denom) continue;
    if (threadIdx.x == 0) atomicMax(maxoff, (unsigned long long)__double_as_longlong(gabs / denom));
    double zeta = (beta - alpha) / (2.0 * gabs);
    double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    double c = 1.0 / sqrt(1.0 + tt * tt), sn = c * tt;
    // phase = conj(gamma)/|gamma| = e^{-i phi}
    double pr = gr / gabs, pi = -gi / gabs;
    T cq = from_complex<T>(c * pr, c * pi);      // c e^{-i phi}
    T sq = from_complex<T>(-sn * pr, -sn * pi);  // -s e^{-i phi}
    T cc = from_complex<T>(c, 0.0), ss = from_complex<T>(sn, 0.0);
    for (int64_t r = threadIdx.x; r < m; r += blockDim.x) {
      T a = gp[r], b = gq[r];
      T na = mul_(cc, a); fma_(na, sq, b);
      T nb = mul_(ss, a); fma_(nb, cq, b);
      gp[r] = na; gq[r] = nb;
    }
    T* vp = V + p * ldv;
    T* vq = V + q * ldv;
    for (int64_t r = threadIdx.x; r < nv; r += blockDim.x) {
      T a = vp[r], b = vq[r];
      T na = mul_(cc, a); fma_(na, sq, b);
      T nb = mul_(ss, a); fma_(nb, cq, b);
      vp[r] = na; vq[r] = nb;
    }
  }
}

template <typename T>
__global__ void conj_copy_kernel(const T* __restrict__ in, int64_t ldi, T* __restrict__ out, int64_t ldo, int64_t rows, int64_t cols) {
  int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i % rows, c = i / rows;
    out[r + c * ldo] = conj_(in[r + c * ldi]);
  }
}
template <typename T>
static void conj_copy_block(Ctx* ctx, const T* in, int64_t ldi, T* out, int64_t ldo, int64_t rows, int64_t cols) {
  if (rows * cols == 0) return;
  int grid = (int)std::min<int64_t>((rows * cols + 255) / 256, (int64_t)ctx->num_sms * 8);
  conj_copy_kernel<T><<<grid, 256, 0, ctx->stream>>>(in, ldi, out, ldo, rows, cols);
  LAUNCH_CHECK(ctx);
}

// returns number of sweeps
template <typename T>
static int jacobi_onesided(Ctx* ctx, T* G, int64_t ldg, int64_t m, int64_t n, T* V, int64_t ldv, int64_t nv) {
  if (n <= 1) return 0;
  int64_t npad = (n % 2) ? n + 1 : n;
  double tol = std::sqrt((double)std::max<int64_t>(m, 1)) * 2.220446049250313e-16;
  unsigned long long* dmax = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
  int grid = (int)std::min<int64_t>(npad / 2, (int64_t)ctx->num_sms * 8);
  int sweep = 0;
  const int max_sweeps = 40;
  for (; sweep < max_sweeps; ++sweep) {
    NSB_CUDA(cudaMemsetAsync(dmax, 0, sizeof(unsigned long long), ctx->stream));
    for (int64_t r = 0; r < npad - 1; ++r) {
      jacobi_round_kernel<T><<<grid, 256, 0, ctx->stream>>>(G, ldg, m, V, ldv, nv, n, npad, r, tol, dmax);
      LAUNCH_CHECK(ctx);
    }
    NSB_CUDA(cudaMemcpyAsync(ctx->h_pinned, dmax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NSB_CUDA(cudaStreamSynchronize(ctx->stream));
    double off = ctx->h_pinned[0];
    ctx->cnt.jacobi_sweeps++;
    if (off <= tol) { ++sweep; break; }
  }
  return sweep;
}


// ------------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi for small matrices as ONE launch (the latency-bound regime: chi <~ 128, where a truncating
// factorisation used to cost n - 1 launches per sweep and one host synchronisation per sweep).  One warp per column pair of
// the round-robin round; a column pair of up to 32 * JC_ITEMS rows stays in registers between the Gram pass and the rotation.
//   * jacobi_smem_kernel: one CTA of 16 warps, the matrix and the accumulated rotations resident in shared memory (224 KB:
//     up to 118 x 118 real), rounds separated by __syncthreads.  One SM is instruction-issue bound from ~50 pairs per round
//     on (measured 2.3 us per round at n = 64, 4.6 us at n = 100), which is what bounds the useful size.
//   * jacobi_cluster_kernel: beyond that, one thread-block cluster of up to 8 CTAs x 16 warps on the L2-resident matrix,
//     rounds separated by the hardware cluster barrier (barrier.cluster release / acquire) instead of a kernel boundary.
// Convergence (no rotation in a full sweep) is decided on the device; the squared column norms and the sweep count are left
// in `out` ([0, n) norms, [n] sweeps) for the single read-back the truncation rule needs anyway.
// ------------------------------------------------------------------------------------------------
constexpr int JC_WARPS = 16;        // warps per CTA of the cluster kernel
constexpr int JS_THREADS = 512;     // threads of the shared-memory kernel (16 warps: 128 registers each, no spills)
constexpr size_t JS_SMEM_MAX = 224 * 1024;
template <typename T> struct JcItems;
template <> struct JcItems<double> { static constexpr int N = 8; };
template <> struct JcItems<cdouble> { static constexpr int N = 4; };

// One column pair (gp, gq: m rows; vp, vq: nv rows), executed by one warp: Gram entries, rotation, update.  Returns whether
// the pair was rotated (same value in every lane: the xor butterfly leaves identical bits everywhere).
template <typename T, int IT>
__device__ __forceinline__ bool jc_pair_it(T* gp, T* gq, int m, T* vp, T* vq, int nv, double tol2, int lane) {
  const bool in_regs = m <= 32 * IT;
  T a[IT], b[IT];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (in_regs) {
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const int r = lane + 32 * k;
      a[k] = (r < m) ? gp[r] : zero_<T>();
      b[k] = (r < m) ? gq[r] : zero_<T>();
    }
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      s0 += abs2_(a[k]);
      s1 += abs2_(b[k]);
      s2 += re(a[k]) * re(b[k]) + im(a[k]) * im(b[k]);   // conj(a) * b
      s3 += re(a[k]) * im(b[k]) - im(a[k]) * re(b[k]);
    }
  } else {
    for (int r = lane; r < m; r += 32) {
      const T x = gp[r], y = gq[r];
      s0 += abs2_(x);
      s1 += abs2_(y);
      s2 += re(x) * re(y) + im(x) * im(y);
      s3 += re(x) * im(y) - im(x) * re(y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (ScalarTraits<T>::is_complex) s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  const double gabs2 = s2 * s2 + s3 * s3;
  if (gabs2 == 0.0 || gabs2 <= tol2 * s0 * s1) return false;
  // rotation (c, s e^{i phi}) that annihilates a^H b; reciprocal square roots instead of the sqrt / div chain (each FP64 divide
  // or square root is a long dependent sequence, and this scalar chain is the critical path of a round)
  const double rg = rsqrt(gabs2);
  const double zeta = (s1 - s0) * 0.5 * rg;
  if (!(fabs(zeta) < 1e150)) return false;   // rotation angle below 1e-150 (a column of denormal norm): identity, and zeta^2 would overflow
  const double w = 1.0 + zeta * zeta;
  const double tt = copysign(1.0, zeta) / (fabs(zeta) + w * rsqrt(w));
  const double c = rsqrt(1.0 + tt * tt), sn = c * tt;
  const double pr = s2 * rg, pi = -s3 * rg;   // e^{-i phi}
  const T cq = from_complex<T>(c * pr, c * pi), sq = from_complex<T>(-sn * pr, -sn * pi);
  const T cc = from_complex<T>(c, 0.0), ss = from_complex<T>(sn, 0.0);
  if (in_regs) {
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const int r = lane + 32 * k;
      if (r < m) {
        T na = mul_(cc, a[k]); fma_(na, sq, b[k]);
        T nb = mul_(ss, a[k]); fma_(nb, cq, b[k]);
        gp[r] = na; gq[r] = nb;
      }
    }
  } else {
    for (int r = lane; r < m; r += 32) {
      const T x = gp[r], y = gq[r];
      T na = mul_(cc, x); fma_(na, sq, y);
      T nb = mul_(ss, x); fma_(nb, cq, y);
      gp[r] = na; gq[r] = nb;
    }
  }
  for (int r = lane; r < nv; r += 32) {
    const T x = vp[r], y = vq[r];
    T na = mul_(cc, x); fma_(na, sq, y);
    T nb = mul_(ss, x); fma_(nb, cq, y);
    vp[r] = na; vq[r] = nb;
  }
  return true;
}

// register depth by column length: the Gram and rotation loops of a short column do not pay for the longest one
template <typename T>
__device__ __forceinline__ bool jc_pair(T* gp, T* gq, int m, T* vp, T* vq, int nv, double tol2, int lane) {
  constexpr int ITMAX = JcItems<T>::N;
  if (m <= 32) return jc_pair_it<T, 1>(gp, gq, m, vp, vq, nv, tol2, lane);
  if (m <= 64) return jc_pair_it<T, 2>(gp, gq, m, vp, vq, nv, tol2, lane);
  if (m <= 128 || ITMAX == 4) return jc_pair_it<T, 4>(gp, gq, m, vp, vq, nv, tol2, lane);
  return jc_pair_it<T, ITMAX>(gp, gq, m, vp, vq, nv, tol2, lane);
}

__device__ __forceinline__ bool jc_pair_of(int i, int round, int ring, int n, int* p, int* q) {
  int a, b;
  if (i == 0) { a = ring; b = round % ring; }
  else { a = (round + i) % ring; b = (round + ring - i) % ring; }
  if (a > b) { const int t = a; a = b; b = t; }
  *p = a; *q = b;
  return b < n;   // false: padded column
}

template <typename T>
__device__ __forceinline__ void jc_finish(const T* G, int64_t ldg, int m, int n, int sweep, int gw, int nw, int lane, double* out) {
  for (int c = gw; c < n; c += nw) {
    const T* gc = G + (int64_t)c * ldg;
    double s = 0.0;
    for (int r = lane; r < m; r += 32) s += abs2_(gc[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[c] = s;
  }
  if (gw == 0 && lane == 0) out[n] = (double)sweep;
}

template <typename T>
__global__ void __launch_bounds__(JS_THREADS) jacobi_smem_kernel(T* G, int64_t ldg, int m, T* V, int64_t ldv, int nv, int n, int npad,
                                                                 double tol2, int max_sweeps, int v_in_smem, double* out) {
  extern __shared__ __align__(16) unsigned char jc_smem_raw[];
  T* Gs = reinterpret_cast<T*>(jc_smem_raw);
  T* Vs = Gs + (size_t)m * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int npairs = npad / 2, ring = npad - 1;
  for (int e = threadIdx.x; e < m * n; e += blockDim.x) Gs[e] = G[(e % m) + (int64_t)(e / m) * ldg];
  if (v_in_smem)
    for (int e = threadIdx.x; e < nv * n; e += blockDim.x) Vs[e] = V[(e % nv) + (int64_t)(e / nv) * ldv];
  __syncthreads();
  T* Vb = v_in_smem ? Vs : V;
  const int64_t ldvb = v_in_smem ? nv : ldv;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    int my_rot = 0;
    for (int round = 0; round < ring; ++round) {
      for (int i = warp; i < npairs; i += nw) {
        int p, q;
        if (!jc_pair_of(i, round, ring, n, &p, &q)) continue;
        if (jc_pair<T>(Gs + (size_t)p * m, Gs + (size_t)q * m, m, Vb + p * ldvb, Vb + q * ldvb, nv, tol2, lane)) ++my_rot;
      }
      __syncthreads();
    }
    if (!__syncthreads_or(my_rot)) { ++sweep; break; }
  }
  for (int e = threadIdx.x; e < m * n; e += blockDim.x) G[(e % m) + (int64_t)(e / m) * ldg] = Gs[e];
  if (v_in_smem)
    for (int e = threadIdx.x; e < nv * n; e += blockDim.x) V[(e % nv) + (int64_t)(e / nv) * ldv] = Vs[e];
  jc_finish<T>(Gs, m, m, n, sweep, warp, nw, lane, out);
}

template <typename T>
__global__ void __launch_bounds__(JC_WARPS * 32) jacobi_cluster_kernel(T* G, int64_t ldg, int m, T* V, int64_t ldv, int nv, int n,
                                                                       int npad, double tol2, int max_sweeps, unsigned int* rot,
                                                                       double* out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (int)cluster.num_blocks() * JC_WARPS, gw = (int)cluster.block_rank() * JC_WARPS + warp;
  const int npairs = npad / 2, ring = npad - 1;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    unsigned int my_rot = 0;
    for (int round = 0; round < ring; ++round) {
      for (int i = gw; i < npairs; i += nw) {
        int p, q;
        if (!jc_pair_of(i, round, ring, n, &p, &q)) continue;
        if (jc_pair<T>(G + (int64_t)p * ldg, G + (int64_t)q * ldg, m, V + (int64_t)p * ldv, V + (int64_t)q * ldv, nv, tol2, lane)) ++my_rot;
      }
      if (round == ring - 1 && lane == 0 && my_rot) atomicAdd(rot + sweep, my_rot);
      cluster.sync();   // release / acquire at cluster scope: the columns written in this round are visible to their next owners
    }
    const unsigned int total = *reinterpret_cast<volatile unsigned int*>(rot + sweep);   // same value in every thread
    if (total == 0u) { ++sweep; break; }
  }
  jc_finish<T>(G, ldg, m, n, sweep, gw, nw, lane, out);
}

// ------------------------------------------------------------------------------------------------
// Tournament form for 112 < n <= 256 (the README-size bonds): a thread-block cluster of up to 8 CTAs, every warp owns one SLOT
// (a pair of columns of G and of V) in its CTA's shared memory.  A round = load the slot into registers, rotate, cluster
// barrier, store both columns into the slots the round-robin rotation sends them to (circle method: top row shifts right,
// bottom row shifts left, slot 0's top column stays) -- the neighbour's shared memory when the slot lives in another CTA
// (distributed shared memory, st.shared::cluster) -- cluster barrier.  No column ever goes through L2 / HBM between rounds and
// the per-round work of one SM stays at <= 16 pairs (the single-CTA kernel is issue-bound beyond ~50).
// ------------------------------------------------------------------------------------------------
constexpr int JD_WARPS = 16;

template <typename T, int IT>
__global__ void __launch_bounds__(JD_WARPS * 32) jacobi_dsmem_kernel(T* G, int64_t ldg, int m, T* V, int64_t ldv, int nv, int n, int spc,
                                                                     double tol2, int max_sweeps, unsigned int* rot, double* out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char jd_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int S = C * spc;                                  // slots = pairs per round (2 S players, ids >= n are padding)
  T* Gs = reinterpret_cast<T*>(jd_raw);                   // [spc][2][m]
  T* Vs = Gs + (size_t)spc * 2 * m;                       // [spc][2][nv]
  int* ids = reinterpret_cast<int*>(Vs + (size_t)spc * 2 * nv);   // [spc][2]
  const bool active = warp < spc;
  const int slot = rank * spc + warp;
  T a[IT], b[IT], va[IT], vb[IT];
  if (active) {                                           // initial seating: slot s holds columns 2 s (top) and 2 s + 1 (bottom)
    for (int pos = 0; pos < 2; ++pos) {
      const int id = 2 * slot + pos;
      T* g = Gs + ((size_t)warp * 2 + pos) * m;
      T* v = Vs + ((size_t)warp * 2 + pos) * nv;
      for (int r = lane; r < m; r += 32) g[r] = (id < n) ? G[r + (int64_t)id * ldg] : zero_<T>();
      for (int r = lane; r < nv; r += 32) v[r] = (id < n) ? V[r + (int64_t)id * ldv] : zero_<T>();
      if (lane == 0) ids[warp * 2 + pos] = id;
    }
  }
  cluster.sync();
  int sweep = 0;
  const int rounds = 2 * S - 1;
  for (; sweep < max_sweeps; ++sweep) {
    unsigned int my_rot = 0;
    for (int round = 0; round < rounds; ++round) {
      int id0 = n, id1 = n;
      if (active) {
        id0 = ids[warp * 2];
        id1 = ids[warp * 2 + 1];
        const T* g0 = Gs + (size_t)warp * 2 * m;
        const T* g1 = g0 + m;
        const T* v0 = Vs + (size_t)warp * 2 * nv;
        const T* v1 = v0 + nv;
#pragma unroll
        for (int k = 0; k < IT; ++k) {
          const int r = lane + 32 * k;
          a[k] = (r < m) ? g0[r] : zero_<T>();
          b[k] = (r < m) ? g1[r] : zero_<T>();
          va[k] = (r < nv) ? v0[r] : zero_<T>();
          vb[k] = (r < nv) ? v1[r] : zero_<T>();
        }
        if (id0 < n && id1 < n) {
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
          for (int k = 0; k < IT; ++k) {
            s0 += abs2_(a[k]);
            s1 += abs2_(b[k]);
            s2 += re(a[k]) * re(b[k]) + im(a[k]) * im(b[k]);
            s3 += re(a[k]) * im(b[k]) - im(a[k]) * re(b[k]);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            if (ScalarTraits<T>::is_complex) s3 += __shfl_xor_sync(0xffffffffu, s3, o);
          }
          const double gabs2 = s2 * s2 + s3 * s3;
          if (gabs2 != 0.0 && gabs2 > tol2 * s0 * s1) {
            const double rg = rsqrt(gabs2);
            const double zeta = (s1 - s0) * 0.5 * rg;
            if (fabs(zeta) < 1e150) {
              ++my_rot;
              const double w = 1.0 + zeta * zeta;
              const double tt = copysign(1.0, zeta) / (fabs(zeta) + w * rsqrt(w));
              const double c = rsqrt(1.0 + tt * tt), sn = c * tt;
              const double pr = s2 * rg, pi = -s3 * rg;
              const T cq = from_complex<T>(c * pr, c * pi), sq = from_complex<T>(-sn * pr, -sn * pi);
              const T cc = from_complex<T>(c, 0.0), ss = from_complex<T>(sn, 0.0);
#pragma unroll
              for (int k = 0; k < IT; ++k) {
                T na = mul_(cc, a[k]); fma_(na, sq, b[k]);
                T nb = mul_(ss, a[k]); fma_(nb, cq, b[k]);
                a[k] = na; b[k] = nb;
                T nva = mul_(cc, va[k]); fma_(nva, sq, vb[k]);
                T nvb = mul_(ss, va[k]); fma_(nvb, cq, vb[k]);
                va[k] = nva; vb[k] = nvb;
              }
            }
          }
        }
      }
      cluster.sync();          // every slot has been read
      if (active) {
        // circle method: top[0] stays; top[s] -> top[s + 1] (the last one turns into bottom[S - 1]); bottom[s] -> bottom[s - 1]
        // (bottom[0] turns into top[1])
        int d0 = slot, p0 = 0, d1 = slot, p1 = 1;
        if (S > 1) {
          if (slot == 0) { d0 = 0; p0 = 0; d1 = 1; p1 = 0; }
          else {
            if (slot < S - 1) { d0 = slot + 1; p0 = 0; } else { d0 = S - 1; p0 = 1; }
            d1 = slot - 1; p1 = 1;
          }
        }
        {
          T* gd = cluster.map_shared_rank(Gs + ((size_t)(d0 % spc) * 2 + p0) * m, d0 / spc);
          T* vd = cluster.map_shared_rank(Vs + ((size_t)(d0 % spc) * 2 + p0) * nv, d0 / spc);
          int* idd = cluster.map_shared_rank(ids + (d0 % spc) * 2 + p0, d0 / spc);
#pragma unroll
          for (int k = 0; k < IT; ++k) {
            const int r = lane + 32 * k;
            if (r < m) gd[r] = a[k];
            if (r < nv) vd[r] = va[k];
          }
          if (lane == 0) *idd = id0;
        }
        {
          T* gd = cluster.map_shared_rank(Gs + ((size_t)(d1 % spc) * 2 + p1) * m, d1 / spc);
          T* vd = cluster.map_shared_rank(Vs + ((size_t)(d1 % spc) * 2 + p1) * nv, d1 / spc);
          int* idd = cluster.map_shared_rank(ids + (d1 % spc) * 2 + p1, d1 / spc);
#pragma unroll
          for (int k = 0; k < IT; ++k) {
            const int r = lane + 32 * k;
            if (r < m) gd[r] = b[k];
            if (r < nv) vd[r] = vb[k];
          }
          if (lane == 0) *idd = id1;
        }
      }
      if (round == rounds - 1 && lane == 0 && my_rot) atomicAdd(rot + sweep, my_rot);
      cluster.sync();          // every slot has been rewritten
    }
    const unsigned int total = *reinterpret_cast<volatile unsigned int*>(rot + sweep);
    if (total == 0u) { ++sweep; break; }
  }
  if (active) {
    for (int pos = 0; pos < 2; ++pos) {
      const int id = ids[warp * 2 + pos];
      if (id >= n) continue;
      const T* g = Gs + ((size_t)warp * 2 + pos) * m;
      const T* v = Vs + ((size_t)warp * 2 + pos) * nv;
      double sq = 0.0;
      for (int r = lane; r < m; r += 32) { const T x = g[r]; G[r + (int64_t)id * ldg] = x; sq += abs2_(x); }
      for (int r = lane; r < nv; r += 32) V[r + (int64_t)id * ldv] = v[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) out[id] = sq;
    }
  }
  if (rank == 0 && threadIdx.x == 0) out[n] = (double)sweep;
}

template <typename T, int IT>
static void jacobi_dsmem_launch(Ctx* ctx, int csize, size_t smem, T* G, int64_t ldg, int m, T* V, int64_t ldv, int nv, int n, int spc,
                                double tol2, int max_sweeps, unsigned int* rot, double* out) {
  static bool configured[64] = {false};
  bool& c = configured[ctx->device & 63];
  if (!c) { NSB_CUDA(cudaFuncSetAttribute(jacobi_dsmem_kernel<T, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JS_SMEM_MAX)); c = true; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize);
  cfg.blockDim = dim3(JD_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NSB_CUDA(cudaLaunchKernelEx(&cfg, jacobi_dsmem_kernel<T, IT>, G, ldg, m, V, ldv, nv, n, spc, tol2, max_sweeps, rot, out));
  LAUNCH_CHECK(ctx);
}

// returns false when the shape does not fit the tournament kernel (columns longer than the register depth, too many pairs)
template <typename T>
static bool jacobi_dsmem(Ctx* ctx, T* G, int64_t ldg, int64_t m, int64_t n, T* V, int64_t ldv, int64_t nv, double* out_dev) {
  constexpr int ITMAX = JcItems<T>::N;
  const int64_t len = std::max(m, nv);
  if (len > 32 * ITMAX) return false;
  const int npad = (int)((n % 2) ? n + 1 : n);
  const int s0 = std::max(npad / 2, 1);
  // slots per CTA: one SM issues a round of 16 slots in ~1 us; spreading the slots over more CTAs of the cluster shortens the
  // round until the cluster barrier dominates (ctx option jacobi_dsmem_spc = most slots per CTA, while the cluster has room)
  const int spc_pref = std::max(1, std::min(JD_WARPS, ctx->opt.jacobi_dsmem_spc));
  int csize = 1;
  while (csize < 8 && csize * spc_pref < s0) csize *= 2;
  if (csize * JD_WARPS < s0) return false;
  const int spc = (s0 + csize - 1) / csize;
  const size_t smem = sizeof(T) * (size_t)spc * 2 * (size_t)(m + nv) + sizeof(int) * (size_t)spc * 2;
  if (smem > JS_SMEM_MAX) return false;
  const int max_sweeps = 60;
  const double tol = 10.0 * std::sqrt((double)std::max<int64_t>(m, 1)) * 2.220446049250313e-16;
  DevBuf rot(ctx, sizeof(unsigned int) * max_sweeps);
  NSB_CUDA(cudaMemsetAsync(rot.ptr, 0, sizeof(unsigned int) * max_sweeps, ctx->stream));
  unsigned int* r = (unsigned int*)rot.ptr;
  if (len <= 64) jacobi_dsmem_launch<T, 2>(ctx, csize, smem, G, ldg, (int)m, V, ldv, (int)nv, (int)n, spc, tol * tol, max_sweeps, r, out_dev);
  else if (len <= 128 || ITMAX == 4) jacobi_dsmem_launch<T, 4>(ctx, csize, smem, G, ldg, (int)m, V, ldv, (int)nv, (int)n, spc, tol * tol, max_sweeps, r, out_dev);
  else jacobi_dsmem_launch<T, ITMAX>(ctx, csize, smem, G, ldg, (int)m, V, ldv, (int)nv, (int)n, spc, tol * tol, max_sweeps, r, out_dev);
  return true;
}

// columns of G (m x n) orthogonalised in place, rotations accumulated in V (nv x n); out_dev: n + 1 doubles (see above)
template <typename T>
static void jacobi_single_launch(Ctx* ctx, T* G, int64_t ldg, int64_t m, int64_t n, T* V, int64_t ldv, int64_t nv, double* out_dev) {
  const int npad = (int)((n % 2) ? n + 1 : n);
  const int npairs = std::max(npad / 2, 1);
  const int max_sweeps = 60;
  // the Gram entries of orthogonal columns carry rounding noise ~ eps sqrt(m) (max over n^2 / 2 pairs several times that)
  const double tol = 10.0 * std::sqrt((double)std::max<int64_t>(m, 1)) * 2.220446049250313e-16;
  const double tol2 = tol * tol;
  const size_t gbytes = sizeof(T) * (size_t)m * n, vbytes = sizeof(T) * (size_t)nv * n;
  if (gbytes + vbytes <= JS_SMEM_MAX) {   // (measured: with V left in global memory the single CTA loses to the cluster form)
    const int v_in_smem = 1;
    const size_t smem = gbytes + vbytes;
    static bool configured[2][64] = {{false}};
    bool& c = configured[ScalarTraits<T>::is_complex ? 1 : 0][ctx->device & 63];
    if (!c) { NSB_CUDA(cudaFuncSetAttribute(jacobi_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JS_SMEM_MAX)); c = true; }
    int threads = 32 * std::min(npairs, JS_THREADS / 32);
    threads = std::max(threads, 128);   // (the staging loops and the final norms use every warp)
    jacobi_smem_kernel<T><<<1, threads, smem, ctx->stream>>>(G, ldg, (int)m, V, ldv, (int)nv, (int)n, npad, tol2, max_sweeps, v_in_smem, out_dev);
    LAUNCH_CHECK(ctx);
    return;
  }
  DevBuf rot(ctx, sizeof(unsigned int) * max_sweeps);
  NSB_CUDA(cudaMemsetAsync(rot.ptr, 0, sizeof(unsigned int) * max_sweeps, ctx->stream));
  int csize = 1;
  while (csize < 8 && csize * JC_WARPS < npairs) csize *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize);
  cfg.blockDim = dim3(JC_WARPS * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NSB_CUDA(cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel<T>, G, ldg, (int)m, V, ldv, (int)nv, (int)n, npad, tol2, max_sweeps,
                              (unsigned int*)rot.ptr, out_dev));
  LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------------
// blocked one-sided Jacobi: column blocks of width B are paired (round-robin over blocks); for each pair the
// 2B x 2B Gram matrix is formed by the DMMA GEMM, diagonalised inside one CTA (two-sided cyclic Jacobi on the
// upper triangle in shared memory, rotations accumulated in R), and the pair is updated with one GEMM
// [G_I G_J] <- [G_I G_J] R (same for V).  All O(m n^2) work per sweep is GEMM.
// ------------------------------------------------------------------------------------------------
template <typename T> struct JacobiBlk;
template <> struct JacobiBlk<double> { static constexpr int B = 64; };
template <> struct JacobiBlk<cdouble> { static constexpr int B = 32; };

template <typename T> __device__ __forceinline__ T cmul_conj_a(T a, T b);   // conj(a) * b
template <> __device__ __forceinline__ double cmul_conj_a<double>(double a, double b) { return a * b; }
template <> __device__ __forceinline__ cdouble cmul_conj_a<cdouble>(cdouble a, cdouble b) { return mul_(conj_(a), b); }

template <typename T, int N2>
__global__ void __launch_bounds__(1024) herm_eig_kernel(const T* __restrict__ Sg, T* __restrict__ Rg,
                                                        unsigned long long* __restrict__ maxoff, double abs_floor,
                                                        double outer_tol, int inner_cap, double* __restrict__ evals = nullptr,
                                                        int force = 0) {
  // evals (optional): N2 diagonal entries of the rotated matrix per problem (eigenvalue i <-> column i of R).
  // force: always run the inner sweeps (stand-alone eigensolver use: no outer iteration finishes the job).
  // abs_floor = eps * (largest diagonal entry of the global Gram matrix): couplings below it cannot change any
  // sigma^2 by more than LAPACK-level absolute accuracy and are treated as converged.
  constexpr int NP = N2 / 2, RING = N2 - 1, TRI = N2 * (N2 + 1) / 2;
  constexpr int LOG_N2 = (N2 == 128) ? 7 : 6, LOG_NP = LOG_N2 - 1;
  static_assert(N2 == 128 || N2 == 64, "N2 must be 64 or 128");
  extern __shared__ __align__(16) char smraw[];
  T* tri = reinterpret_cast<T*>(smraw);           // upper triangle, idx(i<=j) = i + j(j+1)/2
  T* R = tri + TRI;                               // N2 x N2 column-major
  T* J = R + N2 * N2;                             // NP x 4 : j11 j12 j21 j22
  __shared__ unsigned long long smax;
  __shared__ double red[40];
  __shared__ short pq[N2];                        // p[a] = pq[2a], q[a] = pq[2a+1]
  const int tid = threadIdx.x, nth = blockDim.x;
  const T* S = Sg + (size_t)blockIdx.x * N2 * N2;
  auto tix = [](int i, int j) { return i + ((j * (j + 1)) >> 1); };
  auto get = [&](int i, int j) -> T { return i <= j ? tri[tix(i, j)] : conj_(tri[tix(j, i)]); };
  auto put = [&](int i, int j, T v) { if (i <= j) tri[tix(i, j)] = v; else tri[tix(j, i)] = conj_(v); };
  for (int e = tid; e < N2 * N2; e += nth) {
    int i = e & (N2 - 1), j = e >> LOG_N2;
    if (i <= j) {
      T v = S[i + (size_t)j * N2];
      if (i == j) v = from_complex<T>(re(v), 0.0);
      tri[tix(i, j)] = v;
    }
    R[e] = from_complex<T>(i == j ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  {  // off-diagonal measure of the incoming Gram matrix (outer convergence), relative with the absolute floor
    double v = 0.0;
    for (int e = tid; e < N2 * N2; e += nth) {
      int i = e & (N2 - 1), j = e >> LOG_N2;
      if (i < j) {
        double a = sqrt(abs2_(tri[tix(i, j)]));
        double den = fmax(sqrt(fabs(re(tri[tix(i, i)]) * re(tri[tix(j, j)]))), abs_floor / outer_tol);
        if (a > 0.0 && den > 0.0) v = fmax(v, a / den);
      }
    }
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double m = 0.0;
      for (int w = 0; w < (nth + 31) / 32; ++w) m = fmax(m, red[w]);
      atomicMax(maxoff, (unsigned long long)__double_as_longlong(m));
      red[32] = m;
    }
    __syncthreads();
    if (!force && red[32] <= outer_tol) {   // this pair is already orthogonal to the outer tolerance: R = I exactly
      T* Ro = Rg + (size_t)blockIdx.x * N2 * N2;
      for (int e = tid; e < N2 * N2; e += nth) Ro[e] = R[e];
      return;
    }
  }
  const double tol = 2.220446049250313e-16 * (N2 / 2);   // rounding noise of the updated couplings is O(eps sqrt(N2))
  double afl = abs_floor;
  if (force) {   // stand-alone use: absolute floor from this problem's own scale, eps * max |S_ij|
    double v = 0.0;
    for (int e = tid; e < TRI; e += nth) v = fmax(v, sqrt(abs2_(tri[e])));
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double m = 0.0;
    for (int w = 0; w < (nth + 31) / 32; ++w) m = fmax(m, red[w]);
    afl = fmax(afl, 2.220446049250313e-16 * m);
    __syncthreads();
  }
  for (int sweep = 0; sweep < inner_cap; ++sweep) {
    if (tid == 0) { smax = 0ull; atomicAdd(maxoff + 1, 1ull); }
    __syncthreads();
    for (int r = 0; r < RING; ++r) {
      // phase A: rotation of every pair from its diagonal 2x2 block
      int rotated = 0;
      if (tid < NP) {
        int a = tid, p, q;
        if (a == 0) { p = RING; q = r; }
        else { p = r + a; if (p >= RING) p -= RING; q = r + RING - a; if (q >= RING) q -= RING; }
        pq[2 * a] = (short)p; pq[2 * a + 1] = (short)q;
        double alpha = re(tri[tix(p, p)]), beta = re(tri[tix(q, q)]);
        T g = get(p, q);
        const double g2 = abs2_(g), ab = fabs(alpha * beta);
        T j11 = from_complex<T>(1.0, 0.0), j12 = zero_<T>(), j21 = zero_<T>(), j22 = from_complex<T>(1.0, 0.0);
        if (g2 > afl * afl && g2 > tol * tol * ab) {
          // squared relative coupling for the sweep-level convergence test (no sqrt / divide on the critical path)
          atomicMax(&smax, (unsigned long long)__double_as_longlong(ab > 0.0 ? g2 / ab : 1.0));
          const double gabs = sqrt(g2), ginv = 1.0 / gabs;
          double zeta = (beta - alpha) * (0.5 * ginv);
          double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double c = rsqrt(1.0 + tt * tt), sn = c * tt;
          double pr = re(g) * ginv, pi = -im(g) * ginv;       // conj(phase)
          j11 = from_complex<T>(c, 0.0);
          j12 = from_complex<T>(sn, 0.0);
          j21 = from_complex<T>(-sn * pr, -sn * pi);
          j22 = from_complex<T>(c * pr, c * pi);
          tri[tix(p, p)] = from_complex<T>(alpha - tt * gabs, 0.0);
          tri[tix(q, q)] = from_complex<T>(beta + tt * gabs, 0.0);
          put(p, q, zero_<T>());
        }
        J[4 * a] = j11; J[4 * a + 1] = j12; J[4 * a + 2] = j21; J[4 * a + 3] = j22;
        rotated = (re(j12) != 0.0 || im(j12) != 0.0) ? 1 : 0;
      }
      if (!__syncthreads_or(rotated)) continue;   // no pair rotated in this step: nothing to update
      // phase B: off-diagonal pair-pair blocks  B' = Ja^H B Jb, and the columns of R
      for (int blk = tid; blk < NP * NP; blk += nth) {
        int a = blk >> LOG_NP, b = blk & (NP - 1);
        if (a >= b) continue;
        int pa = pq[2 * a], qa = pq[2 * a + 1], pb = pq[2 * b], qb = pq[2 * b + 1];
        T b11 = get(pa, pb), b12 = get(pa, qb), b21 = get(qa, pb), b22 = get(qa, qb);
        T a11 = J[4 * a], a12 = J[4 * a + 1], a21 = J[4 * a + 2], a22 = J[4 * a + 3];
        T c11 = J[4 * b], c12 = J[4 * b + 1], c21 = J[4 * b + 2], c22 = J[4 * b + 3];
        T m11 = add_(cmul_conj_a<T>(a11, b11), cmul_conj_a<T>(a21, b21));
        T m12 = add_(cmul_conj_a<T>(a11, b12), cmul_conj_a<T>(a21, b22));
        T m21 = add_(cmul_conj_a<T>(a12, b11), cmul_conj_a<T>(a22, b21));
        T m22 = add_(cmul_conj_a<T>(a12, b12), cmul_conj_a<T>(a22, b22));
        put(pa, pb, add_(mul_(m11, c11), mul_(m12, c21)));
        put(pa, qb, add_(mul_(m11, c12), mul_(m12, c22)));
        put(qa, pb, add_(mul_(m21, c11), mul_(m22, c21)));
        put(qa, qb, add_(mul_(m21, c12), mul_(m22, c22)));
      }
      for (int it = tid; it < N2 * NP; it += nth) {
        int i = it & (N2 - 1), a = it >> LOG_N2;
        int p = pq[2 * a], q = pq[2 * a + 1];
        T x = R[i + p * N2], y = R[i + q * N2];
        R[i + p * N2] = add_(mul_(x, J[4 * a]), mul_(y, J[4 * a + 2]));
        R[i + q * N2] = add_(mul_(x, J[4 * a + 1]), mul_(y, J[4 * a + 3]));
      }
      __syncthreads();
    }
    double off = __longlong_as_double((long long)smax);   // squared
    __syncthreads();
    if (off <= tol * tol) break;
  }
  T* Ro = Rg + (size_t)blockIdx.x * N2 * N2;
  for (int e = tid; e < N2 * N2; e += nth) Ro[e] = R[e];
  if (evals) for (int i = tid; i < N2; i += nth) evals[(size_t)blockIdx.x * N2 + i] = re(tri[tix(i, i)]);
}

// Batched dense symmetric eigensolver for 128 x 128 problems (leaves of the divide & conquer in eigh.cu):
// S, R: batch x 128 x 128 column-major; evals: batch x 128.  Eigenvalue i belongs to column i of R (unsorted).
void herm_eig_batch128(Ctx* ctx, const double* S, double* R, double* evals, int batch, double abs_floor, int max_sweeps) {
  if (batch <= 0) return;
  constexpr int N2 = 128;
  auto kern = herm_eig_kernel<double, N2>;
  size_t smem = sizeof(double) * ((size_t)N2 * (N2 + 1) / 2 + (size_t)N2 * N2 + 4 * (N2 / 2));
  static bool configured_dev[64] = {false};
  bool& configured = configured_dev[ctx->device & 63];   // function attributes are per device
  if (!configured) { NSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; }
  unsigned long long* dmax = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
  NSB_CUDA(cudaMemsetAsync(dmax, 0, 2 * sizeof(unsigned long long), ctx->stream));
  kern<<<(unsigned)batch, 1024, smem, ctx->stream>>>(S, R, dmax, abs_floor, 1e-300, max_sweeps, evals, 1);
  LAUNCH_CHECK(ctx);
}

template <typename T>
__global__ void block_gather_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows, int64_t bw,
                                    const int32_t* __restrict__ src_slot_of_dst, int nslots) {
  int64_t per = rows * bw, total = per * nslots;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t d = i / per, off = i % per;
    dst[i] = src[(int64_t)src_slot_of_dst[d] * per + off];
  }
}


template <typename T>
static int jacobi_blocked(Ctx* ctx, T* G, int64_t m, int64_t n, T* V, int64_t nv) {
  constexpr int B = JacobiBlk<T>::B, N2 = 2 * B;
  int64_t nblk = (n + B - 1) / B;
  if (nblk < 2) nblk = 2;
  if (nblk % 2) ++nblk;
  const int64_t npad = nblk * B, npairs = nblk / 2, h = npairs;
  DevBuf GA(ctx, sizeof(T) * m * npad), GB(ctx, sizeof(T) * m * npad), VA(ctx, sizeof(T) * nv * npad), VB(ctx, sizeof(T) * nv * npad);
  DevBuf Sb(ctx, sizeof(T) * N2 * N2 * npairs), Rb(ctx, sizeof(T) * N2 * N2 * npairs), mapb(ctx, sizeof(int32_t) * nblk);
  vec_zero<T>(ctx, m * npad, (T*)GA.ptr);
  vec_zero<T>(ctx, nv * npad, (T*)VA.ptr);
  // de Rijk ordering: columns enter sorted by decreasing norm (graded spectra converge in far fewer sweeps)
  double smax2 = 0.0;
  {
    DevBuf nrm(ctx, sizeof(double) * n), sidx(ctx, sizeof(int32_t) * n);
    col_norms2<T>(ctx, G, m, n, m, (double*)nrm.ptr);
    std::vector<double> hn(n);
    NSB_CUDA(cudaMemcpyAsync(hn.data(), nrm.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    for (double x : hn) smax2 += x;   // ||G||_F^2 >= sigma_max^2, preserved by the rotations
    std::vector<int32_t> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return hn[a] > hn[b]; });
    NSB_CUDA(cudaMemcpyAsync(sidx.ptr, ord.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    gather_cols<T>(ctx, G, m, m, (const int32_t*)sidx.ptr, n, nullptr, (T*)GA.ptr, m);
    gather_cols<T>(ctx, V, nv, nv, (const int32_t*)sidx.ptr, n, nullptr, (T*)VA.ptr, nv);
    ctx->sync();
  }
  // round-robin block permutation (constant over rounds): slot 2i = top[i], slot 2i+1 = bottom[i]
  std::vector<int32_t> src_of_dst(nblk), blockid(nblk), tmp(nblk);
  for (int64_t s = 0; s < nblk; ++s) { src_of_dst[s] = (int32_t)s; blockid[s] = (int32_t)s; }
  if (h > 1) {
    src_of_dst[0] = 0;
    src_of_dst[2] = 1;                                              // bottom[0] -> top[1]
    for (int64_t i = 2; i < h; ++i) src_of_dst[2 * i] = (int32_t)(2 * (i - 1));        // top[i-1] -> top[i]
    for (int64_t i = 0; i + 1 < h; ++i) src_of_dst[2 * i + 1] = (int32_t)(2 * (i + 1) + 1);  // bottom[i+1] -> bottom[i]
    src_of_dst[2 * (h - 1) + 1] = (int32_t)(2 * (h - 1));           // top[h-1] -> bottom[h-1]
  }
  NSB_CUDA(cudaMemcpyAsync(mapb.ptr, src_of_dst.data(), sizeof(int32_t) * nblk, cudaMemcpyHostToDevice, ctx->stream));
  auto kern = herm_eig_kernel<T, N2>;
  size_t smem = sizeof(T) * ((size_t)N2 * (N2 + 1) / 2 + (size_t)N2 * N2 + 4 * (N2 / 2));
  static bool configured_dev[64] = {false};
  bool& configured = configured_dev[ctx->device & 63];   // function attributes are per device
  if (!configured) { NSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; }
  const T one = from_complex<T>(1.0, 0.0), zero = zero_<T>();
  // the Gram entries of orthogonal columns carry rounding noise ~ eps sqrt(m) (max over n^2/2 pairs several times that)
  const double tol = 10.0 * std::sqrt((double)std::max<int64_t>(m, 1)) * 2.220446049250313e-16;
  unsigned long long* dmax = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
  T *ga = (T*)GA.ptr, *gb = (T*)GB.ptr, *va = (T*)VA.ptr, *vb = (T*)VB.ptr;
  const double abs_floor = 2.220446049250313e-16 * smax2;
  int sweep = 0;
  const int64_t rounds = std::max<int64_t>(nblk - 1, 1);
  for (; sweep < 30; ++sweep) {
    NSB_CUDA(cudaMemsetAsync(dmax, 0, 2 * sizeof(unsigned long long), ctx->stream));
    for (int64_t r = 0; r < rounds; ++r) {
      gemm<T>(ctx, OP_C, OP_N, N2, N2, m, one, ga, m, m * N2, ga, m, m * N2, zero, (T*)Sb.ptr, N2, (int64_t)N2 * N2, npairs);
      kern<<<(unsigned)npairs, 1024, smem, ctx->stream>>>((const T*)Sb.ptr, (T*)Rb.ptr, dmax, abs_floor, tol, ctx->opt.jacobi_inner_cap, nullptr, 0);
      LAUNCH_CHECK(ctx);
      gemm<T>(ctx, OP_N, OP_N, m, N2, N2, one, ga, m, m * N2, (const T*)Rb.ptr, N2, (int64_t)N2 * N2, zero, gb, m, m * N2, npairs);
      gemm<T>(ctx, OP_N, OP_N, nv, N2, N2, one, va, nv, nv * N2, (const T*)Rb.ptr, N2, (int64_t)N2 * N2, zero, vb, nv, nv * N2, npairs);
      int grid = ctx->num_sms * 8;
      block_gather_kernel<T><<<grid, 256, 0, ctx->stream>>>(gb, ga, m, B, (const int32_t*)mapb.ptr, (int)nblk);
      LAUNCH_CHECK(ctx);
      block_gather_kernel<T><<<grid, 256, 0, ctx->stream>>>(vb, va, nv, B, (const int32_t*)mapb.ptr, (int)nblk);
      LAUNCH_CHECK(ctx);
      for (int64_t s = 0; s < nblk; ++s) tmp[s] = blockid[src_of_dst[s]];
      blockid.swap(tmp);
    }
    NSB_CUDA(cudaMemcpyAsync(ctx->h_pinned, dmax, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NSB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->cnt.jacobi_sweeps++;
    if (getenv("NSB_DEBUG_JACOBI")) {
      unsigned long long inner; memcpy(&inner, &ctx->h_pinned[1], 8);
      fprintf(stderr, "[jacobi_blocked] n=%ld m=%ld sweep %d offmax %.3e tol %.3e floor %.3e inner sweeps/pair-solve %.2f\n", (long)n, (long)m,
              sweep, ctx->h_pinned[0], tol, abs_floor, (double)inner / (double)(rounds * npairs));
    }
    if (ctx->h_pinned[0] <= tol) { ++sweep; break; }
  }
  // collect the n real columns (padding columns never mix: their Gram rows are exactly zero)
  std::vector<int32_t> cols;
  cols.reserve(n);
  for (int64_t s = 0; s < nblk; ++s)
    for (int j = 0; j < B; ++j)
      if ((int64_t)blockid[s] * B + j < n) cols.push_back((int32_t)(s * B + j));
  DevBuf idx(ctx, sizeof(int32_t) * n);
  NSB_CUDA(cudaMemcpyAsync(idx.ptr, cols.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  gather_cols<T>(ctx, ga, m, m, (const int32_t*)idx.ptr, n, nullptr, G, m);
  gather_cols<T>(ctx, va, nv, nv, (const int32_t*)idx.ptr, n, nullptr, V, nv);
  ctx->sync();
  return sweep;
}

// Large matrices (rows <= cols, rows >= g_eigh_min_n): the reference's `eigen` route (App. A.4) -- rho = M M^H by
// one GEMM, Hermitian eigen-decomposition (tridiagonalisation + divide & conquer, eigh.cu), truncation rule on the
// eigenvalues, back-transformation of the kept eigenvectors only, C = U^H M by one GEMM.
template <typename T>
static FactorInfo factorize_left_eigh(Ctx* ctx, const T* M, int64_t rows, int64_t cols, int64_t ld, bool trans_in, double cutoff,
                                      int64_t mindim, int64_t maxdim, bool sqrt_spectrum, DevBuf& U, DevBuf& C,
                                      std::vector<double>& spectrum, const FactorDist* dist) {
  FactorInfo info;
  info.decomp = 2;   // the density-matrix `eigen` route (whatever label the reference's cutoff rule gives; 3 = refined, below)
  const int64_t n = rows;
  const T one = from_complex<T>(1.0, 0.0), zero = zero_<T>();
  // sqrt_spectrum with a square input: M is itself the Hermitian PSD density matrix of `eigen(rho; ...)`
  // (src/subspace/densitymatrix.jl:53): decompose it directly, its eigenvalues are the spectrum
  const bool direct = sqrt_spectrum && rows == cols;
  Eigh<T> eg;
  {
    DevBuf rho(ctx, sizeof(T) * (size_t)n * n);
    // the Gram matrix only needs its lower triangle when the symmetric tridiagonalisation kernel will read it
    const bool refine = !direct && !sqrt_spectrum && cutoff > 0.0 && cutoff <= 1e-12;
    const int G = (dist && !direct && !refine) ? dist->nranks : 1;
    const bool gram_split = G > 1 && n % G == 0 && n / G >= 128;
    if (!(G > 1)) dist = nullptr;
    const bool lower = !direct && !gram_split && Eigh<T>::reads_lower_only(ctx, n, n) && ctx->gemm_impl != GEMM_NAIVE;
    if (gram_split) {   // rho[:, slab] by its owner, all-gather completes the (full) matrix
      const int64_t nc = n / G, c0 = nc * dist->rank;
      if (!trans_in) gemm<T>(ctx, OP_N, OP_C, n, nc, cols, one, M, ld, 0, M + c0, ld, 0, zero, (T*)rho.ptr + c0 * n, n, 0, 1);
      else gemm<T>(ctx, OP_T, OP_CONJ, n, nc, cols, one, M, ld, 0, M + c0 * ld, ld, 0, zero, (T*)rho.ptr + c0 * n, n, 0, 1);
      dist->allgather_inplace(rho.ptr, sizeof(T) * (size_t)n * nc);
    } else if (direct) {
      if (!trans_in) copy_block<T>(ctx, M, ld, (T*)rho.ptr, n, n, n);
      else transpose_conj<T>(ctx, M, cols, rows, ld, (T*)rho.ptr, n, false);
    } else if (!trans_in) {
      gemm<T>(ctx, OP_N, OP_C, n, n, cols, one, M, ld, 0, M, ld, 0, zero, (T*)rho.ptr, n, 0, 1, GEMM_AUTO, nullptr, lower ? GEMM_LOWER_ONLY : 0);
    } else {   // logical M(r, c) = buf[c + r ld]:  rho = buf^T conj(buf)
      gemm<T>(ctx, OP_T, OP_CONJ, n, n, cols, one, M, ld, 0, M, ld, 0, zero, (T*)rho.ptr, n, 0, 1, GEMM_AUTO, nullptr, lower ? GEMM_LOWER_ONLY : 0);
    }
    eg.factor(ctx, (T*)rho.ptr, n, n, lower);
  }
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return eg.w[a] > eg.w[b]; });
  spectrum.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    const double lam = eg.w[order[i]];
    spectrum[i] = (sqrt_spectrum && !direct) ? std::sqrt(std::max(lam, 0.0)) : lam;
  }
  if (!direct && !sqrt_spectrum && cutoff > 0.0 && cutoff <= 1e-12) {
    // The reference takes LAPACK SVD for cutoff <= 1e-12 (App. A.4) because the eigenvalues of rho carry an absolute error of
    // eps * lambda_1, which is the size of the weights such a cutoff decides on.  Refinement: all eigenvectors, G = M^H U; the
    // squared column norms of G are the Rayleigh quotients u_i^H rho u_i = |u_i^H M|^2, computed from M itself -- their error
    // is quadratic in the eigenvector error (which only mixes states of nearly equal weight), so the truncation rule sees
    // every sigma_i^2 with relative accuracy.  U = kept columns, C = (kept columns of G)^H.
    info.decomp = 3;
    DevBuf Uall(ctx, sizeof(T) * (size_t)n * n), G(ctx, sizeof(T) * (size_t)cols * n), nrm(ctx, sizeof(double) * n);
    eg.vectors(order.data(), n, (T*)Uall.ptr, n);
    if (!trans_in) gemm<T>(ctx, OP_C, OP_N, cols, n, rows, one, M, ld, 0, (const T*)Uall.ptr, n, 0, zero, (T*)G.ptr, cols, 0, 1);
    else gemm<T>(ctx, OP_CONJ, OP_N, cols, n, rows, one, M, ld, 0, (const T*)Uall.ptr, n, 0, zero, (T*)G.ptr, cols, 0, 1);
    col_norms2<T>(ctx, (const T*)G.ptr, cols, n, cols, (double*)nrm.ptr);
    std::vector<double> P(n);
    NSB_CUDA(cudaMemcpyAsync(P.data(), nrm.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    std::vector<int32_t> o2(n);
    std::iota(o2.begin(), o2.end(), 0);
    std::stable_sort(o2.begin(), o2.end(), [&](int32_t a, int32_t b) { return P[a] > P[b]; });
    for (int64_t i = 0; i < n; ++i) spectrum[i] = P[o2[i]];
    double terr2 = 0.0;
    const int64_t nk = truncate_spectrum(spectrum, cutoff, mindim, maxdim, &terr2);
    info.newdim = nk;
    info.truncerr = terr2;
    DevBuf idx(ctx, sizeof(int32_t) * nk), tmp(ctx, sizeof(T) * (size_t)cols * nk);
    NSB_CUDA(cudaMemcpyAsync(idx.ptr, o2.data(), sizeof(int32_t) * nk, cudaMemcpyHostToDevice, ctx->stream));
    U = DevBuf(ctx, sizeof(T) * rows * nk);
    C = DevBuf(ctx, sizeof(T) * nk * cols);
    gather_cols<T>(ctx, (const T*)Uall.ptr, n, rows, (const int32_t*)idx.ptr, nk, nullptr, (T*)U.ptr, rows);
    gather_cols<T>(ctx, (const T*)G.ptr, cols, cols, (const int32_t*)idx.ptr, nk, nullptr, (T*)tmp.ptr, cols);
    transpose_conj<T>(ctx, (const T*)tmp.ptr, cols, nk, cols, (T*)C.ptr, nk, true);
    ctx->sync();
    return info;
  }
  double terr = 0.0;
  const int64_t nkeep = truncate_spectrum(spectrum, cutoff, mindim, maxdim, &terr);
  for (double& x : spectrum) x = std::max(x, 0.0);
  info.newdim = nkeep;
  info.truncerr = terr;
  if (dist && nkeep >= 64 * dist->nranks) {
    // kept eigenvectors by column slabs (equal chunks, the buffer is padded to a multiple of the rank count) + all-gather
    const int G = dist->nranks;
    const int64_t kc = (nkeep + G - 1) / G, k0 = std::min<int64_t>(nkeep, kc * dist->rank), kn = std::min<int64_t>(nkeep, k0 + kc) - k0;
    U = DevBuf(ctx, sizeof(T) * rows * kc * G);
    if (kn > 0) eg.vectors(order.data() + k0, kn, (T*)U.ptr + k0 * rows, rows);
    dist->allgather_inplace(U.ptr, sizeof(T) * (size_t)rows * kc);
    if (trans_in && dist->allow_c_transposed) {
      // C^T (cols x nkeep) = buf conj(U): split over the same column slabs of U
      C = DevBuf(ctx, sizeof(T) * cols * kc * G);
      if (kn > 0) gemm<T>(ctx, OP_N, OP_CONJ, cols, kn, rows, one, M, ld, 0, (const T*)U.ptr + k0 * rows, rows, 0, zero, (T*)C.ptr + k0 * cols, cols, 0, 1);
      dist->allgather_inplace(C.ptr, sizeof(T) * (size_t)cols * kc);
      info.c_transposed = true;
    } else if (!trans_in && cols % G == 0) {
      const int64_t nc = cols / G, c0 = nc * dist->rank;
      C = DevBuf(ctx, sizeof(T) * nkeep * cols);
      gemm<T>(ctx, OP_C, OP_N, nkeep, nc, rows, one, (const T*)U.ptr, rows, 0, M + c0 * ld, ld, 0, zero, (T*)C.ptr + c0 * nkeep, nkeep, 0, 1);
      dist->allgather_inplace(C.ptr, sizeof(T) * (size_t)nkeep * nc);
    } else {
      C = DevBuf(ctx, sizeof(T) * nkeep * cols);
      gemm<T>(ctx, OP_C, trans_in ? OP_T : OP_N, nkeep, cols, rows, one, (const T*)U.ptr, rows, 0, M, ld, 0, zero, (T*)C.ptr, nkeep, 0, 1);
    }
    ctx->sync();
    return info;
  }
  U = DevBuf(ctx, sizeof(T) * rows * nkeep);
  C = DevBuf(ctx, sizeof(T) * nkeep * cols);
  eg.vectors(order.data(), nkeep, (T*)U.ptr, rows);
  gemm<T>(ctx, OP_C, trans_in ? OP_T : OP_N, nkeep, cols, rows, one, (const T*)U.ptr, rows, 0, M, ld, 0, zero, (T*)C.ptr, nkeep, 0, 1);
  ctx->sync();
  return info;
}

template <typename T>
FactorInfo factorize_left(Ctx* ctx, const T* M, int64_t rows, int64_t cols, int64_t ld, bool trans_in, double cutoff,
                          int64_t mindim, int64_t maxdim, bool sqrt_spectrum, DevBuf& U, DevBuf& C,
                          std::vector<double>& spectrum, const FactorDist* dist) {
  // the density matrix of `eigen(rho; ...)` (square, sqrt_spectrum) is an eigenproblem already: the tridiagonalisation
  // route beats the latency-bound Jacobi iteration from ~100 rows on (README workload: 1-site DMRG 10.9 s -> 4.7 s)
  const bool direct_eig = sqrt_spectrum && rows == cols && ctx->opt.eigh_direct_min_n > 0 && rows >= ctx->opt.eigh_direct_min_n;
  if ((ctx->opt.eigh_min_n > 0 && rows <= cols && rows >= ctx->opt.eigh_min_n) || direct_eig) {
    ctx->cnt.svd_calls++;
    return factorize_left_eigh<T>(ctx, M, rows, cols, ld, trans_in, cutoff, mindim, std::min<int64_t>(maxdim, rows), sqrt_spectrum, U, C, spectrum, dist);
  }
  FactorInfo info;
  ctx->cnt.svd_calls++;
  const int64_t k = std::min(rows, cols);
  NSB_REQUIRE(k > 0, NSB_EINVAL, "factorize: empty matrix");
  info.decomp = (cutoff <= 1e-12) ? 1 : 2;
  maxdim = std::min<int64_t>(maxdim, k);
  if (rows > cols) {
    // tall matrix: thin QR first, then factorise the square R from the left so that U = Q U_r is orthonormal
    // by construction (product of orthogonal transformations)
    DevBuf Mc(ctx, sizeof(T) * rows * cols), Q(ctx, sizeof(T) * rows * cols), R(ctx, sizeof(T) * cols * cols);
    {
      HostProf hp(ctx, "factorize.tall_qr");
      if (!trans_in) copy_block<T>(ctx, M, ld, (T*)Mc.ptr, rows, rows, cols);
      else transpose_conj<T>(ctx, M, cols, rows, ld, (T*)Mc.ptr, rows, false);
      qr_thin<T>(ctx, (T*)Mc.ptr, rows, cols, rows, (T*)Q.ptr, rows, (T*)R.ptr, cols);
    }
    DevBuf Ur;
    ctx->cnt.svd_calls--;
    FactorInfo fi = factorize_left<T>(ctx, (const T*)R.ptr, cols, cols, cols, false, cutoff, mindim, maxdim, sqrt_spectrum, Ur, C, spectrum);
    U = DevBuf(ctx, sizeof(T) * rows * fi.newdim);
    gemm<T>(ctx, OP_N, OP_N, rows, fi.newdim, cols, from_complex<T>(1.0, 0.0), (const T*)Q.ptr, rows, 0, (const T*)Ur.ptr, cols, 0,
            zero_<T>(), (T*)U.ptr, rows, 0, 1);
    ctx->sync();
    return fi;
  }
  const bool left = true;           // rows <= cols: rotate the row side, U = accumulated rotations
  const int64_t n = rows, m = cols;
  HostProf hp_all(ctx, "factorize.jacobi_route");
  DevBuf G(ctx, sizeof(T) * m * n), V(ctx, sizeof(T) * n * n), Qm;
  std::vector<int32_t> pcol;        // column permutation of the preconditioned path
  const bool precond = (n >= ctx->opt.jacobi_block_min_n) && ctx->opt.jacobi_precondition && (n >= ctx->opt.jacobi_precondition_min_n);
  if (!precond) {   // G = M^H (cols x rows)
    if (!trans_in) transpose_conj<T>(ctx, M, rows, cols, ld, (T*)G.ptr, m, true);
    else conj_copy_block<T>(ctx, M, ld, (T*)G.ptr, m, cols, rows);            // stored (cols x rows): conj only
  } else {
    // Drmac-Veselic preconditioning: M P = Qm Rm (Householder QR, columns pre-sorted by decreasing norm), then
    // one-sided Jacobi on X = Rm^H (lower trapezoidal), whose columns are already nearly orthogonal:
    //   X J = W Sigma  =>  M P = (Qm J) Sigma W^H,  U = Qm J is a product of orthogonal transformations.
    DevBuf Mw(ctx, sizeof(T) * rows * cols);
    if (!trans_in) copy_block<T>(ctx, M, ld, (T*)Mw.ptr, rows, rows, cols);
    else transpose_conj<T>(ctx, M, cols, rows, ld, (T*)Mw.ptr, rows, false);
    Qm = DevBuf(ctx, sizeof(T) * rows * rows);
    DevBuf Rm(ctx, sizeof(T) * rows * cols);
    if (ctx->opt.jacobi_pivot) {
      qr_pivoted_thin<T>(ctx, (T*)Mw.ptr, rows, cols, rows, (T*)Qm.ptr, rows, (T*)Rm.ptr, rows, pcol);
    } else {
      pcol.resize(cols);
      std::iota(pcol.begin(), pcol.end(), 0);
      qr_thin<T>(ctx, (T*)Mw.ptr, rows, cols, rows, (T*)Qm.ptr, rows, (T*)Rm.ptr, rows);
    }
    transpose_conj<T>(ctx, (T*)Rm.ptr, rows, cols, rows, (T*)G.ptr, m, true);   // G = Rm^H (cols x rows)
    ctx->sync();
  }
  set_identity<T>(ctx, (T*)V.ptr, n, n, n);
  DevBuf norms(ctx, sizeof(double) * (n + 1));
  std::vector<double> P(n + 1);
  // small matrices: the whole Jacobi iteration as one cluster launch (no launch per round, no host round trip per sweep)
  const bool cluster = !precond && ctx->opt.jacobi_cluster_max_n > 0 && n >= 2 && n <= ctx->opt.jacobi_cluster_max_n && m <= 8192;
  // measured per sweep on B200 (graded test matrices, profiles/r02_perf_small_svd.log): the tournament kernel wins from n ~ 48
  // up to its limit of 256 columns (n = 100: 0.23 ms against 0.51 single CTA and 0.63 blocked; n = 166: 0.43 against 1.8 blocked)
  bool tournament = false;
  if (!precond && n >= ctx->opt.jacobi_dsmem_min_n && n >= 2 && n <= ctx->opt.jacobi_dsmem_max_n) {
    HostProf hp(ctx, "factorize.jacobi_dsmem");
    tournament = jacobi_dsmem<T>(ctx, (T*)G.ptr, m, m, n, (T*)V.ptr, n, n, (double*)norms.ptr);
  }
  if (tournament) {
    NSB_CUDA(cudaMemcpyAsync(P.data(), norms.ptr, sizeof(double) * (n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    info.sweeps = (int)P[n];
    ctx->cnt.jacobi_sweeps += info.sweeps;
  } else if (cluster) {
    HostProf hp(ctx, "factorize.jacobi_1launch");
    jacobi_single_launch<T>(ctx, (T*)G.ptr, m, m, n, (T*)V.ptr, n, n, (double*)norms.ptr);
    NSB_CUDA(cudaMemcpyAsync(P.data(), norms.ptr, sizeof(double) * (n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    info.sweeps = (int)P[n];
    ctx->cnt.jacobi_sweeps += info.sweeps;
  } else {
    HostProf hp(ctx, "factorize.jacobi_rounds");
    if (n >= ctx->opt.jacobi_block_min_n) info.sweeps = jacobi_blocked<T>(ctx, (T*)G.ptr, m, n, (T*)V.ptr, n);
    else info.sweeps = jacobi_onesided<T>(ctx, (T*)G.ptr, m, m, n, (T*)V.ptr, n, n);
    col_norms2<T>(ctx, (T*)G.ptr, m, n, m, (double*)norms.ptr);
    NSB_CUDA(cudaMemcpyAsync(P.data(), norms.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
  }
  P.resize(n);
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return P[a] > P[b]; });
  spectrum.resize(k);
  for (int64_t i = 0; i < k; ++i) spectrum[i] = sqrt_spectrum ? std::sqrt(std::max(P[order[i]], 0.0)) : P[order[i]];
  double terr = 0.0;
  int64_t nkeep = truncate_spectrum(spectrum, cutoff, mindim, maxdim, &terr);
  info.newdim = nkeep;
  info.truncerr = terr;

  HostProf hp_out(ctx, "factorize.gather_UC");
  DevBuf idx(ctx, sizeof(int32_t) * nkeep), scl(ctx, sizeof(double) * nkeep);
  NSB_CUDA(cudaMemcpyAsync(idx.ptr, order.data(), sizeof(int32_t) * nkeep, cudaMemcpyHostToDevice, ctx->stream));
  U = DevBuf(ctx, sizeof(T) * rows * nkeep);
  C = DevBuf(ctx, sizeof(T) * nkeep * cols);
  (void)left;
  {
    // U = V[:, order] (rows x nkeep)  [times Qm when preconditioned];  C = (G[:, order])^H  [columns un-permuted]
    DevBuf tmp(ctx, sizeof(T) * cols * nkeep);
    gather_cols<T>(ctx, (T*)G.ptr, m, cols, (int32_t*)idx.ptr, nkeep, nullptr, (T*)tmp.ptr, cols);
    if (!precond) {
      gather_cols<T>(ctx, (T*)V.ptr, n, rows, (int32_t*)idx.ptr, nkeep, nullptr, (T*)U.ptr, rows);
      transpose_conj<T>(ctx, (T*)tmp.ptr, cols, nkeep, cols, (T*)C.ptr, nkeep, true);
    } else {
      DevBuf Vk(ctx, sizeof(T) * rows * nkeep), tmp2(ctx, sizeof(T) * cols * nkeep), inv(ctx, sizeof(int32_t) * cols);
      gather_cols<T>(ctx, (T*)V.ptr, n, rows, (int32_t*)idx.ptr, nkeep, nullptr, (T*)Vk.ptr, rows);
      gemm<T>(ctx, OP_N, OP_N, rows, nkeep, rows, from_complex<T>(1.0, 0.0), (const T*)Qm.ptr, rows, 0, (const T*)Vk.ptr, rows, 0,
              zero_<T>(), (T*)U.ptr, rows, 0, 1);
      std::vector<int32_t> ip(cols);
      for (int64_t j = 0; j < cols; ++j) ip[pcol[j]] = (int32_t)j;       // original column c sits at permuted position ip[c]
      NSB_CUDA(cudaMemcpyAsync(inv.ptr, ip.data(), sizeof(int32_t) * cols, cudaMemcpyHostToDevice, ctx->stream));
      gather_rows<T>(ctx, (T*)tmp.ptr, cols, (const int32_t*)inv.ptr, cols, nkeep, (T*)tmp2.ptr, cols);
      transpose_conj<T>(ctx, (T*)tmp2.ptr, cols, nkeep, cols, (T*)C.ptr, nkeep, true);
    }
    ctx->sync();
  }
  return info;
}

#define INST(T)                                                                                              \
  template void qr_thin<T>(Ctx*, T*, int64_t, int64_t, int64_t, T*, int64_t, T*, int64_t);                   \
  template FactorInfo factorize_left<T>(Ctx*, const T*, int64_t, int64_t, int64_t, bool, double, int64_t, int64_t, \
                                        bool, DevBuf&, DevBuf&, std::vector<double>&, const FactorDist*);
INST(double)
INST(cdouble)

}  // namespace nsb
