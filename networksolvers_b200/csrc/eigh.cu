// Hermitian eigensolver on device: blocked tridiagonalisation, divide & conquer, WY back-transformation (eigh.h).
// The formulas follow tools/proto_eigh.py statement by statement (NumPy prototype, checked against LAPACK);
// the D&C bookkeeping (dc_secular.h) is additionally exercised on the CPU by tests/test_cpu_dc.py.
#include "eigh.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>

#include <cooperative_groups.h>

#include "dc_secular.h"
#include "gemm.h"
#include "linalg.h"
#include "ops.h"

namespace nsb {


#define LAUNCH_CHECK(ctx) do { (ctx)->cnt.kernel_launches++; NSB_CUDA(cudaGetLastError()); } while (0)

static bool eigh_debug() { static int v = -1; if (v < 0) v = getenv("NSB_DEBUG_EIGH") ? 1 : 0; return v == 1; }
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

namespace {

constexpr int MAXNB = 128;

__device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
__device__ __forceinline__ double sub_(double a, double b) { return a - b; }
__device__ __forceinline__ cdouble sub_(cdouble a, cdouble b) { return make_cuDoubleComplex(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double shfl_xor_T(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ cdouble shfl_xor_T(cdouble v, int o) {
  return make_cuDoubleComplex(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}
__device__ __forceinline__ double neg_(double a) { return -a; }
__device__ __forceinline__ cdouble neg_(cdouble a) { return make_cuDoubleComplex(-a.x, -a.y); }

// block-wide sums of NV doubles per thread (blockDim <= 1024); result valid in all threads
template <int NV>
__device__ __forceinline__ void blk_sum(double (&v)[NV], double* sh /* NV * 32 + NV doubles */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[i * 32 + w] = v[i];
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int j = 0; j < nw; ++j) s += sh[threadIdx.x * 32 + j];
    sh[NV * 32 + threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = sh[NV * 32 + i];
}

// ------------------------------------------------------------------------------------------------
// stage 1: tridiagonalisation (LAPACK xHETRD 'L' convention, full storage)
// ------------------------------------------------------------------------------------------------
// (1) column j of A is brought up to date with the i reflectors of the current panel:
//     A[j:, j] -= V[j:, :i] conj(W[j, :i]) + W[j:, :i] conj(V[j, :i]);  d[j] = Re A[j, j]
template <typename T>
__global__ void __launch_bounds__(256) trd_col_update_kernel(T* __restrict__ A, int64_t lda, int64_t n, int64_t j,
                                                             const T* __restrict__ V, const T* __restrict__ W, int64_t ldp,
                                                             int i, double* __restrict__ d_out) {
  __shared__ T sv[MAXNB], sw[MAXNB];
  for (int k = threadIdx.x; k < i; k += blockDim.x) {
    sv[k] = conj_(V[j + (int64_t)k * ldp]);
    sw[k] = conj_(W[j + (int64_t)k * ldp]);
  }
  __syncthreads();
  const int64_t r = j + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  T acc = zero_<T>();
  for (int k = 0; k < i; ++k) {
    fma_(acc, V[r + (int64_t)k * ldp], sw[k]);
    fma_(acc, W[r + (int64_t)k * ldp], sv[k]);
  }
  const T a = sub_(A[r + j * lda], acc);
  A[r + j * lda] = a;
  if (r == j) d_out[j] = re(a);
}

// (2) Householder reflector from x = A[j+1:, j]:  (I - tau v v^H)^H x = beta e_0, beta real, v[0] = 1.
//     vcol is the panel column (row 0 based); rows <= j stay zero.
template <typename T>
__global__ void __launch_bounds__(256) trd_house_kernel(const T* __restrict__ A, int64_t lda, int64_t n, int64_t j,
                                                        T* __restrict__ vcol, T* __restrict__ taus, double* __restrict__ e_out) {
  __shared__ double sh[33];
  const T* x = A + j * lda;
  double s[1] = {0.0};
  for (int64_t r = j + 2 + threadIdx.x; r < n; r += blockDim.x) s[0] += abs2_(x[r]);
  blk_sum<1>(s, sh);
  const T alpha = x[j + 1];
  const double sigma = s[0], ar = re(alpha), ai = im(alpha);
  if (sigma == 0.0 && ai == 0.0) {
    for (int64_t r = j + 2 + threadIdx.x; r < n; r += blockDim.x) vcol[r] = zero_<T>();
    if (threadIdx.x == 0) { vcol[j + 1] = from_complex<T>(1.0, 0.0); taus[j] = zero_<T>(); e_out[j] = ar; }
    return;
  }
  const double beta = -copysign(sqrt(ar * ar + ai * ai + sigma), ar);
  const double dr = ar - beta, di = ai, den = dr * dr + di * di;
  const T scale = from_complex<T>(dr / den, -di / den);   // 1 / (alpha - beta)
  for (int64_t r = j + 2 + threadIdx.x; r < n; r += blockDim.x) vcol[r] = mul_(scale, x[r]);
  if (threadIdx.x == 0) {
    vcol[j + 1] = from_complex<T>(1.0, 0.0);
    taus[j] = from_complex<T>((beta - ar) / beta, -ai / beta);
    e_out[j] = beta;
  }
}

// (3) out[c] = sum_r conj(col_c[r]) v[r] over the columns of [A_trail (m x m) | W (m x i) | V (m x i)]:
//     the Hermitian matrix-vector product as column dot products (coalesced), plus W^H v and V^H v.
template <typename T, int CPB>
__global__ void __launch_bounds__(256) trd_gemv_kernel(const T* __restrict__ At, int64_t lda, int64_t m,
                                                       const T* __restrict__ Wp, const T* __restrict__ Vp, int64_t ldp, int i,
                                                       const T* __restrict__ v, T* __restrict__ out) {
  __shared__ double sh[2 * CPB * 32 + 2 * CPB];
  const int64_t ncol = m + 2 * (int64_t)i, c0 = (int64_t)blockIdx.x * CPB;
  const T* cols[CPB];
#pragma unroll
  for (int q = 0; q < CPB; ++q) {
    const int64_t c = c0 + q;
    cols[q] = c < m ? At + c * lda : (c < m + i ? Wp + (c - m) * ldp : (c < ncol ? Vp + (c - m - i) * ldp : nullptr));
  }
  double acc[2 * CPB];
#pragma unroll
  for (int q = 0; q < 2 * CPB; ++q) acc[q] = 0.0;
  for (int64_t r = threadIdx.x; r < m; r += blockDim.x) {
    const T x = v[r];
#pragma unroll
    for (int q = 0; q < CPB; ++q) {
      if (cols[q]) {
        const T a = cols[q][r];
        acc[2 * q] += re(a) * re(x) + im(a) * im(x);
        acc[2 * q + 1] += re(a) * im(x) - im(a) * re(x);
      }
    }
  }
  blk_sum<2 * CPB>(acc, sh);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < CPB; ++q)
      if (c0 + q < ncol) out[c0 + q] = from_complex<T>(acc[2 * q], acc[2 * q + 1]);
  }
}

// (4a) w0 = tau (y - V p1 - W p2), p1 = W^H v, p2 = V^H v;  per-CTA partial of w0^H v
template <typename T>
__global__ void __launch_bounds__(256) trd_w1_kernel(const T* __restrict__ y, const T* __restrict__ Vp, T* __restrict__ Wp,
                                                     int64_t ldp, int i, int64_t m, const T* __restrict__ tau_j,
                                                     double* __restrict__ partial) {
  __shared__ T p1[MAXNB], p2[MAXNB];
  __shared__ double sh[2 * 32 + 2];
  for (int k = threadIdx.x; k < i; k += blockDim.x) { p1[k] = y[m + k]; p2[k] = y[m + i + k]; }
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double dot[2] = {0.0, 0.0};
  if (r < m) {
    T acc = y[r];
    for (int k = 0; k < i; ++k) {
      fma_(acc, neg_(Vp[r + (int64_t)k * ldp]), p1[k]);
      fma_(acc, neg_(Wp[r + (int64_t)k * ldp]), p2[k]);
    }
    const T w0 = mul_(*tau_j, acc);
    Wp[r + (int64_t)i * ldp] = w0;
    const T vv = Vp[r + (int64_t)i * ldp];
    dot[0] = re(w0) * re(vv) + im(w0) * im(vv);
    dot[1] = re(w0) * im(vv) - im(w0) * re(vv);
  }
  blk_sum<2>(dot, sh);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = dot[0]; partial[2 * blockIdx.x + 1] = dot[1]; }
}

// (4b) w = w0 - 1/2 tau (w0^H v) v
template <typename T>
__global__ void __launch_bounds__(256) trd_w2_kernel(const T* __restrict__ Vp, T* __restrict__ Wp, int64_t ldp, int i, int64_t m,
                                                     const T* __restrict__ tau_j, const double* __restrict__ partial, int nparts) {
  double sr = 0.0, si = 0.0;
  for (int k = 0; k < nparts; ++k) { sr += partial[2 * k]; si += partial[2 * k + 1]; }
  const T alpha = mul_(*tau_j, from_complex<T>(-0.5 * sr, -0.5 * si));
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m) {
    T w = Wp[r + (int64_t)i * ldp];
    fma_(w, alpha, Vp[r + (int64_t)i * ldp]);
    Wp[r + (int64_t)i * ldp] = w;
  }
}

// ------------------------------------------------------------------------------------------------
// One tridiagonalisation panel as a single cooperative kernel: per column two grid-wide barriers.
//   phase AD(i): finish w_{i-1} = tau (y - V p1 - W p2) + alpha v on the thread's own rows, bring column j = p + i
//                up to date, partial sums of |A[j+2:, j]|^2                                         -> barrier
//   phase C(i):  every CTA derives (beta, tau, scale) from the partials, v = scale x (v[0] = 1) is formed on the
//                fly and stored, persistent column-dot products y = A_trail^H v, p1 = W^H v, p2 = V^H v,
//                partial sums of y^H v                                                              -> barrier
// alpha = -1/2 tau (w0^H v) needs no third barrier: w0^H v = conj(tau) (y^H v - 2 Re p1^H p2) because V^H v = p2
// and W^H v = p1.  Same arithmetic as the five-kernel path up to the order of the reductions.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct TrdPanelArgs {
  T* A; int64_t lda, n, p; int w;
  T* Vp; T* Wp; int64_t ldp;      // panel column 0, row 0 based
  T* taus; double* d; double* e;
  T* y;                           // >= n + 2 MAXNB scratch: y (m), p1 (i), p2 (i) of the current column
  double* part;                   // gridDim.x partials of sigma
  double* part2;                  // 2 gridDim.x partials of y^H v
};

template <typename T, int CPB>
__global__ void __launch_bounds__(256) trd_panel_kernel(const TrdPanelArgs<T> a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  constexpr int GEMV_CHUNK = 128;
  constexpr int CSET = ScalarTraits<T>::is_complex ? 8 : 16;   // columns streamed together by one warp pass
  constexpr int NACC = ScalarTraits<T>::is_complex ? 2 : 1;
  __shared__ T sv[MAXNB], sw[MAXNB], p1[MAXNB], p2[MAXNB];
  __shared__ double sh[2 * CPB * 32 + 2 * CPB];
  __shared__ double spart[GEMV_CHUNK * 16];   // per column: 8 warp partials (re), 8 (im)
  const int tid = threadIdx.x, nblk = gridDim.x;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + tid, gsize = (int64_t)nblk * blockDim.x;
  const int64_t n = a.n, lda = a.lda, ldp = a.ldp;
  const T one = from_complex<T>(1.0, 0.0);
  const bool vec_ok = (lda % 2 == 0) && (ldp % 2 == 0) && (((uintptr_t)a.A & 15) == 0) && (((uintptr_t)a.Wp & 15) == 0) &&
                      (((uintptr_t)a.Vp & 15) == 0);
  T tau_prev = zero_<T>();
  for (int i = 0; i <= a.w; ++i) {   // i == w: only finishes the last w
    const int64_t j = a.p + i;
    const int ip = i - 1;
    // ---------------- phase AD ----------------
    T alpha = zero_<T>();
    if (i > 0) {
      const int64_t mprev = n - j;   // length of y of column j - 1 (rows j .. n-1)
      for (int k = tid; k < ip; k += blockDim.x) { p1[k] = a.y[mprev + k]; p2[k] = a.y[mprev + ip + k]; }
      double s2[2] = {0.0, 0.0};
      for (int k = tid; k < nblk; k += blockDim.x) { s2[0] += a.part2[2 * k]; s2[1] += a.part2[2 * k + 1]; }
      blk_sum<2>(s2, sh);            // (its barriers also publish p1 / p2)
      double cross = 0.0;
      for (int k = 0; k < ip; ++k) cross += re(p1[k]) * re(p2[k]) + im(p1[k]) * im(p2[k]);
      const T w0hv = mul_(conj_(tau_prev), from_complex<T>(s2[0] - 2.0 * cross, s2[1]));
      alpha = mul_(tau_prev, from_complex<T>(-0.5 * re(w0hv), -0.5 * im(w0hv)));
    }
    if (i < a.w) {
      for (int k = tid; k < ip; k += blockDim.x) { sv[k] = conj_(a.Vp[j + (int64_t)k * ldp]); sw[k] = conj_(a.Wp[j + (int64_t)k * ldp]); }
      if (i > 0 && tid < 32) {       // row j of the column being finished: V[j, ip] = 1, W[j, ip] = w_{ip}[j]
        double ar = 0.0, ai = 0.0;
        for (int k = tid; k < ip; k += 32) {
          T t = mul_(a.Vp[j + (int64_t)k * ldp], p1[k]);
          fma_(t, a.Wp[j + (int64_t)k * ldp], p2[k]);
          ar += re(t); ai += im(t);
        }
        for (int o = 16; o > 0; o >>= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, o); ai += __shfl_xor_sync(0xffffffffu, ai, o); }
        if (tid == 0) {
          T wj = mul_(tau_prev, sub_(a.y[0], from_complex<T>(ar, ai)));
          wj = add_(wj, alpha);
          sv[ip] = one;
          sw[ip] = conj_(wj);
        }
      }
    }
    __syncthreads();
    double sig[1] = {0.0};
    // eight threads per row share the k loop (the panel columns), so that all of the grid's threads take part
    {
      const int kp = tid & 7;
      const int64_t rows_per_pass = gsize >> 3;
      for (int64_t rb = j + ((gtid - (tid & 31)) >> 3); rb < n; rb += rows_per_pass) {   // rb is warp-uniform
        const int64_t r = rb + ((tid & 31) >> 3);
        const bool valid = r < n;
        T accw = zero_<T>(), accu = zero_<T>();
        if (valid) {
          for (int k = kp; k < ip; k += 8) {
            const T vk = a.Vp[r + (int64_t)k * ldp], wk = a.Wp[r + (int64_t)k * ldp];
            fma_(accw, vk, p1[k]);
            fma_(accw, wk, p2[k]);
            if (i < a.w) { fma_(accu, vk, sw[k]); fma_(accu, wk, sv[k]); }
          }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { accw = add_(accw, shfl_xor_T(accw, o)); accu = add_(accu, shfl_xor_T(accu, o)); }
        if (valid && kp == 0) {
          if (i > 0) {
            const T vip = a.Vp[r + (int64_t)ip * ldp];
            T wr = mul_(tau_prev, sub_(a.y[r - j], accw));
            fma_(wr, alpha, vip);
            a.Wp[r + (int64_t)ip * ldp] = wr;
            if (i < a.w) { fma_(accu, vip, sw[ip]); fma_(accu, wr, sv[ip]); }
          }
          if (i < a.w) {
            const T av = sub_(a.A[r + j * lda], accu);
            a.A[r + j * lda] = av;
            if (r == j) a.d[j] = re(av);
            if (r >= j + 2) sig[0] += abs2_(av);
          }
        }
      }
    }
    if (i == a.w) break;
    blk_sum<1>(sig, sh);
    if (tid == 0) a.part[blockIdx.x] = sig[0];
    grid.sync();
    // ---------------- phase C ----------------
    const int64_t m = n - j - 1;
    double s1[1] = {0.0};
    for (int k = tid; k < nblk; k += blockDim.x) s1[0] += a.part[k];
    blk_sum<1>(s1, sh);
    const double sigma = s1[0];
    const T* xcol = a.A + (j + 1) + j * lda;
    const T alpha0 = xcol[0];
    const double ar = re(alpha0), ai = im(alpha0);
    T tau = zero_<T>(), scale = zero_<T>();
    double beta = ar;
    if (!(sigma == 0.0 && ai == 0.0)) {
      beta = -copysign(sqrt(ar * ar + ai * ai + sigma), ar);
      const double dr = ar - beta, di = ai, den = dr * dr + di * di;
      scale = from_complex<T>(dr / den, -di / den);
      tau = from_complex<T>((beta - ar) / beta, -ai / beta);
    }
    tau_prev = tau;
    if (gtid == 0) { a.taus[j] = tau; a.e[j] = beta; }
    T* vcol = a.Vp + (int64_t)i * ldp + (j + 1);
    for (int64_t rr = gtid; rr < m; rr += gsize) vcol[rr] = (rr == 0) ? one : mul_(scale, xcol[rr]);
    // column dot products: every CTA owns a contiguous range of the ncol columns; its 8 warps stream disjoint row
    // classes of CPB columns at a time (no block barrier inside the stream), partials meet in shared memory
    const int64_t ncol = m + 2 * (int64_t)i;
    const T* At = a.A + (j + 1) + (j + 1) * lda;
    const T* Wr = a.Wp + (j + 1);
    const T* Vr = a.Vp + (j + 1);
    const int64_t per = (ncol + nblk - 1) / nblk;
    const int64_t cbeg = imin64(ncol, (int64_t)blockIdx.x * per), cend = imin64(ncol, cbeg + per);
    const int warp = tid >> 5, lane = tid & 31;
    double yhv[2] = {0.0, 0.0};
    for (int64_t cb = cbeg; cb < cend; cb += GEMV_CHUNK) {   // (one pass unless a CTA owns more than GEMV_CHUNK columns)
      const int64_t ce = imin64(cend, cb + GEMV_CHUNK);
      for (int64_t c0 = cb; c0 < ce;) {
        // a set of up to CSET consecutive columns of one of the three matrices: base + q * stride
        const T* base;
        int64_t stride, segend;
        if (c0 < m) { base = At + c0 * lda; stride = lda; segend = m; }
        else if (c0 < m + i) { base = Wr + (c0 - m) * ldp; stride = ldp; segend = m + i; }
        else { base = Vr + (c0 - m - i) * ldp; stride = ldp; segend = ncol; }
        const int nset = (int)imin64(CSET, imin64(segend, ce) - c0);
        double acc[CSET * NACC];
#pragma unroll
        for (int q = 0; q < CSET * NACC; ++q) acc[q] = 0.0;
        bool done = false;
        if constexpr (!ScalarTraits<T>::is_complex) {
          if (vec_ok) {   // 128-bit loads: all columns (and x) share the 16-byte phase of row j + 1
            const int64_t start = (j + 1) & 1;
            const int64_t npairs = (m - start) >> 1;
            const double sc = re(scale);
            if (tid == 0 && start == 1) {   // unaligned head row rr = 0 (x = 1)
#pragma unroll
              for (int q = 0; q < CSET; ++q) if (q < nset) acc[q] += re(base[q * stride]);
            }
            if (tid == 32 && ((m - start) & 1)) {   // odd tail row
              const int64_t rr = m - 1;
              const double x = (rr == 0) ? 1.0 : sc * re(xcol[rr]);
#pragma unroll
              for (int q = 0; q < CSET; ++q) if (q < nset) acc[q] += re(base[q * stride + rr]) * x;
            }
            for (int64_t pi = tid; pi < npairs; pi += blockDim.x) {
              const int64_t rr = start + 2 * pi;
              const double2 xv = *reinterpret_cast<const double2*>(xcol + rr);
              const double x0 = (rr == 0) ? 1.0 : sc * xv.x, x1 = sc * xv.y;
              double2 v[CSET];
#pragma unroll
              for (int q = 0; q < CSET; ++q)
                v[q] = (q < nset) ? *reinterpret_cast<const double2*>(base + q * stride + rr) : make_double2(0.0, 0.0);
#pragma unroll
              for (int q = 0; q < CSET; ++q) acc[q] += v[q].x * x0 + v[q].y * x1;
            }
            done = true;
          }
        }
        if (!done) {
          for (int64_t rr = tid; rr < m; rr += blockDim.x) {
            const T x = (rr == 0) ? one : mul_(scale, xcol[rr]);
            T v[CSET];
#pragma unroll
            for (int q = 0; q < CSET; ++q) v[q] = (q < nset) ? base[q * stride + rr] : zero_<T>();
#pragma unroll
            for (int q = 0; q < CSET; ++q) {
              acc[NACC * q] += re(v[q]) * re(x) + im(v[q]) * im(x);
              if (NACC == 2) acc[NACC * q + NACC - 1] += re(v[q]) * im(x) - im(v[q]) * re(x);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < CSET * NACC; ++q)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < CSET; ++q) {
            if (q < nset) {
              const int64_t c = c0 + q;
              spart[(c - cb) * 16 + warp] = acc[NACC * q];
              spart[(c - cb) * 16 + 8 + warp] = (NACC == 2) ? acc[NACC * q + NACC - 1] : 0.0;
            }
          }
        }
        c0 += nset;
      }
      __syncthreads();
      for (int64_t c = cb + tid; c < ce; c += blockDim.x) {
        double yr = 0.0, yi = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) { yr += spart[(c - cb) * 16 + w8]; yi += spart[(c - cb) * 16 + 8 + w8]; }
        a.y[c] = from_complex<T>(yr, yi);
        if (c < m) {
          const T vc = (c == 0) ? one : mul_(scale, xcol[c]);
          yhv[0] += yr * re(vc) + yi * im(vc);       // conj(y_c) v_c
          yhv[1] += yr * im(vc) - yi * re(vc);
        }
      }
      __syncthreads();
    }
    blk_sum<2>(yhv, sh);
    const double yhv0 = yhv[0], yhv1 = yhv[1];
    if (tid == 0) { a.part2[2 * blockIdx.x] = yhv0; a.part2[2 * blockIdx.x + 1] = yhv1; }
    grid.sync();
  }
}

// ------------------------------------------------------------------------------------------------
// Symmetric variant of the panel kernel (real FP64, even n, 16-byte aligned columns): phase C reads only the lower
// triangle of the trailing matrix (plus the square diagonal blocks), half the HBM traffic of the kernel above.
//   * coordinates: u = absolute row - R0, R0 = (j + 1) rounded down to even (128-bit loads stay aligned; the extra
//     row u = 0 of an odd j + 1 enters with x = 0), mu = n - R0 rows.
//   * work unit (I, J): column block J (TC columns) x row chunk I (SYM_RC = 512 rows = one pair of rows per thread),
//     rows from the top of the diagonal block down.  Element A[u, c], u >= c, gives  y_c += A[u, c] x_u  (dot) and, for
//     u > c,  y_u += A[u, c] x_c  (the mirrored element, kept in registers); the upper triangle is never read.
//     Units are dealt round-robin to the CTAs; every unit stores its TC dot partials and its 512 row partials, and the
//     next phase AD sums, for row u, the partials of its column block (over I) and of its row chunk (over J) in a
//     fixed order (deterministic, no atomics).
//   * the 2 i panel columns (W^H v, V^H v) are units of 16 columns x 2048 rows with partials over the row chunks.
// tools/proto_trd_sym.py is the NumPy statement of the same index scheme (checked on the CPU by tests/test_cpu_dc.py).
// ------------------------------------------------------------------------------------------------
constexpr int SYM_RC = 512;

struct SymCfg { int R0, mu, nJ, nI, UA, nIW, UW, U, s, TC, q, nsetW; };   // n < 2^31

__host__ __device__ __forceinline__ int sym_sumfloor(int J, int q) {   // sum_{J' < J} floor(J' / q)
  const int a = J / q, b = J % q;
  return q * (a * (a - 1) / 2) + a * b;
}

__device__ __forceinline__ SymCfg sym_cfg(int n, int j, int i, int G, int force_tc) {
  // Unit width: the widest column block that still fills one round of the grid (measured at n = 8192: fixed 64 -> 315 ms,
  // 32 -> 330 ms, 16 -> 363 ms, 128 -> 354 ms; wide units amortise the per-unit barrier and partial stores, narrow ones
  // keep all CTAs busy once the trailing matrix is small).
  SymCfg c{};
  const int R0 = (j + 1) & ~1, mu = n - R0;
  const int tcs[3] = {64, 32, 16};
  for (int t = 0; t < 3; ++t) {
    const int TC = force_tc ? force_tc : tcs[t];
    c.R0 = R0; c.mu = mu; c.s = j + 1 - R0; c.TC = TC; c.q = SYM_RC / TC;
    c.nJ = (mu + TC - 1) / TC; c.nI = (mu + SYM_RC - 1) / SYM_RC;
    c.UA = c.nJ * c.nI - sym_sumfloor(c.nJ, c.q);
    c.nIW = (mu + 4 * SYM_RC - 1) / (4 * SYM_RC); c.nsetW = (i + 15) / 16; c.UW = 2 * c.nsetW * c.nIW;
    c.U = c.UA + c.UW;
    if (force_tc || c.U >= G) break;
  }
  return c;
}

struct TrdSymArgs {
  double* A; int64_t lda, n, p; int w;
  double* Vp; double* Wp; int64_t ldp;
  double* taus; double* d; double* e;
  double* part;     // gridDim.x partials of sigma
  double* part2;    // gridDim.x partials of y^T v
  double* dotP;     // [J][I][TC] dot partials
  double* zP;       // [I][J][SYM_RC] row partials
  double* pP;       // [2][MAXNB][nIW] partials of W^T v, V^T v
  int force_tc;
};

// Sum 16 per-lane values over the 32 lanes of a warp with a butterfly that halves the number of live values per
// level (8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 16 x 5): on return every lane holds the warp total of column
// sym_col_of_lane(lane).  Shuffles share the LSU data pipe with the 128-bit loads, so their count matters.
__device__ __forceinline__ int sym_col_of_lane(int lane) { return ((lane & 1) << 3) | ((lane & 2) << 1) | ((lane & 4) >> 1) | ((lane & 8) >> 3); }
__device__ __forceinline__ double warp_reduce16(double (&acc)[16], int lane) {
  const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4, b3 = lane & 8;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const double send = b0 ? acc[q] : acc[q + 8], keep = b0 ? acc[q + 8] : acc[q];
    acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double send = b1 ? acc[q] : acc[q + 4], keep = b1 ? acc[q + 4] : acc[q];
    acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const double send = b2 ? acc[q] : acc[q + 2], keep = b2 ? acc[q + 2] : acc[q];
    acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const double send = b3 ? acc[0] : acc[1], keep = b3 ? acc[1] : acc[0];
    acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  return acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 16);
}

__global__ void __launch_bounds__(256, 2) trd_panel_sym_kernel(const TrdSymArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double sv[MAXNB], sw[MAXNB], p1[MAXNB], p2[MAXNB];
  __shared__ double sh[2 * 32 + 2];
  __shared__ double spart[2 * 128 * 8];   // per column of the unit: 8 warp partials (two buffers, alternating per unit)
  __shared__ __align__(16) double sxw[8][16];   // per warp: x of the 16 columns of the current set
  const int tid = threadIdx.x, nblk = gridDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + tid, gsize = (int64_t)nblk * blockDim.x;
  const int64_t n = a.n, lda = a.lda, ldp = a.ldp;
  double tau_prev = 0.0;
  for (int i = 0; i <= a.w; ++i) {   // i == w: only finishes the last w
    const int64_t j = a.p + i;
    const int ip = i - 1;
    // ---------------- phase AD ----------------
    double alpha = 0.0;
    SymCfg cp{};
    if (i > 0) {
      cp = sym_cfg((int)n, (int)j - 1, ip, nblk, a.force_tc);
      for (int k = tid; k < ip; k += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int iw = 0; iw < cp.nIW; ++iw) { s1 += a.pP[k * cp.nIW + iw]; s2 += a.pP[(MAXNB + k) * cp.nIW + iw]; }
        p1[k] = s1; p2[k] = s2;
      }
      double s2[1] = {0.0};
      for (int k = tid; k < nblk; k += blockDim.x) s2[0] += a.part2[k];
      blk_sum<1>(s2, sh);            // (its barriers also publish p1 / p2)
      double cross = 0.0;
      for (int k = 0; k < ip; ++k) cross += p1[k] * p2[k];
      const double w0hv = tau_prev * (s2[0] - 2.0 * cross);
      alpha = tau_prev * (-0.5 * w0hv);
    }
    if (i < a.w) {
      for (int k = tid; k < ip; k += blockDim.x) { sv[k] = a.Vp[j + (int64_t)k * ldp]; sw[k] = a.Wp[j + (int64_t)k * ldp]; }
      if (i > 0 && tid < 32) {       // row j of the column being finished: V[j, ip] = 1, W[j, ip] = w_{ip}[j]
        double ar = 0.0;
        for (int k = tid; k < ip; k += 32) ar += a.Vp[j + (int64_t)k * ldp] * p1[k] + a.Wp[j + (int64_t)k * ldp] * p2[k];
        double yj = 0.0;             // y of row j = local row cp.s of the previous column: column block 0, no row partials
        for (int I = tid; I < cp.nI; I += 32) yj += a.dotP[I * cp.TC + cp.s];
        for (int o = 16; o > 0; o >>= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, o); yj += __shfl_xor_sync(0xffffffffu, yj, o); }
        if (tid == 0) {
          const double wj = tau_prev * (yj - ar) + alpha;
          sv[ip] = 1.0;
          sw[ip] = wj;
        }
      }
    }
    __syncthreads();
    double sig[1] = {0.0};
    {
      const int kp = tid & 7;
      const int64_t rows_per_pass = gsize >> 3;
      for (int64_t rb = j + ((gtid - (tid & 31)) >> 3); rb < n; rb += rows_per_pass) {   // rb is warp-uniform
        const int64_t r = rb + ((tid & 31) >> 3);
        const bool valid = r < n;
        double accw = 0.0, accu = 0.0, accy = 0.0;
        double a_rj = 0.0, vip = 0.0;   // issued early: the loads overlap the panel loops below
        if (valid && kp == 0) {
          if (i < a.w) a_rj = a.A[r + j * lda];
          if (i > 0) vip = a.Vp[r + (int64_t)ip * ldp];
        }
        if (valid) {
          for (int k = kp; k < ip; k += 8) {
            const double vk = a.Vp[r + (int64_t)k * ldp], wk = a.Wp[r + (int64_t)k * ldp];
            accw += vk * p1[k] + wk * p2[k];
            if (i < a.w) accu += vk * sw[k] + wk * sv[k];
          }
          if (i > 0) {   // y of row r from the partials of the previous column's units
            const int u = (int)r - cp.R0, Ju = u / cp.TC, Iu = u / SYM_RC;
            const int i0 = Ju / cp.q, nd = cp.nI - i0, nt = nd + Ju + 1;   // row partials of blocks 0 .. Ju (own block: strictly-lower part)
            const double* dp = a.dotP + (int64_t)(Ju * cp.nI + i0) * cp.TC + (u - Ju * cp.TC);
            const double* zp = a.zP + (int64_t)(Iu * cp.nJ) * SYM_RC + (u - Iu * SYM_RC);
            for (int t = kp; t < nt; t += 8) accy += (t < nd) ? dp[t * cp.TC] : zp[(int64_t)(t - nd) * SYM_RC];
          }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          accw += __shfl_xor_sync(0xffffffffu, accw, o);
          accu += __shfl_xor_sync(0xffffffffu, accu, o);
          accy += __shfl_xor_sync(0xffffffffu, accy, o);
        }
        if (valid && kp == 0) {
          if (i > 0) {
            const double wr = tau_prev * (accy - accw) + alpha * vip;
            a.Wp[r + (int64_t)ip * ldp] = wr;
            if (i < a.w) accu += vip * sw[ip] + wr * sv[ip];
          }
          if (i < a.w) {
            const double av = a_rj - accu;
            a.A[r + j * lda] = av;
            if (r == j) a.d[j] = av;
            if (r >= j + 2) sig[0] += av * av;
          }
        }
      }
    }
    if (i == a.w) break;
    blk_sum<1>(sig, sh);
    if (tid == 0) a.part[blockIdx.x] = sig[0];
    grid.sync();
    // ---------------- phase C ----------------
    const int64_t m = n - j - 1;
    const double* xcol = a.A + (j + 1) + j * lda;
    const double ar = xcol[0];       // issued before the reduction below (its latency hides behind the block barriers)
    double s1[1] = {0.0};
    for (int k = tid; k < nblk; k += blockDim.x) s1[0] += a.part[k];
    blk_sum<1>(s1, sh);
    const double sigma = s1[0];
    double tau = 0.0, sc = 0.0, beta = ar;
    if (sigma != 0.0) {
      beta = -copysign(sqrt(ar * ar + sigma), ar);
      sc = 1.0 / (ar - beta);
      tau = (beta - ar) / beta;
    }
    tau_prev = tau;
    if (gtid == 0) { a.taus[j] = tau; a.e[j] = beta; }
    double* vcol = a.Vp + (int64_t)i * ldp + (j + 1);
    for (int64_t rr = gtid; rr < m; rr += gsize) vcol[rr] = (rr == 0) ? 1.0 : sc * xcol[rr];
    const SymCfg c = sym_cfg((int)n, (int)j, i, nblk, a.force_tc);
    const double* colj = a.A + c.R0 + j * lda;            // x source: x_u = sc * colj[u] (1 at u = s, 0 above)
    const double* Ablk = a.A + c.R0 + (int64_t)c.R0 * lda;         // element (u, uc) at Ablk[u + uc * lda]
    const int s = c.s, TC = c.TC;
    double yhv[1] = {0.0};
    const int lq = (TC == 128) ? 2 : ((TC == 64) ? 3 : ((TC == 32) ? 4 : 5));   // q = SYM_RC / TC = 1 << lq
    auto prefix = [&](int J) { const int aa = J >> lq, bb = J & (c.q - 1); return J * c.nI - (c.q * ((aa * (aa - 1)) >> 1) + aa * bb); };
    int buf = 0;
    for (int unit = blockIdx.x; unit < c.U; unit += nblk, buf ^= 1) {
      // one block barrier per unit: the warp partials alternate between two buffers, so the writers of unit k + 2 have
      // passed the barrier of unit k + 1, which the readers of unit k reach only after reading
      double* sp = spart + buf * (128 * 8);
      if (unit < c.UW) {
        // ---- panel columns: 16 columns of W or V x 2048 rows ----
        const int t = unit / c.nIW, iw = unit - t * c.nIW;
        const int which = t / c.nsetW, set = t - which * c.nsetW;
        const int k0 = set * 16, nset = min(16, i - k0);
        const double* base = (which ? a.Vp : a.Wp) + c.R0 + (int64_t)k0 * ldp;
        double acc[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = 0.0;
        for (int it = 0; it < 4; ++it) {
          const int u = iw * (4 * SYM_RC) + 2 * (tid + 256 * it);
          if (u < c.mu) {
            const double2 xv = *reinterpret_cast<const double2*>(colj + u);
            const double x0 = (u < s) ? 0.0 : ((u == s) ? 1.0 : sc * xv.x), x1 = (u + 1 == s) ? 1.0 : sc * xv.y;
            double2 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = (q < nset) ? *reinterpret_cast<const double2*>(base + u + (int64_t)q * ldp) : make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] += v[q].x * x0 + v[q].y * x1;
          }
        }
        {
          const double tot = warp_reduce16(acc, lane);
          if (lane < 16) sp[sym_col_of_lane(lane) * 8 + warp] = tot;
        }
        __syncthreads();
        if (tid < nset) {
          double ys = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) ys += sp[tid * 8 + w8];
          a.pP[(which * MAXNB + k0 + tid) * c.nIW + iw] = ys;
        }
      } else {
        // ---- unit (I, J) of the trailing matrix ----
        const int ka = unit - c.UW;
        int lo = 0, hi = c.nJ - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (prefix(mid) <= ka) lo = mid; else hi = mid - 1;
        }
        const int J = lo, I = (J >> lq) + (ka - prefix(J));
        const int cbeg = J * TC, cend = min(cbeg + TC, c.mu);
        const int rbeg = max(I * SYM_RC, cbeg), rend = min((I + 1) * SYM_RC, c.mu);
        const int u = I * SYM_RC + 2 * tid;
        const bool active = (u >= rbeg) && (u < rend);
        const bool wactive = __any_sync(0xffffffffu, active);
        const bool wdiag = __any_sync(0xffffffffu, active && (u < cbeg + TC));   // warp touches the diagonal block: per-element masks
        double x0 = 0.0, x1 = 0.0;
        if (active) {
          const double2 xv = *reinterpret_cast<const double2*>(colj + u);
          x0 = (u < s) ? 0.0 : ((u == s) ? 1.0 : sc * xv.x);
          x1 = (u + 1 == s) ? 1.0 : sc * xv.y;
        }
        double z0 = 0.0, z1 = 0.0;
        for (int set = 0; set * 16 < TC; ++set) {
          const int c0 = cbeg + 16 * set;
          const int nset = min(16, cend - c0);
          if (nset <= 0) break;
          if (wactive) {      // x of the 16 columns of the set, staged per warp in shared memory (read back as broadcasts)
            __syncwarp();
            if (lane < 16) {
              const int uc = c0 + lane;
              sxw[warp][lane] = (lane >= nset || uc < s) ? 0.0 : ((uc == s) ? 1.0 : sc * colj[uc]);
            }
            __syncwarp();
          }
          double acc[16];
          double2 v[16];
          if (active) {
            const double* base = Ablk + u + (int64_t)c0 * lda;
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = (q < nset) ? *reinterpret_cast<const double2*>(base + (int64_t)q * lda) : make_double2(0.0, 0.0);
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = make_double2(0.0, 0.0);
          }
          if (!wdiag) {
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = v[q].x * x0 + v[q].y * x1;
            if (active) {
#pragma unroll
              for (int q = 0; q < 16; q += 2) {
                const double2 xc = *reinterpret_cast<const double2*>(&sxw[warp][q]);
                z0 += v[q].x * xc.x + v[q + 1].x * xc.y;
                z1 += v[q].y * xc.x + v[q + 1].y * xc.y;
              }
            }
          } else {
            // diagonal block: only the lower triangle is valid.  Element (u, uc) feeds the dot product of column uc for
            // u >= uc and the row partial of row u for u > uc (x0 = x1 = 0 on inactive lanes).
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int uc = c0 + q;
              const double xc = sxw[warp][q];
              acc[q] = ((u >= uc) ? v[q].x * x0 : 0.0) + ((u + 1 >= uc) ? v[q].y * x1 : 0.0);
              if (active) {
                z0 += (u > uc) ? v[q].x * xc : 0.0;
                z1 += (u + 1 > uc) ? v[q].y * xc : 0.0;
              }
            }
          }
          {
            const double tot = warp_reduce16(acc, lane);
            if (lane < 16) sp[(16 * set + sym_col_of_lane(lane)) * 8 + warp] = tot;
          }
        }
        __syncthreads();
        if (tid < TC && cbeg + tid < cend) {
          double ys = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) ys += sp[tid * 8 + w8];
          a.dotP[(int64_t)(J * c.nI + I) * TC + tid] = ys;
          const int uc = cbeg + tid;
          yhv[0] += ys * ((uc < s) ? 0.0 : ((uc == s) ? 1.0 : sc * colj[uc]));
        }
        if (active) {
          *reinterpret_cast<double2*>(a.zP + (int64_t)(I * c.nJ + J) * SYM_RC + 2 * tid) = make_double2(z0, z1);
          yhv[0] += z0 * x0 + z1 * x1;
        }
      }
    }
    blk_sum<1>(yhv, sh);
    if (tid == 0) a.part2[blockIdx.x] = yhv[0];
    grid.sync();
  }
}

template <typename T>
__global__ void trd_last_diag_kernel(const T* __restrict__ A, int64_t lda, int64_t n, double* __restrict__ d_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) d_out[n - 1] = re(A[(n - 1) + (n - 1) * lda]);
}

// ------------------------------------------------------------------------------------------------
// stage 3 helpers: T factor of a reflector block (H_0 ... H_{w-1} = I - V T V^H), real -> T conversion
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) larft_kernel(const T* __restrict__ Sall /* V^H V per block, ld nb */, int nb, int64_t nref,
                                                    const T* __restrict__ tau_all, T* __restrict__ Tall, int reps) {
  // One CTA per reflector block b (columns b nb .. b nb + w - 1, w = min(nb, nref - b nb)).  T lives in shared memory
  // (w x w), column i of S is staged per step; thread r owns row r of T.  The result is written `reps` times side by
  // side (copy s at columns s w .. s w + w - 1, ld nb): [T T ... T] is the left operand of the slab-summing GEMM of the
  // split-K back-transformation.
  extern __shared__ __align__(16) char larft_sm[];
  const int64_t b = blockIdx.x;
  const int w = (int)((nref - b * nb) < (int64_t)nb ? (nref - b * nb) : (int64_t)nb);
  const T* S = Sall + (size_t)b * nb * nb;
  const T* tau = tau_all + b * nb;
  T* Tm = Tall + (size_t)b * nb * nb * reps;
  T* Ts = reinterpret_cast<T*>(larft_sm);
  T* sc = Ts + (size_t)w * w;
  for (int e = threadIdx.x; e < w * w; e += blockDim.x) Ts[e] = zero_<T>();
  __syncthreads();
  for (int i = 0; i < w; ++i) {
    for (int k = threadIdx.x; k < i; k += blockDim.x) sc[k] = S[k + (size_t)i * nb];
    __syncthreads();
    const T ti = tau[i];
    for (int r = threadIdx.x; r <= i; r += blockDim.x) {
      if (r < i) {
        T acc = zero_<T>();
        for (int k = r; k < i; ++k) fma_(acc, Ts[r + k * w], sc[k]);
        Ts[r + i * w] = mul_(neg_(ti), acc);
      } else {
        Ts[i + i * w] = ti;
      }
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < w * w * reps; e += blockDim.x) {
    const int s = e / (w * w), rc = e - s * (w * w), r = rc % w, c = rc / w;
    Tm[r + ((size_t)s * w + c) * nb] = Ts[rc];
  }
}

template <typename T>
__global__ void real_to_T_kernel(const double* __restrict__ in, int64_t ldi, T* __restrict__ out, int64_t ldo, int64_t rows,
                                 int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e % rows, c = e / rows;
    out[r + c * ldo] = from_complex<T>(in[r + c * ldi], 0.0);
  }
}

// ------------------------------------------------------------------------------------------------
// stage 2: divide & conquer kernels (FP64)
// ------------------------------------------------------------------------------------------------
constexpr int LEAF = 128;

__global__ void __launch_bounds__(256) dc_leaf_build_kernel(const double* __restrict__ dmod, const double* __restrict__ e,
                                                            const int64_t* __restrict__ bounds, double* __restrict__ S) {
  const int64_t lo = bounds[blockIdx.x], hi = bounds[blockIdx.x + 1];
  const int s = (int)(hi - lo);
  double* Sl = S + (size_t)blockIdx.x * LEAF * LEAF;
  for (int idx = threadIdx.x; idx < LEAF * LEAF; idx += blockDim.x) {
    const int r = idx % LEAF, c = idx / LEAF;
    double v = 0.0;
    if (r < s && c < s) {
      if (r == c) v = dmod[lo + r];
      else if (r == c + 1) v = e[lo + c];
      else if (c == r + 1) v = e[lo + r];
    }
    Sl[idx] = v;
  }
}

__global__ void __launch_bounds__(256) dc_leaf_scatter_kernel(const double* __restrict__ R, const double* __restrict__ ev,
                                                              const int64_t* __restrict__ bounds, double* __restrict__ Z,
                                                              int64_t ldz, double* __restrict__ D) {
  const int64_t lo = bounds[blockIdx.x], hi = bounds[blockIdx.x + 1];
  const int s = (int)(hi - lo);
  const double* Rl = R + (size_t)blockIdx.x * LEAF * LEAF;
  for (int idx = threadIdx.x; idx < s * s; idx += blockDim.x) {
    const int r = idx % s, c = idx / s;
    Z[(lo + r) + (lo + c) * ldz] = Rl[r + c * LEAF];
  }
  for (int c = threadIdx.x; c < s; c += blockDim.x) D[lo + c] = ev[(size_t)blockIdx.x * LEAF + c];
}

__global__ void dc_zrows_kernel(const double* __restrict__ Z, int64_t ldz, const int32_t* __restrict__ rowsel,
                                double* __restrict__ zout, int64_t n) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) zout[c] = Z[rowsel[c] + c * ldz];
}

// Givens rotations of the deflation step on the eigenvector columns (each thread owns one row; in order)
__global__ void __launch_bounds__(256) dc_rot_kernel(double* __restrict__ Zb, int64_t ldz, int64_t N, int nrot,
                                                     const int32_t* __restrict__ rp, const int32_t* __restrict__ rn,
                                                     const double* __restrict__ rc, const double* __restrict__ rs) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  for (int t = 0; t < nrot; ++t) {
    const int64_t p = rp[t], q = rn[t];
    const double c = rc[t], s = rs[t];
    const double x = Zb[r + p * ldz], y = Zb[r + q * ldz];
    Zb[r + p * ldz] = c * x + s * y;
    Zb[r + q * ldz] = c * y - s * x;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dc::secular_root (dc_secular.h) with the sums over the poles split across the 32 lanes of a warp; the scalar
// iteration (bracket, two-pole rational step, bisection safeguard, stopping rule) is identical and executed
// redundantly by every lane (the butterfly sums are bitwise equal in all lanes).
__device__ double secular_root_warp(int K, int i, const double* __restrict__ d, const double* __restrict__ z2, double rho,
                                    double* __restrict__ delta, int64_t stride) {
  const int lane = threadIdx.x & 31;
  if (K == 1) {
    const double tau = rho * z2[0];
    if (lane == 0) delta[0] = -tau;
    return d[0] + tau;
  }
  const bool last = (i == K - 1);
  int org;
  double lo, hi, tau;
  bool done = false;
  if (last) {
    org = K - 1;
    double sz = 0.0;
    for (int j = lane; j < K; j += 32) sz += z2[j];
    sz = warp_sum(sz);
    lo = 0.0;
    hi = rho * sz;
    tau = 0.5 * hi;
  } else {
    const double gap = d[i + 1] - d[i], half = 0.5 * gap, di = d[i];
    double f = 0.0;
    for (int j = lane; j < K; j += 32) f += z2[j] / ((d[j] - di) - half);
    f = 1.0 + rho * warp_sum(f);
    if (f > 0.0) { org = i; lo = 0.0; hi = half; }
    else { org = i + 1; lo = -half; hi = 0.0; }
    tau = 0.5 * (lo + hi);
    if (f == 0.0) { tau = -half; done = true; }
  }
  const double dorg = d[org];
  const int ip = last ? K - 2 : i;
  for (int it = 0; !done && it < 100; ++it) {
    double psi = 0.0, phi = 0.0, dpsi = 0.0, dphi = 0.0;
    for (int j = lane; j < K; j += 32) {
      const double dl = (d[j] - dorg) - tau, t = z2[j] / dl, t2 = t / dl;
      if (j <= ip) { psi += t; dpsi += t2; } else { phi += t; dphi += t2; }
    }
    psi = rho * warp_sum(psi); phi = rho * warp_sum(phi); dpsi = rho * warp_sum(dpsi); dphi = rho * warp_sum(dphi);
    const double g = 1.0 + psi + phi;
    const double erretm = 1.0 + fabs(psi) + fabs(phi);
    if (fabs(g) <= 4.0 * dc::DC_EPS * erretm) break;
    if (g > 0.0) hi = tau; else lo = tau;
    const double dA = (d[ip] - dorg) - tau, dB = (d[ip + 1] - dorg) - tau;
    const double S = dpsi * dA * dA, s_ = psi - dpsi * dA;
    const double R = dphi * dB * dB, r_ = phi - dphi * dB;
    const double c0 = 1.0 + s_ + r_;
    const double a = c0;
    const double b = -(c0 * (dA + dB) + S + R);
    const double cc = c0 * dA * dB + S * dB + R * dA;
    double tn = 0.0;
    bool ok = false;
    if (a == 0.0) {
      if (b != 0.0) { tn = tau - cc / b; ok = (tn > lo && tn < hi); }
    } else {
      const double disc = b * b - 4.0 * a * cc;
      if (disc >= 0.0) {
        const double sq = sqrt(disc);
        const double q = -0.5 * (b + (b >= 0.0 ? sq : -sq));
        if (q != 0.0) { tn = tau + cc / q; ok = (tn > lo && tn < hi); }
        if (!ok) { tn = tau + q / a; ok = (tn > lo && tn < hi); }
      }
    }
    if (!ok) tn = 0.5 * (lo + hi);
    const double width = hi - lo, big = fmax(fabs(lo), fabs(hi));
    const bool stuck = (tn == tau) || (width <= 2.0 * dc::DC_EPS * big);
    tau = tn;
    if (stuck) break;
  }
  for (int j = lane; j < K; j += 32) delta[(int64_t)j * stride] = (d[j] - dorg) - tau;
  return dorg + tau;
}

// Dt[i + j ldt] = d_j - lambda_i.  Small problems: one thread per root (the shared host/device function);
// large ones: one warp per root.
__global__ void __launch_bounds__(128) dc_secular_kernel(int K, const double* __restrict__ dl, const double* __restrict__ z2,
                                                         double rho, double* __restrict__ Dt, int64_t ldt, double* __restrict__ lam) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) lam[i] = dc::secular_root(K, i, dl, z2, rho, Dt + i, ldt);
}
__global__ void __launch_bounds__(128) dc_secular_warp_kernel(int K, const double* __restrict__ dl, const double* __restrict__ z2,
                                                              double rho, double* __restrict__ Dt, int64_t ldt, double* __restrict__ lam) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // warp-uniform
  if (i >= K) return;
  const double l = secular_root_warp(K, i, dl, z2, rho, Dt + i, ldt);
  if ((threadIdx.x & 31) == 0) lam[i] = l;
}

// Gu-Eisenstat: zhat_j = sign(z_j) sqrt( -prod_i (d_j - lam_i) / prod_{i != j} (d_j - d_i) ), one CTA per j
__global__ void __launch_bounds__(128) dc_zhat_kernel(int K, const double* __restrict__ Dt, int64_t ldt, const double* __restrict__ dl,
                                                      const double* __restrict__ zz, double* __restrict__ zh) {
  __shared__ double sh[4];
  const int j = blockIdx.x;
  const double dj = dl[j];
  double pr = 1.0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const double num = Dt[i + (int64_t)j * ldt];
    pr *= (i == j) ? num : num / (dj - dl[i]);
  }
  for (int o = 16; o > 0; o >>= 1) pr *= __shfl_xor_sync(0xffffffffu, pr, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = pr;
  __syncthreads();
  if (threadIdx.x == 0) {
    double p = 1.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) p *= sh[w];
    zh[j] = copysign(sqrt(fabs(p)), zz[j]);
  }
}

// invn[i] = 1 / || zhat_j / (d_j - lam_i) ||_j : tiles of 32 roots (coalesced along i) x 8 slices of j
__global__ void __launch_bounds__(256) dc_vecnorm_kernel(int K, const double* __restrict__ Dt, int64_t ldt, const double* __restrict__ zh,
                                                         double* __restrict__ invn) {
  __shared__ double sh[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  double s = 0.0;
  if (i < K) {
    for (int j = ty; j < K; j += 8) {
      const double v = zh[j] / Dt[i + (int64_t)j * ldt];
      s += v * v;
    }
  }
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < K) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sh[q][tx];
    invn[i] = 1.0 / sqrt(t);
  }
}

__global__ void dc_vecscale_kernel(int K, double* __restrict__ Dt, int64_t ldt, const double* __restrict__ zh,
                                   const double* __restrict__ invn) {
  const int64_t total = (int64_t)K * K;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % K), j = (int)(e / K);
    const int64_t a = i + (int64_t)j * ldt;
    Dt[a] = zh[j] / Dt[a] * invn[i];   // Ut[i, j]: component j of eigenvector i
  }
}

template <typename V>
void h2d(Ctx* ctx, void* dst, const std::vector<V>& src, size_t count) {
  if (count) NSB_CUDA(cudaMemcpyAsync(dst, src.data(), sizeof(V) * count, cudaMemcpyHostToDevice, ctx->stream));
}

// T = Z diag(D) Z^T for the tridiagonal (d, e); Z (n x n, ld n) on device, D on the host (unsorted).
void dc_solve(Ctx* ctx, int64_t n, const std::vector<double>& d, const std::vector<double>& e, DevBuf& Zout,
              std::vector<double>& D, int64_t* nondeflated) {
  std::vector<int64_t> b = dc::leaf_bounds(n, LEAF);
  const int nleaf = (int)b.size() - 1;
  std::vector<double> dmod(d);
  double tnorm = 0.0;
  for (int64_t i = 0; i < n; ++i) tnorm = std::max(tnorm, std::fabs(d[i]) + (i > 0 ? std::fabs(e[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(e[i]) : 0.0));
  for (int k = 1; k < nleaf; ++k) {
    const int64_t x = b[k];
    dmod[x - 1] -= std::fabs(e[x - 1]);
    dmod[x] -= std::fabs(e[x - 1]);
  }
  const size_t nn = (size_t)n * n;
  DevBuf Z1(ctx, sizeof(double) * nn), Z2(ctx, sizeof(double) * nn);
  DevBuf dmod_d(ctx, sizeof(double) * n), e_d(ctx, sizeof(double) * std::max<int64_t>(n, 1)), b_d(ctx, sizeof(int64_t) * b.size());
  DevBuf D_d(ctx, sizeof(double) * n);
  D.assign(n, 0.0);
  {
    DevBuf S(ctx, sizeof(double) * (size_t)nleaf * LEAF * LEAF), R(ctx, sizeof(double) * (size_t)nleaf * LEAF * LEAF),
        ev(ctx, sizeof(double) * (size_t)nleaf * LEAF);
    h2d(ctx, dmod_d.ptr, dmod, n);
    h2d(ctx, e_d.ptr, e, e.size());
    h2d(ctx, b_d.ptr, b, b.size());
    dc_leaf_build_kernel<<<nleaf, 256, 0, ctx->stream>>>((const double*)dmod_d.ptr, (const double*)e_d.ptr, (const int64_t*)b_d.ptr, (double*)S.ptr);
    LAUNCH_CHECK(ctx);
    herm_eig_batch128(ctx, (const double*)S.ptr, (double*)R.ptr, (double*)ev.ptr, nleaf, 0.0, 40);
    NSB_CUDA(cudaMemsetAsync(Z1.ptr, 0, sizeof(double) * nn, ctx->stream));
    NSB_CUDA(cudaMemsetAsync(Z2.ptr, 0, sizeof(double) * nn, ctx->stream));
    dc_leaf_scatter_kernel<<<nleaf, 256, 0, ctx->stream>>>((const double*)R.ptr, (const double*)ev.ptr, (const int64_t*)b_d.ptr, (double*)Z1.ptr, n,
                                                         (double*)D_d.ptr);
    LAUNCH_CHECK(ctx);
    NSB_CUDA(cudaMemcpyAsync(D.data(), D_d.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
  }
  const bool dbg = eigh_debug();
  double tl = dbg ? now_s() : 0.0;
  if (dbg) fprintf(stderr, "[eigh/dc] %d leaves solved\n", nleaf);
  if (nondeflated) *nondeflated = 0;
  if (nleaf > 1) {
    DevBuf Dt(ctx, sizeof(double) * nn), Zg(ctx, sizeof(double) * nn);
    DevBuf z_d(ctx, sizeof(double) * n), lam_d(ctx, sizeof(double) * n), dl_d(ctx, sizeof(double) * n), z2_d(ctx, sizeof(double) * n),
        zz_d(ctx, sizeof(double) * n), zh_d(ctx, sizeof(double) * n), invn_d(ctx, sizeof(double) * n), rc_d(ctx, sizeof(double) * n),
        rs_d(ctx, sizeof(double) * n);
    DevBuf rowsel_d(ctx, sizeof(int32_t) * n), nd_d(ctx, sizeof(int32_t) * n), df_d(ctx, sizeof(int32_t) * n), rp_d(ctx, sizeof(int32_t) * n),
        rn_d(ctx, sizeof(int32_t) * n);
    std::vector<double> z(n), lam(n), h_dl(n), h_z2(n), h_zz(n), h_rc(n), h_rs(n);
    std::vector<int32_t> rowsel(n), h_nd(n), h_df(n), h_rp(n), h_rn(n);
    struct MInfo { int64_t lo, hi; int K, ndf, nrot; double rho; std::vector<double> Ddf; };
    double* Zc = (double*)Z1.ptr;
    double* Zn = (double*)Z2.ptr;
    dc::MergePlan mp;
    const double isq = 1.0 / std::sqrt(2.0);
    while (b.size() > 2) {
      const size_t nm = (b.size() - 1) / 2;
      for (size_t k = 0; k < nm; ++k) {
        const int64_t lo = b[2 * k], mid = b[2 * k + 1], hi = b[2 * k + 2];
        for (int64_t c = lo; c < mid; ++c) rowsel[c] = (int32_t)(mid - 1);
        for (int64_t c = mid; c < hi; ++c) rowsel[c] = (int32_t)mid;
      }
      h2d(ctx, rowsel_d.ptr, rowsel, n);
      dc_zrows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(Zc, n, (const int32_t*)rowsel_d.ptr, (double*)z_d.ptr, n);
      LAUNCH_CHECK(ctx);
      NSB_CUDA(cudaMemcpyAsync(z.data(), z_d.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->sync();
      std::vector<MInfo> ms(nm);
      for (size_t k = 0; k < nm; ++k) {
        const int64_t lo = b[2 * k], mid = b[2 * k + 1], hi = b[2 * k + 2], N = hi - lo;
        const double beta = e[mid - 1], sgn = beta < 0.0 ? -1.0 : 1.0;
        MInfo& m = ms[k];
        m.lo = lo; m.hi = hi; m.rho = 2.0 * std::fabs(beta);
        for (int64_t c = lo; c < mid; ++c) z[c] *= isq;
        for (int64_t c = mid; c < hi; ++c) z[c] *= sgn * isq;
        dc::plan_merge(D.data() + lo, z.data() + lo, N, m.rho, mp);
        m.K = (int)mp.nd.size(); m.ndf = (int)mp.df.size(); m.nrot = (int)mp.rot_p.size();
        for (int t = 0; t < m.K; ++t) {
          const int32_t j = mp.nd[t];
          h_nd[lo + t] = j; h_dl[lo + t] = mp.D[j]; h_zz[lo + t] = mp.z[j]; h_z2[lo + t] = mp.z[j] * mp.z[j];
        }
        m.Ddf.resize(m.ndf);
        for (int t = 0; t < m.ndf; ++t) { h_df[lo + t] = mp.df[t]; m.Ddf[t] = mp.D[mp.df[t]]; }
        for (int t = 0; t < m.nrot; ++t) { h_rp[lo + t] = mp.rot_p[t]; h_rn[lo + t] = mp.rot_n[t]; h_rc[lo + t] = mp.rot_c[t]; h_rs[lo + t] = mp.rot_s[t]; }
        if (nondeflated) *nondeflated += m.K;
      }
      h2d(ctx, nd_d.ptr, h_nd, n); h2d(ctx, df_d.ptr, h_df, n); h2d(ctx, dl_d.ptr, h_dl, n); h2d(ctx, zz_d.ptr, h_zz, n);
      h2d(ctx, z2_d.ptr, h_z2, n); h2d(ctx, rp_d.ptr, h_rp, n); h2d(ctx, rn_d.ptr, h_rn, n); h2d(ctx, rc_d.ptr, h_rc, n);
      h2d(ctx, rs_d.ptr, h_rs, n);
      for (size_t k = 0; k < nm; ++k) {
        const MInfo& m = ms[k];
        const int64_t lo = m.lo, N = m.hi - m.lo, off = lo + lo * n;
        const int K = m.K;
        double* Zb = Zc + off;
        double* Znb = Zn + off;
        if (m.nrot > 0) {
          dc_rot_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(Zb, n, N, m.nrot, (const int32_t*)rp_d.ptr + lo, (const int32_t*)rn_d.ptr + lo,
                                                                              (const double*)rc_d.ptr + lo, (const double*)rs_d.ptr + lo);
          LAUNCH_CHECK(ctx);
        }
        if (m.ndf > 0) gather_cols<double>(ctx, Zb, n, N, (const int32_t*)df_d.ptr + lo, m.ndf, nullptr, Znb + (int64_t)K * n, n);
        if (K > 0) {
          double* Dtb = (double*)Dt.ptr + off;
          double* Zgb = (double*)Zg.ptr + off;
          const double* dl = (const double*)dl_d.ptr + lo;
          gather_cols<double>(ctx, Zb, n, N, (const int32_t*)nd_d.ptr + lo, K, nullptr, Zgb, n);
          if (K <= 128) dc_secular_kernel<<<(K + 127) / 128, 128, 0, ctx->stream>>>(K, dl, (const double*)z2_d.ptr + lo, m.rho, Dtb, n, (double*)lam_d.ptr + lo);
          else dc_secular_warp_kernel<<<(K + 3) / 4, 128, 0, ctx->stream>>>(K, dl, (const double*)z2_d.ptr + lo, m.rho, Dtb, n, (double*)lam_d.ptr + lo);
          LAUNCH_CHECK(ctx);
          dc_zhat_kernel<<<K, 128, 0, ctx->stream>>>(K, Dtb, n, dl, (const double*)zz_d.ptr + lo, (double*)zh_d.ptr + lo);
          LAUNCH_CHECK(ctx);
          dc_vecnorm_kernel<<<(K + 31) / 32, 256, 0, ctx->stream>>>(K, Dtb, n, (const double*)zh_d.ptr + lo, (double*)invn_d.ptr + lo);
          LAUNCH_CHECK(ctx);
          const int grid = (int)std::min<int64_t>(((int64_t)K * K + 255) / 256, (int64_t)ctx->num_sms * 8);
          dc_vecscale_kernel<<<grid, 256, 0, ctx->stream>>>(K, Dtb, n, (const double*)zh_d.ptr + lo, (const double*)invn_d.ptr + lo);
          LAUNCH_CHECK(ctx);
          gemm<double>(ctx, OP_N, OP_T, N, K, K, 1.0, Zgb, n, 0, Dtb, n, 0, 0.0, Znb, n, 0, 1);
        }
      }
      NSB_CUDA(cudaMemcpyAsync(lam.data(), lam_d.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->sync();
      std::vector<int64_t> nbnd;
      nbnd.push_back(b[0]);
      for (size_t k = 0; k < nm; ++k) {
        const MInfo& m = ms[k];
        for (int t = 0; t < m.K; ++t) D[m.lo + t] = lam[m.lo + t];
        for (int t = 0; t < m.ndf; ++t) D[m.lo + m.K + t] = m.Ddf[t];
        nbnd.push_back(m.hi);
      }
      b.swap(nbnd);
      std::swap(Zc, Zn);
      if (dbg) {
        int64_t ksum = 0, rsum = 0;
        for (const MInfo& m : ms) { ksum += m.K; rsum += m.nrot; }
        fprintf(stderr, "[eigh/dc] level with %zu merges: %.1f ms, non-deflated %ld of %ld, rotations %ld\n", nm, (now_s() - tl) * 1e3,
                (long)ksum, (long)n, (long)rsum);
        tl = now_s();
      }
    }
    if (Zc == (double*)Z2.ptr) std::swap(Z1, Z2);
  }
  (void)tnorm;
  Zout = std::move(Z1);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// driver
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void mirror_lower_kernel(T* __restrict__ A, int64_t lda, int64_t n) {   // A[r, c] = conj(A[c, r]) for r < c
  const int64_t total = n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e % n, c = e / n;
    if (r < c) A[r + c * lda] = conj_(A[c + r * lda]);
  }
}

template <typename T>
bool Eigh<T>::reads_lower_only(const Ctx* ctx, int64_t n, int64_t lda) {
  return !ScalarTraits<T>::is_complex && ctx->opt.eigh_coop && ctx->opt.eigh_sym && n >= 4 && n % 2 == 0 && lda % 2 == 0 && n < (1ll << 30);
}

template <typename T>
void Eigh<T>::factor(Ctx* c, T* A, int64_t n_, int64_t lda, bool lower_only_input) {
  ctx = c;
  n = n_;
  NSB_REQUIRE(n >= 1, NSB_EINVAL, "eigh: empty matrix");
  const int nb = std::max(2, std::min(ctx->opt.eigh_nb, MAXNB)) & ~1;
  const T one = from_complex<T>(1.0, 0.0), mone = from_complex<T>(-1.0, 0.0);
  Vall = DevBuf(ctx, sizeof(T) * (size_t)n * n);
  taus = DevBuf(ctx, sizeof(T) * n);
  NSB_CUDA(cudaMemsetAsync(Vall.ptr, 0, sizeof(T) * (size_t)n * n, ctx->stream));
  NSB_CUDA(cudaMemsetAsync(taus.ptr, 0, sizeof(T) * n, ctx->stream));
  DevBuf d_d(ctx, sizeof(double) * n), e_d(ctx, sizeof(double) * n);
  NSB_CUDA(cudaMemsetAsync(e_d.ptr, 0, sizeof(double) * n, ctx->stream));
  DevBuf Wb(ctx, sizeof(T) * (size_t)n * nb), yb(ctx, sizeof(T) * (n + 2 * nb));
  const int maxparts = (int)((n + 255) / 256) + 1;
  DevBuf part(ctx, sizeof(double) * 2 * maxparts);
  T* Vp0 = (T*)Vall.ptr;
  T* Wp = (T*)Wb.ptr;
  T* dtau = (T*)taus.ptr;
  const int64_t nref = n - 1;
  constexpr int CPB = 4;
  const bool dbg = eigh_debug();
  double t0 = 0.0;
  if (dbg) { ctx->sync(); t0 = now_s(); }
  // cooperative panel kernel: grid sized from the occupancy query so that every CTA is resident
  int coop_grid = 0;
  DevBuf part2;
  if (ctx->opt.eigh_coop) {
    int per_sm = 0, coop_ok = 0;
    NSB_CUDA(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, ctx->device));
    NSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trd_panel_kernel<T, CPB>, 256, 0));
    per_sm = std::min(per_sm, std::max(1, ctx->opt.eigh_coop_ctas));
    if (coop_ok && per_sm >= 1) {
      coop_grid = per_sm * ctx->num_sms;
      part = DevBuf(ctx, sizeof(double) * coop_grid);
      part2 = DevBuf(ctx, sizeof(double) * 2 * coop_grid);
    }
  }
  // symmetric (half-traffic) panel kernel: real FP64, even n, 16-byte aligned columns
  bool use_sym = false;
  DevBuf dotP, zP, pP, pack;
  if constexpr (!ScalarTraits<T>::is_complex) {
    if (reads_lower_only(ctx, n, lda) && ((uintptr_t)A % 16) == 0 && ((uintptr_t)Vp0 % 16) == 0 && ((uintptr_t)Wp % 16) == 0) {
      int per_sm = 0, coop_ok = 0;
      NSB_CUDA(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, ctx->device));
      NSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trd_panel_sym_kernel, 256, 0));
      per_sm = std::min(per_sm, std::max(1, ctx->opt.eigh_coop_ctas));
      if (coop_ok && per_sm >= 1) {
        use_sym = true;
        coop_grid = per_sm * ctx->num_sms;
        part = DevBuf(ctx, sizeof(double) * coop_grid);
        part2 = DevBuf(ctx, sizeof(double) * coop_grid);
        dotP = DevBuf(ctx, sizeof(double) * (size_t)(n + 128) * (size_t)(n / SYM_RC + 2));
        zP = DevBuf(ctx, sizeof(double) * (size_t)(n / SYM_RC + 2) * (size_t)(n / 16 + 2) * SYM_RC);
        pP = DevBuf(ctx, sizeof(double) * 2 * MAXNB * (size_t)(n / (4 * SYM_RC) + 2));
        pack = DevBuf(ctx, sizeof(double) * 4 * (size_t)n * nb);   // [V W] and [W V] of the rank-2w update
      }
    }
  }
  if (lower_only_input && !use_sym) {
    mirror_lower_kernel<T><<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(A, lda, n);
    LAUNCH_CHECK(ctx);
  }
  // L2 residency (ctx option eigh_l2_persist): every column of a panel streams the whole trailing matrix once, so whatever part
  // of it stays in the 126 MB L2 between columns is HBM traffic saved.  Plain LRU keeps nothing of a matrix larger than the
  // cache; an access-policy window on the first columns of the trailing matrix (the longest ones) with hit ratio
  // (persisting capacity / bytes touched inside the window) pins that share, the rest streams past it.
  int64_t l2_persist = 0, l2_window = 0;
  if (ctx->opt.eigh_l2_persist && use_sym && (int64_t)sizeof(T) * n * n / 2 > (int64_t)(96ull << 20)) {
    int maxp = 0, maxw = 0;
    cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
    cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    if (maxp > 0 && maxw > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp) == cudaSuccess) {
      size_t got = 0;
      cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
      l2_persist = (int64_t)got;
      l2_window = maxw;
    }
    (void)cudaGetLastError();
    if (dbg) fprintf(stderr, "[eigh] L2 persistence: capacity %.1f MB, max window %.1f MB\n", l2_persist / 1048576.0, l2_window / 1048576.0);
  }
  auto set_l2_window = [&](int64_t q) {      // q: first column / row of the trailing matrix
    if (l2_persist <= 0) return;
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof(av));
    const int64_t mt = n - q;
    const int64_t col_bytes = (int64_t)sizeof(T) * lda;
    int64_t wcols = std::min<int64_t>(mt, l2_window / col_bytes);
    // bytes of the lower triangle inside the window: columns q .. q + wcols - 1, rows from the diagonal down
    const double touched = (double)sizeof(T) * ((double)wcols * (double)mt - 0.5 * (double)wcols * (double)wcols);
    if (wcols <= 0 || touched <= (double)l2_persist * 0.75 || mt * mt * (int64_t)sizeof(T) / 2 <= (int64_t)(64ull << 20)) {
      av.accessPolicyWindow.num_bytes = 0;   // the whole trailing triangle fits: leave it to the normal policy
    } else {
      av.accessPolicyWindow.base_ptr = (void*)(A + q + q * lda);
      av.accessPolicyWindow.num_bytes = (size_t)(wcols * col_bytes - (int64_t)sizeof(T) * q);
      av.accessPolicyWindow.hitRatio = (float)std::min(1.0, 0.9 * (double)l2_persist / touched);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
    (void)cudaGetLastError();
  };
  for (int64_t p = 0; p < nref; p += nb) {
    const int w = (int)std::min<int64_t>(nb, nref - p);
    T* Vp = Vp0 + p * n;   // panel columns p .. p + w - 1 of Vall (ld n)
    set_l2_window(p);
    if (use_sym) {
      if constexpr (!ScalarTraits<T>::is_complex) {
        TrdSymArgs pa;
        pa.A = A; pa.lda = lda; pa.n = n; pa.p = p; pa.w = w; pa.Vp = Vp; pa.Wp = Wp; pa.ldp = n;
        pa.taus = dtau; pa.d = (double*)d_d.ptr; pa.e = (double*)e_d.ptr;
        pa.part = (double*)part.ptr; pa.part2 = (double*)part2.ptr;
        pa.dotP = (double*)dotP.ptr; pa.zP = (double*)zP.ptr; pa.pP = (double*)pP.ptr; pa.force_tc = ctx->opt.eigh_sym_tc;
        void* kargs[] = {(void*)&pa};
        NSB_CUDA(cudaLaunchCooperativeKernel((void*)trd_panel_sym_kernel, dim3(coop_grid), dim3(256), kargs, 0, ctx->stream));
        ctx->cnt.kernel_launches++;
      }
    } else if (coop_grid > 0) {
      TrdPanelArgs<T> pa;
      pa.A = A; pa.lda = lda; pa.n = n; pa.p = p; pa.w = w; pa.Vp = Vp; pa.Wp = Wp; pa.ldp = n;
      pa.taus = dtau; pa.d = (double*)d_d.ptr; pa.e = (double*)e_d.ptr; pa.y = (T*)yb.ptr;
      pa.part = (double*)part.ptr; pa.part2 = (double*)part2.ptr;
      void* kargs[] = {(void*)&pa};
      NSB_CUDA(cudaLaunchCooperativeKernel((void*)trd_panel_kernel<T, CPB>, dim3(coop_grid), dim3(256), kargs, 0, ctx->stream));
      ctx->cnt.kernel_launches++;
    }
    for (int i = 0; coop_grid == 0 && i < w; ++i) {
      const int64_t j = p + i, m = n - j - 1;
      trd_col_update_kernel<T><<<(unsigned)((n - j + 255) / 256), 256, 0, ctx->stream>>>(A, lda, n, j, Vp, Wp, n, i, (double*)d_d.ptr);
      LAUNCH_CHECK(ctx);
      trd_house_kernel<T><<<1, 256, 0, ctx->stream>>>(A, lda, n, j, Vp + (int64_t)i * n, dtau, (double*)e_d.ptr);
      LAUNCH_CHECK(ctx);
      const int64_t ncol = m + 2 * (int64_t)i;
      trd_gemv_kernel<T, CPB><<<(unsigned)((ncol + CPB - 1) / CPB), 256, 0, ctx->stream>>>(
          A + (j + 1) + (j + 1) * lda, lda, m, Wp + (j + 1), Vp + (j + 1), n, i, Vp + (int64_t)i * n + (j + 1), (T*)yb.ptr);
      LAUNCH_CHECK(ctx);
      const int nparts = (int)((m + 255) / 256);
      trd_w1_kernel<T><<<nparts, 256, 0, ctx->stream>>>((const T*)yb.ptr, Vp + (j + 1), Wp + (j + 1), n, i, m, dtau + j, (double*)part.ptr);
      LAUNCH_CHECK(ctx);
      trd_w2_kernel<T><<<nparts, 256, 0, ctx->stream>>>(Vp + (j + 1), Wp + (j + 1), n, i, m, dtau + j, (const double*)part.ptr, nparts);
      LAUNCH_CHECK(ctx);
    }
    const int64_t q = p + w, mt = n - q;
    if (mt > 0 && use_sym) {
      // A_trail -= [V W] [W V]^H as one GEMM with K = 2 w, lower-triangular output tiles only (the symmetric panel kernel
      // never reads the upper triangle)
      T* P1 = (T*)pack.ptr;
      T* P2 = P1 + (size_t)n * 2 * nb;
      copy_block<T>(ctx, Vp + q, n, P1, mt, mt, w);
      copy_block<T>(ctx, Wp + q, n, P1 + (size_t)mt * w, mt, mt, w);
      copy_block<T>(ctx, Wp + q, n, P2, mt, mt, w);
      copy_block<T>(ctx, Vp + q, n, P2 + (size_t)mt * w, mt, mt, w);
      gemm<T>(ctx, OP_N, OP_C, mt, mt, 2 * w, mone, P1, mt, 0, P2, mt, 0, one, A + q + q * lda, lda, 0, 1, GEMM_AUTO, nullptr, GEMM_LOWER_ONLY);
    } else if (mt > 0) {   // A_trail -= V W^H + W V^H (full square: both triangles stay valid for the column dot products)
      gemm<T>(ctx, OP_N, OP_C, mt, mt, w, mone, Vp + q, n, 0, Wp + q, n, 0, one, A + q + q * lda, lda, 0, 1);
      gemm<T>(ctx, OP_N, OP_C, mt, mt, w, mone, Wp + q, n, 0, Vp + q, n, 0, one, A + q + q * lda, lda, 0, 1);
    }
  }
  if (l2_persist > 0) {
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof(av));
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
    cudaCtxResetPersistingL2Cache();
    (void)cudaGetLastError();
  }
  trd_last_diag_kernel<T><<<1, 32, 0, ctx->stream>>>(A, lda, n, (double*)d_d.ptr);
  LAUNCH_CHECK(ctx);
  std::vector<double> d(n), e(std::max<int64_t>(n - 1, 0));
  NSB_CUDA(cudaMemcpyAsync(d.data(), d_d.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (n > 1) NSB_CUDA(cudaMemcpyAsync(e.data(), e_d.ptr, sizeof(double) * (n - 1), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  Wb.release(); yb.release();
  const double t1 = dbg ? now_s() : 0.0;
  dc_solve(ctx, n, d, e, Z, w, &dc_nondeflated);
  if (dbg) {
    ctx->sync();
    fprintf(stderr, "[eigh] n=%ld nb=%d %s grid %d tridiagonalise %.1f ms, divide&conquer %.1f ms (non-deflated %ld)\n", (long)n, nb,
            use_sym ? "sym" : "full", coop_grid, (t1 - t0) * 1e3, (now_s() - t1) * 1e3, (long)dc_nondeflated);
  }
}

template <typename T>
void Eigh<T>::vectors(const int32_t* idx_host, int64_t k, T* U, int64_t ldu) {
  if (k <= 0) return;
  // block width of the back-transformation: independent of the tridiagonalisation panels (any run of consecutive
  // reflectors has a compact-WY form); wide blocks make the three GEMMs per block efficient
  const int nb = std::max(2, std::min(ctx->opt.eigh_wb, ScalarTraits<T>::is_complex ? 96 : 160)) & ~1;   // T (nb x nb) must fit in shared memory
  const T one = from_complex<T>(1.0, 0.0), mone = from_complex<T>(-1.0, 0.0), zero = zero_<T>();
  DevBuf idx(ctx, sizeof(int32_t) * k);
  NSB_CUDA(cudaMemcpyAsync(idx.ptr, idx_host, sizeof(int32_t) * k, cudaMemcpyHostToDevice, ctx->stream));
  if (ScalarTraits<T>::is_complex) {
    DevBuf Zs(ctx, sizeof(double) * (size_t)n * k);
    gather_cols<double>(ctx, (const double*)Z.ptr, n, n, (const int32_t*)idx.ptr, k, nullptr, (double*)Zs.ptr, n);
    const int grid = (int)std::min<int64_t>((n * k + 255) / 256, (int64_t)ctx->num_sms * 8);
    real_to_T_kernel<T><<<grid, 256, 0, ctx->stream>>>((const double*)Zs.ptr, n, U, ldu, n, k);
    LAUNCH_CHECK(ctx);
    ctx->sync();   // Zs is released at scope exit (stream-ordered free would also do; keep it simple)
  } else {
    gather_cols<double>(ctx, (const double*)Z.ptr, n, n, (const int32_t*)idx.ptr, k, nullptr, reinterpret_cast<double*>(U), ldu);
  }
  const int64_t nref = n - 1;
  if (nref <= 0) { ctx->sync(); return; }
  const T* Vp0 = (const T*)Vall.ptr;
  const T* dtau = (const T*)taus.ptr;
  const int64_t nblocks = (nref + nb - 1) / nb, nfull = nref / nb;
  // Y = V_b^H U has only ceil(w / 128) x ceil(k / BN) output tiles but a contraction as long as the column height:
  // split the rows into `split` slabs (one strided batch), stack the partial products and let the T-factor GEMM sum
  // them ([T T .. T] x stack), so that all SMs work on it.
  const int64_t tiles_y = ((nb + 127) / 128) * ((k + (ScalarTraits<T>::is_complex ? 63 : 127)) / (ScalarTraits<T>::is_complex ? 64 : 128));
  const int split = (int)std::max<int64_t>(1, std::min<int64_t>(std::max(1, ctx->opt.eigh_split), (int64_t)ctx->num_sms / std::max<int64_t>(tiles_y, 1)));
  DevBuf Sall(ctx, sizeof(T) * (size_t)nb * nb * nblocks), Tall(ctx, sizeof(T) * (size_t)nb * nb * split * nblocks);
  DevBuf Y(ctx, sizeof(T) * (size_t)nb * split * k), Y2(ctx, sizeof(T) * (size_t)nb * k);
  const size_t larft_smem = sizeof(T) * ((size_t)nb * nb + nb);
  {
    static bool configured[2][64] = {{false}};
    bool& c = configured[ScalarTraits<T>::is_complex ? 1 : 0][ctx->device & 63];
    if (!c) { NSB_CUDA(cudaFuncSetAttribute(larft_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(T) * (ScalarTraits<T>::is_complex ? (96 * 96 + 96) : (160 * 160 + 160))))); c = true; }
  }
  const bool dbg = eigh_debug();
  double t0 = 0.0;
  if (dbg) { ctx->sync(); t0 = now_s(); }
  // Gram matrices S_b = V_b^H V_b of all blocks in one strided batch over the full column height (the rows of V_b above
  // its first reflector are zero), then all T factors in one launch (one CTA per block).
  NSB_CUDA(cudaMemsetAsync(Sall.ptr, 0, sizeof(T) * (size_t)nb * nb * nblocks, ctx->stream));
  if (nfull > 0)
    gemm<T>(ctx, OP_C, OP_N, nb, nb, n, one, Vp0, n, (int64_t)nb * n, Vp0, n, (int64_t)nb * n, zero, (T*)Sall.ptr, nb, (int64_t)nb * nb, nfull);
  if (nblocks > nfull) {
    const int64_t p = nfull * nb, w = nref - p;
    gemm<T>(ctx, OP_C, OP_N, w, w, n - p, one, Vp0 + p * n + p, n, 0, Vp0 + p * n + p, n, 0, zero, (T*)Sall.ptr + (size_t)nfull * nb * nb, nb, 0, 1);
  }
  larft_kernel<T><<<(unsigned)nblocks, 256, larft_smem, ctx->stream>>>((const T*)Sall.ptr, nb, nref, dtau, (T*)Tall.ptr, split);
  LAUNCH_CHECK(ctx);
  for (int64_t b = nblocks - 1; b >= 0; --b) {
    const int64_t p = b * nb;
    const int w = (int)std::min<int64_t>(nb, nref - p);
    const T* Vp = Vp0 + p * n + p;   // rows p.. (row p of this panel is zero: harmless, keeps the operands 16-byte aligned)
    const int64_t mp = n - p;
    const T* Tb = (const T*)Tall.ptr + (size_t)b * nb * nb * split;
    int64_t sb = std::min<int64_t>(split, std::max<int64_t>(1, mp / 256));   // slabs of at least 256 rows
    int64_t kc = (((mp + sb - 1) / sb) + 15) / 16 * 16;
    const int64_t nf = mp / kc, rem = mp - nf * kc, st = nf + (rem > 0 ? 1 : 0);
    const int64_t ldy = st * w;
    if (nf > 0) gemm<T>(ctx, OP_C, OP_N, w, k, kc, one, Vp, n, kc, U + p, ldu, kc, zero, (T*)Y.ptr, ldy, w, nf);
    if (rem > 0) gemm<T>(ctx, OP_C, OP_N, w, k, rem, one, Vp + nf * kc, n, 0, U + p + nf * kc, ldu, 0, zero, (T*)Y.ptr + nf * w, ldy, 0, 1);
    gemm<T>(ctx, OP_N, OP_N, w, k, st * w, one, Tb, nb, 0, (const T*)Y.ptr, ldy, 0, zero, (T*)Y2.ptr, w, 0, 1);
    gemm<T>(ctx, OP_N, OP_N, mp, k, w, mone, Vp, n, 0, (const T*)Y2.ptr, w, 0, one, U + p, ldu, 0, 1);
  }
  ctx->sync();
  if (dbg) fprintf(stderr, "[eigh] n=%ld k=%ld back-transformation %.1f ms (split %d)\n", (long)n, (long)k, (now_s() - t0) * 1e3, split);
}

template struct Eigh<double>;
template struct Eigh<cdouble>;

}  // namespace nsb
