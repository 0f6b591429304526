// Householder QR and one-sided Jacobi SVD on device (see linalg.h).
#include "linalg.h"
#include "ops.h"

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

template <typename T>
void qr_thin(Ctx* ctx, T* A, int64_t rows, int64_t cols, int64_t lda, T* Q, int64_t ldq, T* R, int64_t ldr) {
  int64_t k = std::min(rows, cols);
  if (k == 0) return;
  ctx->cnt.qr_calls++;
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

template <typename T>
FactorInfo factorize_left(Ctx* ctx, const T* M, int64_t rows, int64_t cols, int64_t ld, bool trans_in, double cutoff,
                          int64_t mindim, int64_t maxdim, bool sqrt_spectrum, DevBuf& U, DevBuf& C,
                          std::vector<double>& spectrum) {
  FactorInfo info;
  ctx->cnt.svd_calls++;
  const int64_t k = std::min(rows, cols);
  NSB_REQUIRE(k > 0, NSB_EINVAL, "factorize: empty matrix");
  info.decomp = (cutoff <= 1e-12) ? 1 : 2;
  maxdim = std::min<int64_t>(maxdim, k);
  const bool left = rows <= cols;   // rotate the smaller side
  const int64_t n = left ? rows : cols, m = left ? cols : rows;
  DevBuf G(ctx, sizeof(T) * m * n), V(ctx, sizeof(T) * n * n);
  if (left) {   // G = M^H (cols x rows)
    if (!trans_in) transpose_conj<T>(ctx, M, rows, cols, ld, (T*)G.ptr, m, true);
    else conj_copy_block<T>(ctx, M, ld, (T*)G.ptr, m, cols, rows);            // stored (cols x rows): conj only
  } else {      // G = M (rows x cols)
    if (!trans_in) copy_block<T>(ctx, M, ld, (T*)G.ptr, m, rows, cols);
    else transpose_conj<T>(ctx, M, cols, rows, ld, (T*)G.ptr, m, false);      // stored (cols x rows): transpose
  }
  set_identity<T>(ctx, (T*)V.ptr, n, n, n);
  info.sweeps = jacobi_onesided<T>(ctx, (T*)G.ptr, m, m, n, (T*)V.ptr, n, n);

  DevBuf norms(ctx, sizeof(double) * n);
  col_norms2<T>(ctx, (T*)G.ptr, m, n, m, (double*)norms.ptr);
  std::vector<double> P(n);
  NSB_CUDA(cudaMemcpyAsync(P.data(), norms.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return P[a] > P[b]; });
  spectrum.resize(k);
  for (int64_t i = 0; i < k; ++i) spectrum[i] = sqrt_spectrum ? std::sqrt(std::max(P[order[i]], 0.0)) : P[order[i]];
  double terr = 0.0;
  int64_t nkeep = truncate_spectrum(spectrum, cutoff, mindim, maxdim, &terr);
  info.newdim = nkeep;
  info.truncerr = terr;

  DevBuf idx(ctx, sizeof(int32_t) * nkeep), scl(ctx, sizeof(double) * nkeep);
  NSB_CUDA(cudaMemcpyAsync(idx.ptr, order.data(), sizeof(int32_t) * nkeep, cudaMemcpyHostToDevice, ctx->stream));
  U = DevBuf(ctx, sizeof(T) * rows * nkeep);
  C = DevBuf(ctx, sizeof(T) * nkeep * cols);
  if (left) {
    // U = V[:, order] (rows x nkeep);  C = (G[:, order])^H
    gather_cols<T>(ctx, (T*)V.ptr, n, rows, (int32_t*)idx.ptr, nkeep, nullptr, (T*)U.ptr, rows);
    DevBuf tmp(ctx, sizeof(T) * cols * nkeep);
    gather_cols<T>(ctx, (T*)G.ptr, m, cols, (int32_t*)idx.ptr, nkeep, nullptr, (T*)tmp.ptr, cols);
    transpose_conj<T>(ctx, (T*)tmp.ptr, cols, nkeep, cols, (T*)C.ptr, nkeep, true);
    ctx->sync();
  } else {
    // M V = W Sigma:  U = G[:, order] / sigma;  C = (V[:, order] * sigma)^H
    std::vector<double> inv(nkeep), sig(nkeep);
    for (int64_t i = 0; i < nkeep; ++i) { sig[i] = std::sqrt(std::max(P[order[i]], 0.0)); inv[i] = sig[i] > 0 ? 1.0 / sig[i] : 0.0; }
    NSB_CUDA(cudaMemcpyAsync(scl.ptr, inv.data(), sizeof(double) * nkeep, cudaMemcpyHostToDevice, ctx->stream));
    gather_cols<T>(ctx, (T*)G.ptr, m, rows, (int32_t*)idx.ptr, nkeep, (double*)scl.ptr, (T*)U.ptr, rows);
    ctx->sync();
    NSB_CUDA(cudaMemcpyAsync(scl.ptr, sig.data(), sizeof(double) * nkeep, cudaMemcpyHostToDevice, ctx->stream));
    DevBuf tmp(ctx, sizeof(T) * cols * nkeep);
    gather_cols<T>(ctx, (T*)V.ptr, n, cols, (int32_t*)idx.ptr, nkeep, (double*)scl.ptr, (T*)tmp.ptr, cols);
    transpose_conj<T>(ctx, (T*)tmp.ptr, cols, nkeep, cols, (T*)C.ptr, nkeep, true);
    ctx->sync();
  }
  return info;
}

#define INST(T)                                                                                              \
  template void qr_thin<T>(Ctx*, T*, int64_t, int64_t, int64_t, T*, int64_t, T*, int64_t);                   \
  template FactorInfo factorize_left<T>(Ctx*, const T*, int64_t, int64_t, int64_t, bool, double, int64_t, int64_t, \
                                        bool, DevBuf&, DevBuf&, std::vector<double>&);
INST(double)
INST(cdouble)

}  // namespace nsb
