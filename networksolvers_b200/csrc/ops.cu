// Bandwidth-bound kernels (see ops.h).  All grids are sized in multiples of the SM count and use
// grid-stride loops; global accesses are coalesced along the fastest output mode.
#include "ops.h"

namespace nsb {

static inline int grid_for(Ctx* ctx, int64_t n, int threads, int per_sm = 8) {
  int64_t blocks = (n + threads - 1) / threads;
  int64_t cap = (int64_t)ctx->num_sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
#define LAUNCH_CHECK(ctx) do { (ctx)->cnt.kernel_launches++; NSB_CUDA(cudaGetLastError()); } while (0)

// ------------------------------------------------------------------------------------------------
// permute
// ------------------------------------------------------------------------------------------------
struct PermDesc { int rank; int64_t odims[MAX_RANK]; int64_t istride[MAX_RANK]; };

template <typename T>
__global__ void permute_kernel(const T* __restrict__ in, T* __restrict__ out, PermDesc d, int64_t total, int conj) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = idx, off = 0;
#pragma unroll 4
    for (int m = 0; m < d.rank; ++m) {
      int64_t i = rem % d.odims[m];
      rem /= d.odims[m];
      off += i * d.istride[m];
    }
    T v = in[off];
    out[idx] = conj ? conj_(v) : v;
  }
}

template <typename T>
void permute(Ctx* ctx, const T* in, T* out, int rank, const int64_t* in_dims, const int* perm, bool conj) {
  NSB_REQUIRE(rank <= MAX_RANK, NSB_EINVAL, "permute: rank too large");
  PermDesc d;
  d.rank = rank;
  int64_t istr[MAX_RANK], total = 1;
  int64_t s = 1;
  for (int m = 0; m < rank; ++m) { istr[m] = s; s *= in_dims[m]; }
  for (int m = 0; m < rank; ++m) { d.odims[m] = in_dims[perm[m]]; d.istride[m] = istr[perm[m]]; total *= in_dims[m]; }
  if (total == 0) return;
  permute_kernel<T><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(in, out, d, total, conj ? 1 : 0);
  LAUNCH_CHECK(ctx);
  ctx->cnt.permute_bytes += 2ull * total * sizeof(T);
}

// ------------------------------------------------------------------------------------------------
// small-operator apply
// ------------------------------------------------------------------------------------------------
struct BigDesc { int nbig; int64_t dims[MAX_RANK]; int64_t xs[MAX_RANK]; int64_t os[MAX_RANK]; };

template <typename T, int NCH>
__global__ void __launch_bounds__(256) small_apply_kernel(const T* __restrict__ X, T* __restrict__ out,
                                                          const T* __restrict__ W, BigDesc bd, int K,
                                                          const int64_t* __restrict__ koff, int N,
                                                          const int64_t* __restrict__ noff, int64_t total) {
  extern __shared__ __align__(16) char sm[];
  T* Ws = reinterpret_cast<T*>(sm);                         // K x NCH
  int64_t* ks = reinterpret_cast<int64_t*>(sm + sizeof(T) * (size_t)K * NCH);
  const int n0 = blockIdx.y * NCH;
  for (int i = threadIdx.x; i < K * NCH; i += blockDim.x) {
    int k = i / NCH, j = i % NCH;
    Ws[i] = (n0 + j < N) ? W[k + (int64_t)(n0 + j) * K] : zero_<T>();
  }
  for (int i = threadIdx.x; i < K; i += blockDim.x) ks[i] = koff[i];
  __syncthreads();
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = idx, xo = 0, oo = 0;
    for (int m = 0; m < bd.nbig; ++m) {
      int64_t i = rem % bd.dims[m];
      rem /= bd.dims[m];
      xo += i * bd.xs[m];
      oo += i * bd.os[m];
    }
    T acc[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc[j] = zero_<T>();
    for (int k = 0; k < K; ++k) {
      T x = X[xo + ks[k]];
#pragma unroll
      for (int j = 0; j < NCH; ++j) fma_(acc[j], Ws[k * NCH + j], x);
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j)
      if (n0 + j < N) out[oo + noff[n0 + j]] = acc[j];
  }
}

template <typename T>
void small_apply(Ctx* ctx, const T* X, T* out, const T* W, int nbig, const int64_t* big_dims,
                 const int64_t* xs_big, const int64_t* os_big, int K, const int64_t* k_off, int N,
                 const int64_t* n_off) {
  NSB_REQUIRE(nbig <= MAX_RANK, NSB_EINVAL, "small_apply: too many modes");
  BigDesc bd;
  bd.nbig = nbig;
  int64_t total = 1;
  for (int m = 0; m < nbig; ++m) { bd.dims[m] = big_dims[m]; bd.xs[m] = xs_big[m]; bd.os[m] = os_big[m]; total *= big_dims[m]; }
  if (total == 0 || N == 0) return;
  // output chunk per pass over X: one pass when N <= 32 (accumulators stay in registers), else chunks of 16
  const int NCH = (N <= 8) ? 8 : (N <= 16 ? 16 : (N <= 32 && !ScalarTraits<T>::is_complex ? 32 : 16));
  size_t smem = sizeof(T) * (size_t)K * NCH + sizeof(int64_t) * (size_t)K;
  NSB_REQUIRE(smem <= 48 * 1024, NSB_EUNSUPPORTED, "small_apply: operator too large for the small-operator path");
  int ny = (N + NCH - 1) / NCH;
  int gx = grid_for(ctx, total, 256, 4);
  dim3 grid(gx, ny);
  if (NCH == 8) small_apply_kernel<T, 8><<<grid, 256, smem, ctx->stream>>>(X, out, W, bd, K, k_off, N, n_off, total);
  else if (NCH == 16) small_apply_kernel<T, 16><<<grid, 256, smem, ctx->stream>>>(X, out, W, bd, K, k_off, N, n_off, total);
  else small_apply_kernel<T, 32><<<grid, 256, smem, ctx->stream>>>(X, out, W, bd, K, k_off, N, n_off, total);
  LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// partial[2*block + {0,1}] = (re, im) of sum conj(x) y over the block's slice
template <typename T>
__global__ void __launch_bounds__(256) dot_partial_kernel(int64_t n, const T* __restrict__ x, const T* __restrict__ y,
                                                          double* __restrict__ partial) {
  double sr = 0.0, si = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T a = x[i], b = y[i];
    sr += re(a) * re(b) + im(a) * im(b);
    si += re(a) * im(b) - im(a) * re(b);
  }
  __shared__ double shr[8], shi[8];
  sr = warp_sum(sr); si = warp_sum(si);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { shr[w] = sr; shi[w] = si; }
  __syncthreads();
  if (w == 0) {
    sr = lane < 8 ? shr[lane] : 0.0; si = lane < 8 ? shi[lane] : 0.0;
    sr = warp_sum(sr); si = warp_sum(si);
    if (lane == 0) { partial[2 * blockIdx.x] = sr; partial[2 * blockIdx.x + 1] = si; }
  }
}

__global__ void __launch_bounds__(256) dot_final_kernel(int nblocks, const double* __restrict__ partial, double* __restrict__ out) {
  double sr = 0.0, si = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { sr += partial[2 * i]; si += partial[2 * i + 1]; }
  __shared__ double shr[8], shi[8];
  sr = warp_sum(sr); si = warp_sum(si);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { shr[w] = sr; shi[w] = si; }
  __syncthreads();
  if (w == 0) {
    sr = lane < 8 ? shr[lane] : 0.0; si = lane < 8 ? shi[lane] : 0.0;
    sr = warp_sum(sr); si = warp_sum(si);
    if (lane == 0) { out[0] = sr; out[1] = si; }
  }
}

// scratch layout: [0, 4096) result slots (2 doubles each), [4096, ...) block partials
static constexpr int SLOT_DOUBLES = 4096;

template <typename T>
static void dot_async(Ctx* ctx, int64_t n, const T* x, const T* y, int slot) {
  NSB_REQUIRE(2 * slot + 1 < SLOT_DOUBLES, NSB_EINTERNAL, "dot slot out of range");
  int blocks = grid_for(ctx, n, 256, 4);
  int maxb = (Ctx::SCRATCH_DOUBLES - SLOT_DOUBLES) / 2;
  if (blocks > maxb) blocks = maxb;
  double* partial = ctx->d_scratch + SLOT_DOUBLES;
  dot_partial_kernel<T><<<blocks, 256, 0, ctx->stream>>>(n, x, y, partial);
  LAUNCH_CHECK(ctx);
  dot_final_kernel<<<1, 256, 0, ctx->stream>>>(blocks, partial, ctx->d_scratch + 2 * slot);
  LAUNCH_CHECK(ctx);
}

static void fetch_slots(Ctx* ctx, int nslots) {
  NSB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_scratch, sizeof(double) * 2 * nslots, cudaMemcpyDeviceToHost, ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
}

template <typename T>
void vec_dot_slot(Ctx* ctx, int64_t n, const T* x, const T* y, int slot) { dot_async<T>(ctx, n, x, y, slot); }
double* dot_slot_ptr(Ctx* ctx, int slot) { return ctx->d_scratch + 2 * slot; }
void dot_slots_fetch(Ctx* ctx, int nslots, double* out_host) {
  fetch_slots(ctx, nslots);
  for (int i = 0; i < 2 * nslots; ++i) out_host[i] = ctx->h_pinned[i];
}

template <typename T>
void vec_dot(Ctx* ctx, int64_t n, const T* x, const T* y, double* re_out, double* im_out) {
  dot_async<T>(ctx, n, x, y, 0);
  fetch_slots(ctx, 1);
  if (re_out) *re_out = ctx->h_pinned[0];
  if (im_out) *im_out = ctx->h_pinned[1];
}

template <typename T>
double vec_nrm2(Ctx* ctx, int64_t n, const T* x) {
  double r;
  vec_dot<T>(ctx, n, x, x, &r, nullptr);
  return sqrt(r > 0 ? r : 0.0);
}

template <typename T>
void vec_multi_dot(Ctx* ctx, int64_t n, int nvec, const T* const* xs, const T* y, T* out_host) {
  for (int i = 0; i < nvec; ++i) dot_async<T>(ctx, n, xs[i], y, i);
  fetch_slots(ctx, nvec);
  for (int i = 0; i < nvec; ++i) out_host[i] = from_complex<T>(ctx->h_pinned[2 * i], ctx->h_pinned[2 * i + 1]);
}

// ------------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void axpy_kernel(int64_t n, T a, const T* __restrict__ x, T* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T v = y[i];
    fma_(v, a, x[i]);
    y[i] = v;
  }
}
template <typename T>
__global__ void scale_kernel(int64_t n, T a, T* __restrict__ x) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = mul_(a, x[i]);
}
struct LinComb { int nvec; const void* xs[32]; double cr[32]; double ci[32]; };
template <typename T>
__global__ void lincomb_kernel(int64_t n, LinComb lc, T* __restrict__ y, int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T acc = accumulate ? y[i] : zero_<T>();
    for (int v = 0; v < lc.nvec; ++v) fma_(acc, from_complex<T>(lc.cr[v], lc.ci[v]), reinterpret_cast<const T*>(lc.xs[v])[i]);
    y[i] = acc;
  }
}

template <typename T> void vec_axpy(Ctx* ctx, int64_t n, T a, const T* x, T* y) {
  if (n == 0) return;
  axpy_kernel<T><<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, a, x, y);
  LAUNCH_CHECK(ctx);
}
template <typename T> void vec_scale(Ctx* ctx, int64_t n, T a, T* x) {
  if (n == 0) return;
  scale_kernel<T><<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, a, x);
  LAUNCH_CHECK(ctx);
}
template <typename T> void vec_copy(Ctx* ctx, int64_t n, const T* x, T* y) {
  if (n == 0 || x == y) return;
  NSB_CUDA(cudaMemcpyAsync(y, x, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx->stream));
}
template <typename T> void vec_zero(Ctx* ctx, int64_t n, T* x) {
  if (n == 0) return;
  NSB_CUDA(cudaMemsetAsync(x, 0, sizeof(T) * n, ctx->stream));
}
template <typename T> void vec_lincomb(Ctx* ctx, int64_t n, int nvec, const T* const* xs, const T* c, T* y) {
  if (n == 0) return;
  int done = 0;
  bool first = true;
  while (done < nvec || first) {
    LinComb lc;
    lc.nvec = std::min(32, nvec - done);
    for (int v = 0; v < lc.nvec; ++v) { lc.xs[v] = xs[done + v]; lc.cr[v] = re(c[done + v]); lc.ci[v] = im(c[done + v]); }
    lincomb_kernel<T><<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, lc, y, first ? 0 : 1);
    LAUNCH_CHECK(ctx);
    done += lc.nvec;
    first = false;
    if (nvec == 0) break;
  }
}

// ------------------------------------------------------------------------------------------------
// matrix helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, int64_t rows, int64_t cols, int64_t ld, T* __restrict__ out,
                                 int64_t ldo, int conj, int64_t tiles_r, int64_t tiles_c) {
  __shared__ T tile[32][33];
  for (int64_t tix = blockIdx.x; tix < tiles_r * tiles_c; tix += gridDim.x) {
    int64_t tr = tix % tiles_r, tc = tix / tiles_r;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      int64_t r = tr * 32 + threadIdx.x, c = tc * 32 + j;
      if (r < rows && c < cols) tile[j][threadIdx.x] = in[r + c * ld];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      int64_t c = tc * 32 + threadIdx.x, r = tr * 32 + j;
      if (r < rows && c < cols) {
        T v = tile[threadIdx.x][j];
        out[c + r * ldo] = conj ? conj_(v) : v;
      }
    }
    __syncthreads();
  }
}
template <typename T>
void transpose_conj(Ctx* ctx, const T* in, int64_t rows, int64_t cols, int64_t ld, T* out, int64_t ldo, bool conj) {
  if (rows == 0 || cols == 0) return;
  int64_t tr = (rows + 31) / 32, tc = (cols + 31) / 32;
  int grid = (int)std::min<int64_t>(tr * tc, (int64_t)ctx->num_sms * 8);
  transpose_kernel<T><<<grid, dim3(32, 8), 0, ctx->stream>>>(in, rows, cols, ld, out, ldo, conj ? 1 : 0, tr, tc);
  LAUNCH_CHECK(ctx);
}

template <typename T>
void copy_block(Ctx* ctx, const T* in, int64_t ldi, T* out, int64_t ldo, int64_t rows, int64_t cols) {
  if (rows == 0 || cols == 0) return;
  NSB_CUDA(cudaMemcpy2DAsync(out, ldo * sizeof(T), in, ldi * sizeof(T), rows * sizeof(T), cols, cudaMemcpyDeviceToDevice, ctx->stream));
}

template <typename T>
__global__ void gather_cols_kernel(const T* __restrict__ in, int64_t ld, int64_t rows, const int32_t* __restrict__ idx,
                                   int64_t ncols, const double* __restrict__ scale, T* __restrict__ out, int64_t ldo) {
  int64_t total = rows * ncols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i % rows, c = i / rows;
    T v = in[r + (int64_t)idx[c] * ld];
    if (scale) v = mul_(from_complex<T>(scale[c], 0.0), v);
    out[r + c * ldo] = v;
  }
}
template <typename T>
void gather_cols(Ctx* ctx, const T* in, int64_t ld, int64_t rows, const int32_t* idx_dev, int64_t ncols,
                 const double* scale_dev, T* out, int64_t ldo) {
  if (rows == 0 || ncols == 0) return;
  gather_cols_kernel<T><<<grid_for(ctx, rows * ncols, 256), 256, 0, ctx->stream>>>(in, ld, rows, idx_dev, ncols, scale_dev, out, ldo);
  LAUNCH_CHECK(ctx);
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ in, int64_t ld, const int32_t* __restrict__ idx, int64_t nrows,
                                   int64_t ncols, T* __restrict__ out, int64_t ldo) {
  int64_t total = nrows * ncols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i % nrows, c = i / nrows;
    out[r + c * ldo] = in[(int64_t)idx[r] + c * ld];
  }
}
template <typename T>
void gather_rows(Ctx* ctx, const T* in, int64_t ld, const int32_t* idx_dev, int64_t nrows, int64_t ncols, T* out, int64_t ldo) {
  if (nrows == 0 || ncols == 0) return;
  gather_rows_kernel<T><<<grid_for(ctx, nrows * ncols, 256), 256, 0, ctx->stream>>>(in, ld, idx_dev, nrows, ncols, out, ldo);
  LAUNCH_CHECK(ctx);
}

template <typename T>
__global__ void concat_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ out, int64_t pre,
                              int64_t a, int64_t b, int64_t post) {
  int64_t ab = a + b, total = pre * ab * post;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i % pre, rest = i / pre;
    int64_t j = rest % ab, q = rest / ab;
    T v;
    if (j < a) v = A[p + pre * (j + a * q)];
    else v = B ? B[p + pre * ((j - a) + b * q)] : zero_<T>();
    out[i] = v;
  }
}
// out[pre, m + 1, post]: slice `pos` of the middle mode comes from B[pre, post], the others from A[pre, m, post] in order
template <typename T>
__global__ void insert_mode_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ out, int64_t pre,
                                   int64_t m, int64_t pos, int64_t post) {
  const int64_t w = m + 1, total = pre * w * post;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i % pre, rest = i / pre;
    const int64_t j = rest % w, q = rest / w;
    out[i] = (j == pos) ? B[p + pre * q] : A[p + pre * ((j - (j > pos ? 1 : 0)) + m * q)];
  }
}
template <typename T>
void insert_mode(Ctx* ctx, const T* A, const T* B, T* out, int64_t pre, int64_t m, int64_t pos, int64_t post) {
  const int64_t total = pre * (m + 1) * post;
  if (total == 0) return;
  insert_mode_kernel<T><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(A, B, out, pre, m, pos, post);
  LAUNCH_CHECK(ctx);
}

// out[w] = max over (a, a') of |E[a, w, a'] - delta(a, a')| for E stored [n, W, n] column-major (an environment whose
// channel w is the identity has out[w] ~ eps).  One CTA per column (w, a'); the maxima meet through atomicMax on the bit
// pattern of the non-negative double (a NaN compares larger than every finite value).
template <typename T>
__global__ void __launch_bounds__(256) ident_dev_kernel(const T* __restrict__ E, int64_t n, int64_t W, unsigned long long* __restrict__ out) {
  __shared__ double red[8];
  for (int64_t col = blockIdx.x; col < W * n; col += gridDim.x) {
    const int64_t w = col % W, ap = col / W;
    double m = 0.0;
    for (int64_t a = threadIdx.x; a < n; a += blockDim.x) {
      const T v = E[a + n * col];
      const double dr = re(v) - (a == ap ? 1.0 : 0.0), di = im(v);
      const double d = fabs(dr) + fabs(di);
      if (!(d <= m)) m = d;
    }
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, m, o); if (!(t <= m)) m = t; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < (int)(blockDim.x >> 5); ++k) if (!(red[k] <= m)) m = red[k];
      atomicMax(out + w, (unsigned long long)__double_as_longlong(m));
    }
  }
}
template <typename T>
void identity_deviation(Ctx* ctx, const T* E, int64_t n, int64_t W, double* out_host) {
  NSB_REQUIRE(W >= 1 && W <= 64, NSB_EINVAL, "identity_deviation: operator link too large");
  unsigned long long* d = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
  NSB_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * W, ctx->stream));
  const int grid = (int)std::min<int64_t>(W * n, (int64_t)ctx->num_sms * 8);
  ident_dev_kernel<T><<<grid, 256, 0, ctx->stream>>>(E, n, W, d);
  LAUNCH_CHECK(ctx);
  NSB_CUDA(cudaMemcpyAsync(ctx->h_pinned, d, sizeof(double) * W, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  for (int64_t w = 0; w < W; ++w) out_host[w] = ctx->h_pinned[w];   // same bit pattern
}

template <typename T>
void concat_mode(Ctx* ctx, const T* A, const T* B, T* out, int64_t pre, int64_t a, int64_t b, int64_t post) {
  int64_t total = pre * (a + b) * post;
  if (total == 0) return;
  concat_kernel<T><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(A, B, out, pre, a, b, post);
  LAUNCH_CHECK(ctx);
}

// Philox-4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4(uint64_t ctr, uint64_t seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
__global__ void fill_normal_kernel(double* __restrict__ x, int64_t n, uint64_t seed, double scale) {
  int64_t npairs = (n + 1) / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4((uint64_t)i, seed, r);
    double u1 = ((double)r[0] * 4294967296.0 + (double)r[1] + 1.0) * (1.0 / 18446744073709551616.0);
    double u2 = ((double)r[2] * 4294967296.0 + (double)r[3] + 0.5) * (1.0 / 18446744073709551616.0);
    double rad = sqrt(-2.0 * log(u1)), s, c;
    sincospi(2.0 * u2, &s, &c);
    x[2 * i] = scale * rad * c;
    if (2 * i + 1 < n) x[2 * i + 1] = scale * rad * s;
  }
}
template <typename T>
void fill_normal(Ctx* ctx, T* x, int64_t n, uint64_t seed, double scale) {
  int64_t nd = n * (int64_t)(sizeof(T) / sizeof(double));
  if (nd == 0) return;
  fill_normal_kernel<<<grid_for(ctx, (nd + 1) / 2, 256), 256, 0, ctx->stream>>>(reinterpret_cast<double*>(x), nd, seed, scale);
  LAUNCH_CHECK(ctx);
}

template <typename T>
__global__ void set_identity_kernel(T* x, int64_t rows, int64_t cols, int64_t ld) {
  int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i % rows, c = i / rows;
    x[r + c * ld] = from_complex<T>(r == c ? 1.0 : 0.0, 0.0);
  }
}
template <typename T>
void set_identity(Ctx* ctx, T* x, int64_t rows, int64_t cols, int64_t ld) {
  if (rows * cols == 0) return;
  set_identity_kernel<T><<<grid_for(ctx, rows * cols, 256), 256, 0, ctx->stream>>>(x, rows, cols, ld);
  LAUNCH_CHECK(ctx);
}

template <typename T>
__global__ void sum_slabs_kernel(const T* __restrict__ in, int nslabs, int64_t slab, T* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slab; i += (int64_t)gridDim.x * blockDim.x) {
    T acc = in[i];
    for (int s = 1; s < nslabs; ++s) acc = add_(acc, in[(int64_t)s * slab + i]);
    out[i] = acc;
  }
}
template <typename T>
void sum_slabs(Ctx* ctx, const T* in, int nslabs, int64_t slab, T* out) {
  if (slab == 0) return;
  sum_slabs_kernel<T><<<grid_for(ctx, slab, 256), 256, 0, ctx->stream>>>(in, nslabs, slab, out);
  LAUNCH_CHECK(ctx);
}

template <typename T>
__global__ void __launch_bounds__(256) col_norms2_kernel(const T* __restrict__ A, int64_t rows, int64_t cols, int64_t ld,
                                                         double* __restrict__ out) {
  for (int64_t c = blockIdx.x; c < cols; c += gridDim.x) {
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) s += abs2_(A[r + c * ld]);
    __shared__ double sh[8];
    s = warp_sum(s);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = s;
    __syncthreads();
    if (w == 0) {
      s = lane < 8 ? sh[lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) out[c] = s;
    }
    __syncthreads();
  }
}
template <typename T>
void col_norms2(Ctx* ctx, const T* A, int64_t rows, int64_t cols, int64_t ld, double* out_dev) {
  if (cols == 0) return;
  int grid = (int)std::min<int64_t>(cols, (int64_t)ctx->num_sms * 8);
  col_norms2_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, rows, cols, ld, out_dev);
  LAUNCH_CHECK(ctx);
}

#define INST(T)                                                                                                        \
  template void permute<T>(Ctx*, const T*, T*, int, const int64_t*, const int*, bool);                                 \
  template void small_apply<T>(Ctx*, const T*, T*, const T*, int, const int64_t*, const int64_t*, const int64_t*, int, \
                               const int64_t*, int, const int64_t*);                                                   \
  template void vec_dot<T>(Ctx*, int64_t, const T*, const T*, double*, double*);                                       \
  template double vec_nrm2<T>(Ctx*, int64_t, const T*);                                                                \
  template void vec_axpy<T>(Ctx*, int64_t, T, const T*, T*);                                                           \
  template void vec_scale<T>(Ctx*, int64_t, T, T*);                                                                    \
  template void vec_copy<T>(Ctx*, int64_t, const T*, T*);                                                              \
  template void vec_zero<T>(Ctx*, int64_t, T*);                                                                        \
  template void vec_lincomb<T>(Ctx*, int64_t, int, const T* const*, const T*, T*);                                     \
  template void vec_multi_dot<T>(Ctx*, int64_t, int, const T* const*, const T*, T*);                                   \
  template void transpose_conj<T>(Ctx*, const T*, int64_t, int64_t, int64_t, T*, int64_t, bool);                       \
  template void copy_block<T>(Ctx*, const T*, int64_t, T*, int64_t, int64_t, int64_t);                                 \
  template void gather_cols<T>(Ctx*, const T*, int64_t, int64_t, const int32_t*, int64_t, const double*, T*, int64_t); \
  template void gather_rows<T>(Ctx*, const T*, int64_t, const int32_t*, int64_t, int64_t, T*, int64_t);              \
  template void concat_mode<T>(Ctx*, const T*, const T*, T*, int64_t, int64_t, int64_t, int64_t);                      \
  template void identity_deviation<T>(Ctx*, const T*, int64_t, int64_t, double*);                                      \
  template void insert_mode<T>(Ctx*, const T*, const T*, T*, int64_t, int64_t, int64_t, int64_t);                      \
  template void fill_normal<T>(Ctx*, T*, int64_t, uint64_t, double);                                                   \
  template void vec_dot_slot<T>(Ctx*, int64_t, const T*, const T*, int);                                               \
  template void set_identity<T>(Ctx*, T*, int64_t, int64_t, int64_t);                                                  \
  template void sum_slabs<T>(Ctx*, const T*, int, int64_t, T*);                                                        \
  template void col_norms2<T>(Ctx*, const T*, int64_t, int64_t, int64_t, double*);
INST(double)
INST(cdouble)

}  // namespace nsb
