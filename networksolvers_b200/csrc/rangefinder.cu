// Blocked randomised range finder on the device (src/sketched_linear_algebra/range_finder.jl:6-64 of the reference).
//
// The reference draws one probe at a time: q = linear_map(random_vector()), `north_pass` Gram-Schmidt passes against all
// previous vectors, stop at the first q whose residual norm falls below `orthogonal_threshold` (or, experimental, keep it and
// stop when it falls below `cutoff`).  Here the probes go through the map in panels (one GEMM for a matrix, a loop of H_eff
// applications for the projected operator), the panel is projected against all accepted vectors by GEMMs, and the vector
// loop inside the panel (projection coefficients, residual norm, acceptance rule) runs entirely on the device: the accept /
// stop decision is a flag in device memory and the host looks at it once per panel.  (A Cholesky-QR inside the panel would
// save launches but cannot see a residual below sqrt(eps) |y|, which is what the 1e-12 threshold of the reference tests.)
#include "rangefinder.h"

#include "gemm.h"
#include "ops.h"

namespace nsb {

#define LAUNCH_CHECK(ctx) do { (ctx)->cnt.kernel_launches++; NSB_CUDA(cudaGetLastError()); } while (0)

__device__ __forceinline__ double rf_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) rf_norm_partial_kernel(int64_t n, const T* __restrict__ x, double* __restrict__ partial) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += abs2_(x[i]);
  s = rf_warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}

// state[0] = stop flag, state[1] = accepted count; norms[k] = residual norm of accepted vector k; scale = 1 / norm or 0
__global__ void rf_guard_kernel(int nblocks, const double* __restrict__ partial, double thr, double cutoff, int32_t* __restrict__ state,
                                double* __restrict__ norms, double* __restrict__ scale) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < nblocks; ++i) t += partial[i];
  const double nrm = sqrt(t > 0.0 ? t : 0.0);
  if (state[0]) { *scale = 0.0; return; }
  if (!(nrm >= thr)) { state[0] = 1; *scale = 0.0; return; }   // residual exhausted: this vector and everything after it is dropped
  norms[state[1]] = nrm;
  state[1] += 1;
  *scale = 1.0 / nrm;
  if (nrm < cutoff) state[0] = 1;                                // experimental cutoff rule: keep this one, then stop
}

template <typename T>
__global__ void rf_scale_kernel(int64_t n, const double* __restrict__ scale, T* __restrict__ x) {
  const double s = *scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = mul_(from_complex<T>(s, 0.0), x[i]);
}

template <typename T>
int64_t range_finder_blocked(Ctx* ctx, int64_t m, int64_t ndom, const RangeMap<T>& apply_map, const T* probes, uint64_t seed,
                             int64_t max_rank, int oversample, int north_pass, double thr, double cutoff, T* Q, std::vector<double>* norms_out) {
  if (max_rank <= 0 || m <= 0 || ndom <= 0) return 0;
  max_rank = std::min(max_rank, std::min(m, ndom));
  const int64_t sketch = std::min(max_rank + (int64_t)oversample, std::min(m, ndom));
  if (sketch <= 0) return 0;
  const T one = from_complex<T>(1.0, 0.0), mone = from_complex<T>(-1.0, 0.0), zero = zero_<T>();
  const int64_t panel = 32;
  DevBuf om(ctx, probes ? 0 : sizeof(T) * (size_t)ndom * panel), coef(ctx, sizeof(T) * (size_t)std::max<int64_t>(sketch, 1) * panel);
  const int nblk = (int)std::min<int64_t>((m + 1023) / 1024, 1024);
  DevBuf part(ctx, sizeof(double) * 1024), st(ctx, sizeof(int32_t) * 2), nrm(ctx, sizeof(double) * (sketch + 1)), scl(ctx, sizeof(double));
  NSB_CUDA(cudaMemsetAsync(st.ptr, 0, sizeof(int32_t) * 2, ctx->stream));
  int64_t have = 0;
  int32_t hstate[2] = {0, 0};
  while (have < sketch) {
    const int64_t pb = std::min(panel, sketch - have);
    T* Y = Q + have * m;
    const T* Om;
    if (probes) Om = probes + have * ndom;
    else { fill_normal<T>(ctx, (T*)om.ptr, ndom * pb, seed + 7919 * (uint64_t)have, 1.0); Om = (const T*)om.ptr; }
    apply_map(Om, Y, pb);
    for (int pass = 0; pass < north_pass && have > 0; ++pass) {   // against everything accepted before this panel
      gemm<T>(ctx, OP_C, OP_N, have, pb, m, one, Q, m, 0, Y, m, 0, zero, (T*)coef.ptr, have, 0, 1);
      gemm<T>(ctx, OP_N, OP_N, m, pb, have, mone, Q, m, 0, (const T*)coef.ptr, have, 0, one, Y, m, 0, 1);
    }
    for (int64_t j = 0; j < pb; ++j) {
      T* y = Y + j * m;
      for (int pass = 0; pass < north_pass && j > 0; ++pass) {
        gemm<T>(ctx, OP_C, OP_N, j, 1, m, one, Y, m, 0, y, m, 0, zero, (T*)coef.ptr, j, 0, 1);
        gemm<T>(ctx, OP_N, OP_N, m, 1, j, mone, Y, m, 0, (const T*)coef.ptr, j, 0, one, y, m, 0, 1);
      }
      rf_norm_partial_kernel<T><<<nblk, 256, 0, ctx->stream>>>(m, y, (double*)part.ptr);
      LAUNCH_CHECK(ctx);
      rf_guard_kernel<<<1, 32, 0, ctx->stream>>>(nblk, (const double*)part.ptr, thr, cutoff, (int32_t*)st.ptr, (double*)nrm.ptr, (double*)scl.ptr);
      LAUNCH_CHECK(ctx);
      rf_scale_kernel<T><<<nblk, 256, 0, ctx->stream>>>(m, (const double*)scl.ptr, y);
      LAUNCH_CHECK(ctx);
    }
    NSB_CUDA(cudaMemcpyAsync(hstate, st.ptr, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();                                                   // the one host look per panel
    have = hstate[1];
    if (hstate[0]) break;
  }
  if (norms_out) {
    norms_out->resize(have);
    if (have > 0) NSB_CUDA(cudaMemcpyAsync(norms_out->data(), nrm.ptr, sizeof(double) * have, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
  }
  return have;
}

#define INST(T)                                                                                                            \
  template int64_t range_finder_blocked<T>(Ctx*, int64_t, int64_t, const RangeMap<T>&, const T*, uint64_t, int64_t, int, int, \
                                            double, double, T*, std::vector<double>*);
INST(double)
INST(cdouble)

}  // namespace nsb
