// Two-stage tridiagonalisation, stage 2 on the device: band -> tridiagonal by bulge chasing (sbr_chase.h holds the
// arithmetic of one task, validated on the CPU through tests/dc_cpu_harness.cpp; tools/proto_sbr.py is the statement of the
// whole two-stage algorithm).
//
// EXPERIMENTAL (round-2 groundwork): reachable only through the test hook nsb_sbr_chase_host; the eigensolver of
// csrc/eigh.cu does not use it yet.  Checked on a B200 for the eigenvalues of the resulting tridiagonal matrix
// (tests/test_gpu_zz_experimental.py::test_experimental_bulge_chasing_kernel).
//
// One persistent cooperative kernel: CTA c runs sweeps c, c + G, c + 2 G, ... (G = grid size, all CTAs resident); step s of
// sweep j starts once sweep j - 1 has finished step min(s + 2, last) (flag done[j - 1] >= s + 3, acquire / release through
// L2).  With G CTAs the sweeps overlap with a lag of three steps: ~ min(G, n / 3 b) tasks in flight.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "sbr_chase.h"

namespace nsb {


namespace {

struct DeviceTeam {
  int tid, size;
  NSB_HD void sync() const {
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
  }
  NSB_HD double sum(double x, double* red) const {   // identical value in every thread (fixed summation order)
#ifdef __CUDA_ARCH__
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = x;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (size + 31) / 32; ++w) s += red[w];
    return s;
#else
    (void)red;
    return x;
#endif
  }
};

constexpr int SBR_MAXB = 128;

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool STAGED>
__global__ void __launch_bounds__(128) sbr_chase_kernel(sbr::Band B, double* __restrict__ V2, double* __restrict__ tau2, int64_t ldtau,
                                                        int* __restrict__ done) {
  __shared__ double v[SBR_MAXB], work[2 * SBR_MAXB], red[32];
  __shared__ double s_tau;
  extern __shared__ __align__(16) double stage[];   // STAGED: 3 b b doubles
  const DeviceTeam tm{(int)threadIdx.x, (int)blockDim.x};
  const int64_t n = B.n;
  const int b = B.b;
  for (int64_t j = blockIdx.x; j + 2 < n; j += gridDim.x) {
    const int ns = (int)sbr::nsteps(n, b, j);
    const int ns_prev = j > 0 ? (int)sbr::nsteps(n, b, j - 1) : 0;
    for (int s = 0; s < ns; ++s) {
      if (j > 0) {
        const int need = min(s + 3, ns_prev);
        if (tm.tid == 0) while (ld_acquire(done + (j - 1)) < need) __nanosleep(32);
        __syncthreads();
      }
      const int len = STAGED ? sbr::chase_task_staged(tm, B, j, s, v, &s_tau, work, red, stage) : sbr::chase_task(tm, B, j, s, v, &s_tau, work, red);
      __syncthreads();
      const int64_t r0 = j + 1 + (int64_t)s * b;
      if (len >= 2) for (int i = tm.tid; i < len; i += tm.size) V2[(r0 + i) + j * n] = v[i];
      if (tm.tid == 0) tau2[s + j * ldtau] = (len >= 2) ? s_tau : 0.0;
      __syncthreads();                      // every write of this task has been issued by its thread
      if (tm.tid == 0) { __threadfence(); st_release(done + j, s + 1); }
    }
  }
}

}  // namespace

// ab (ld x n, ld >= 2 b + 1, bulge rows zero), V2 (n x n, zero), tau2 (ldtau x n) on the device.
void sbr_chase_device(Ctx* ctx, int64_t n, int b, double* ab, int64_t ld, double* V2, double* tau2, int64_t ldtau) {
  NSB_REQUIRE(b >= 1 && b <= SBR_MAXB && ld >= 2 * (int64_t)b + 1 && n >= 1, NSB_EINVAL, "sbr_chase: bad band layout");
  if (n < 3) return;
  DevBuf done(ctx, sizeof(int) * (size_t)n);
  NSB_CUDA(cudaMemsetAsync(done.ptr, 0, sizeof(int) * (size_t)n, ctx->stream));
  int per_sm = 0, coop = 0;
  NSB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  // g_sbr_staged (ctx option "sbr_staged", default 0): the shared-memory form of the task (validated on the CPU, not yet run
  // on hardware); 0: the element-wise form (validated on a B200, slow)
  const bool staged = ctx->opt.sbr_staged != 0;
  const size_t smem = staged ? sizeof(double) * 3 * (size_t)b * b : 0;
  void* kern = staged ? (void*)sbr_chase_kernel<true> : (void*)sbr_chase_kernel<false>;
  if (staged) NSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
  NSB_REQUIRE(coop && per_sm >= 1, NSB_EUNSUPPORTED, "sbr_chase: cooperative launch unavailable");
  const int grid = (int)std::min<int64_t>(n - 2, (int64_t)std::min(per_sm, 4) * ctx->num_sms);   // all CTAs resident: no deadlock
  sbr::Band B{ab, ld, n, b};
  int* dn = (int*)done.ptr;
  void* args[] = {(void*)&B, (void*)&V2, (void*)&tau2, (void*)&ldtau, (void*)&dn};
  const bool dbg = getenv("NSB_DEBUG_EIGH") != nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (dbg) { NSB_CUDA(cudaEventCreate(&e0)); NSB_CUDA(cudaEventCreate(&e1)); NSB_CUDA(cudaEventRecord(e0, ctx->stream)); }
  NSB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(128), args, smem, ctx->stream));
  ctx->cnt.kernel_launches++;
  if (dbg) NSB_CUDA(cudaEventRecord(e1, ctx->stream));
  ctx->sync();   // `done` is released at scope exit
  if (dbg) {
    float ms = 0.f;
    NSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    fprintf(stderr, "[sbr] n=%ld b=%d grid %d %s bulge chasing %.2f ms\n", (long)n, b, grid, staged ? "staged" : "element-wise", ms);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
}

}  // namespace nsb
