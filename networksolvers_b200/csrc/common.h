// nsb200 -- common declarations shared by all translation units of libnsb200.so
#pragma once
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <memory>
#include <complex>
#include <map>

#include "../../include/nsb200.h"

namespace nsb {

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define NSB_CUDA(x)                                                                          \
  do {                                                                                       \
    cudaError_t e__ = (x);                                                                   \
    if (e__ != cudaSuccess)                                                                  \
      throw ::nsb::Error(NSB_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e__) + " (" + \
                                        __FILE__ + ":" + std::to_string(__LINE__) + ")");   \
  } while (0)

#define NSB_REQUIRE(cond, code, msg)                                                    \
  do {                                                                                  \
    if (!(cond)) throw ::nsb::Error((code), std::string(msg) + " [" #cond "] (" +      \
                                                 __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

typedef cuDoubleComplex cdouble;

template <typename T> struct ScalarTraits;
template <> struct ScalarTraits<double> {
  static constexpr bool is_complex = false;
  static constexpr int dtype = NSB_F64;
};
template <> struct ScalarTraits<cdouble> {
  static constexpr bool is_complex = true;
  static constexpr int dtype = NSB_C128;
};

// host <-> device scalar helpers
__host__ __device__ inline double re(double x) { return x; }
__host__ __device__ inline double re(cdouble x) { return x.x; }
__host__ __device__ inline double im(double) { return 0.0; }
__host__ __device__ inline double im(cdouble x) { return x.y; }
__host__ __device__ inline double conj_(double x) { return x; }
__host__ __device__ inline cdouble conj_(cdouble x) { return make_cuDoubleComplex(x.x, -x.y); }
__host__ __device__ inline double mul_(double a, double b) { return a * b; }
__host__ __device__ inline cdouble mul_(cdouble a, cdouble b) {
  return make_cuDoubleComplex(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ inline double add_(double a, double b) { return a + b; }
__host__ __device__ inline cdouble add_(cdouble a, cdouble b) { return make_cuDoubleComplex(a.x + b.x, a.y + b.y); }
__host__ __device__ inline void fma_(double& acc, double a, double b) { acc = fma(a, b, acc); }
__host__ __device__ inline void fma_(cdouble& acc, cdouble a, cdouble b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__host__ __device__ inline double abs2_(double a) { return a * a; }
__host__ __device__ inline double abs2_(cdouble a) { return a.x * a.x + a.y * a.y; }
template <typename T> __host__ __device__ inline T zero_();
template <> __host__ __device__ inline double zero_<double>() { return 0.0; }
template <> __host__ __device__ inline cdouble zero_<cdouble>() { return make_cuDoubleComplex(0.0, 0.0); }
template <typename T> __host__ __device__ inline T from_complex(double r, double i);
template <> __host__ __device__ inline double from_complex<double>(double r, double) { return r; }
template <> __host__ __device__ inline cdouble from_complex<cdouble>(double r, double i) { return make_cuDoubleComplex(r, i); }

// ---------------------------------------------------------------------------------------------
// Context: one device, one stream, stream-ordered memory pool, counters, timers.
// ---------------------------------------------------------------------------------------------
struct Counters {
  uint64_t kernel_launches = 0;
  uint64_t gemm_calls = 0;
  double gemm_flops = 0;       // real flops issued by GEMM kernels
  uint64_t permute_bytes = 0;  // bytes moved by layout permutes (0 on the chain hot path)
  uint64_t matvecs = 0;
  uint64_t env_builds = 0;
  uint64_t qr_calls = 0;
  uint64_t svd_calls = 0;
  uint64_t jacobi_sweeps = 0;
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaMemPool_t pool = nullptr;
  int num_sms = 148;
  size_t smem_optin = 0;
  std::string last_error;
  Counters cnt;
  int gemm_impl = 0;   // 0 = auto, 1 = naive, 2 = dmma cp.async, 3 = dmma TMA
  // tuning / routing knobs of nsb_ctx_set_option: per context (two contexts of one process do not see each other's settings)
  struct Options {
    int64_t gemm_naive_max_work = 20000000;   // GEMM_AUTO: m n k batch below which the plain (non-tensor) kernel is used
    int skip_identity = 1;           // skip the identity channel of the first / last environment of an H_eff application
    int skip_identity_sharded = 1;   // the same inside the multi-GPU (sharded) application
    int merge_site_ops = 1;          // 2-site regions: apply W[a] W[b] as one small-operator pass
    int eigh_min_n = 1024;           // factorize_left takes the Gram + eigh route from this size on (<= 0: never)
    int eigh_direct_min_n = 96;      // same for the Hermitian (density-matrix) input of the expansion's `eigen`
    int eigh_nb = 64;                // panel width of the tridiagonalisation (even, <= 128)
    int eigh_coop = 1;               // tridiagonalisation panels as one cooperative kernel (2 grid barriers per column)
    int eigh_coop_ctas = 3;          // CTAs per SM of the cooperative panel kernel
    int eigh_sym = 1;                // real FP64, even n: symmetric (half-traffic) panel kernel
    int eigh_sym_tc = 0;             // its column-block width (0: chosen per column from {64, 32, 16})
    int eigh_l2_persist = 1;         // tridiagonalisation: pin part of the trailing matrix in L2 (access-policy window)
    int eigh_split = 8;              // maximum number of row slabs of the split-K product Y = V^H U (back-transformation)
    int eigh_wb = 128;               // reflectors per compact-WY block of the back-transformation
    int jacobi_block_min_n = 48;     // column count from which the blocked (GEMM-rich) Jacobi is used
    int jacobi_precondition = 1;     // QR-precondition the blocked Jacobi (Drmac-Veselic)
    int jacobi_inner_cap = 1;        // inner sweeps per pair solve
    int jacobi_pivot = 0;            // column pivoting in the preconditioning QR
    int jacobi_precondition_min_n = 1024;
    int jacobi_dsmem_spc = 4;        // tournament Jacobi: preferred maximum of slots (column pairs) per CTA while the cluster (<= 8 CTAs) has room
                                     // (measured: n = 64 1.98 -> 1.30 ms, n = 128 7.0 -> 5.1 ms against 16 slots per CTA)
    int jacobi_dsmem_min_n = 41;     // column range of the cluster / distributed-shared-memory tournament Jacobi (needs columns of
    int jacobi_dsmem_max_n = 256;    // <= 256 real / 128 complex rows; other shapes fall through to the kernels below)
    int jacobi_cluster_max_n = 112;  // column count up to which the one-sided Jacobi runs as ONE launch (0: never); measured cross-over
                                     // with the blocked (GEMM) Jacobi on B200: faster per sweep up to n ~ 120
    int sbr_staged = 0;              // experimental bulge-chasing kernel: shared-memory form of the task
    int qr_smem = 1;                 // thin QR of a matrix that fits in shared memory (with its Q) as one launch
    int qr_block_min = 64;           // min(rows, cols) from which the blocked compact-WY Householder QR is used
    int shard_envs = 1;              // multi-GPU: environments away from the current region are kept as 1 / G slabs per GPU
    int nccl_sync = 0;               // host-synchronise the stream around every collective
    int qn_block_sparse = 1;         // QN networks: block-sparse storage + grouped sector GEMMs (0: dense storage)
  } opt;
  // per-launch GEMM timing (nsb_gemm_profile_*): CUDA events recorded on the stream around every GEMM launch while enabled;
  // resolved when read.  Feeds the roofline of bench.py from the launches of the timed region itself.
  struct GemmProf { cudaEvent_t e0, e1; double flops; int64_t M, N, K, batch; };
  bool gemm_profile = false;
  std::vector<GemmProf> gemm_prof;
  // multi-GPU (optional): NCCL communicator + rank info, see shard.cu
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  // peer-memory staging windows for the fused GEMM + reduce-scatter epilogue (cudaIpc-mapped)
  void* win_local = nullptr;
  size_t win_bytes = 0;
  void* win_peer[8] = {nullptr};
  char win_peer_handle[8][64] = {{0}};
  int shard_fused = 0;
  // phase timers (ms), CUDA-event based: extract / matvec / krylov_vec / factorize / env
  double timers_ms[NSB_NUM_TIMERS] = {0};
  // small scratch areas for reductions / scalar read-back
  double* d_scratch = nullptr;   // SCRATCH_DOUBLES doubles on the device
  double* h_pinned = nullptr;    // SCRATCH_DOUBLES doubles of pinned host memory
  static constexpr int SCRATCH_DOUBLES = 16384;
  // Large blocks (matvec intermediates, factorisation workspaces) recur with exact sizes every region step: they are
  // recycled through a size-keyed cache inside the context instead of going back to the driver's stream-ordered
  // pool (single stream, so handing a freed block to a later allocation is ordered by the stream itself).
  std::multimap<size_t, void*> big_cache;
  size_t big_cached_bytes = 0;
  size_t big_cache_cap = 24ull << 30;
  static constexpr size_t BIG_MIN = 32ull << 20;
  // helper contexts on the same device (own stream, scratch and block cache): independent small factorisations -- the symmetry
  // sectors of a QN tensor -- run concurrently, one host thread per helper
  std::vector<std::unique_ptr<Ctx>> helpers;
  Ctx* helper(int i);
  ~Ctx();
  Ctx() = default;
  Ctx(const Ctx&) = delete;
  Ctx& operator=(const Ctx&) = delete;
  bool is_helper = false;
  void* alloc(size_t bytes);
  void free(void* p, size_t bytes = 0);
  void flush_big_cache();
  void sync() { NSB_CUDA(cudaStreamSynchronize(stream)); }
};

// RAII device buffer (stream-ordered)
struct DevBuf {
  Ctx* ctx = nullptr;
  void* ptr = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(Ctx* c, size_t b) : ctx(c), bytes(b) { ptr = b ? c->alloc(b) : nullptr; }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); ctx = o.ctx; ptr = o.ptr; bytes = o.bytes; o.ptr = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() { if (ptr && ctx) ctx->free(ptr, bytes); ptr = nullptr; bytes = 0; }
};

struct PhaseTimer {  // accumulates elapsed device time of a phase into ctx->timers_ms[idx]
  Ctx* ctx; int idx; cudaEvent_t e0, e1; bool active;
  PhaseTimer(Ctx* c, int i);
  ~PhaseTimer();
};
extern bool g_timers_enabled;

// Host-side wall-clock profile of named scopes (diagnostic, env NSB_HOST_PROF=1): the scope synchronises the stream on exit so
// that enqueued device work is charged to it; totals are printed to stderr when the process ends.
struct HostProf {
  Ctx* ctx; const char* name; double t0; bool active;
  HostProf(Ctx* c, const char* n);
  ~HostProf();
};

}  // namespace nsb
