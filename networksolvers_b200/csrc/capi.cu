// extern "C" boundary of libnsb200.so (declared in include/nsb200.h).  Nothing throws across it.
#include <ctime>
#include <cstdlib>
#include "nccl_dyn.h"

#include <random>
#include <cuda_profiler_api.h>

#include "net.h"
#include "eigh.h"
#include "rangefinder.h"
#include <algorithm>

namespace nsb {

bool g_timers_enabled = false;
static thread_local std::string tl_error;

void* Ctx::alloc(size_t bytes) {
  if (bytes >= BIG_MIN) {
    auto it = big_cache.find(bytes);
    if (it != big_cache.end()) {
      void* q = it->second;
      big_cache.erase(it);
      big_cached_bytes -= bytes;
      return q;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMallocFromPoolAsync(&p, bytes, pool, stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    // one retry after returning the cached blocks and letting in-flight frees complete
    flush_big_cache();
    cudaStreamSynchronize(stream);
    e = cudaMallocFromPoolAsync(&p, bytes, pool, stream);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error(NSB_ENOMEM, std::string("device allocation of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
  }
  return p;
}
void Ctx::free(void* p, size_t bytes) {
  if (!p) return;
  if (bytes >= BIG_MIN && big_cached_bytes + bytes <= big_cache_cap) {
    big_cache.emplace(bytes, p);
    big_cached_bytes += bytes;
    return;
  }
  cudaFreeAsync(p, stream);
}
Ctx* Ctx::helper(int i) {
  while ((int)helpers.size() <= i) {
    std::unique_ptr<Ctx> h(new Ctx());
    h->device = device; h->pool = pool; h->num_sms = num_sms; h->smem_optin = smem_optin; h->is_helper = true;
    h->gemm_impl = gemm_impl; h->opt = opt;
    NSB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    NSB_CUDA(cudaMalloc(&h->d_scratch, sizeof(double) * Ctx::SCRATCH_DOUBLES));
    NSB_CUDA(cudaMallocHost(&h->h_pinned, sizeof(double) * Ctx::SCRATCH_DOUBLES));
    helpers.push_back(std::move(h));
  }
  Ctx* h = helpers[i].get();
  h->opt = opt; h->gemm_impl = gemm_impl;
  return h;
}
Ctx::~Ctx() {
  if (!is_helper) return;
  flush_big_cache();
  cudaStreamSynchronize(stream);
  if (d_scratch) cudaFree(d_scratch);
  if (h_pinned) cudaFreeHost(h_pinned);
  cudaStreamDestroy(stream);
}
void Ctx::flush_big_cache() {
  for (auto& kv : big_cache) cudaFreeAsync(kv.second, stream);
  big_cache.clear();
  big_cached_bytes = 0;
}

PhaseTimer::PhaseTimer(Ctx* c, int i) : ctx(c), idx(i), e0(nullptr), e1(nullptr), active(g_timers_enabled) {
  if (!active) return;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, ctx->stream);
}
PhaseTimer::~PhaseTimer() {
  if (!active) return;
  cudaEventRecord(e1, ctx->stream);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx->timers_ms[idx] += ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

namespace {
struct HostProfTable {
  bool on;
  std::map<std::string, std::pair<uint64_t, double>> acc;
  HostProfTable() { const char* e = getenv("NSB_HOST_PROF"); on = e && e[0] == '1'; }
  ~HostProfTable() {
    if (!on) return;
    for (auto& kv : acc) fprintf(stderr, "[host-prof] %-28s calls %8llu  total %10.3f ms  avg %9.2f us\n", kv.first.c_str(),
                                 (unsigned long long)kv.second.first, kv.second.second * 1e3, kv.second.second * 1e6 / (double)kv.second.first);
  }
};
HostProfTable g_host_prof;
double wall_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }
}  // namespace
HostProf::HostProf(Ctx* c, const char* n) : ctx(c), name(n), t0(0.0), active(g_host_prof.on && !c->is_helper) {
  if (active) t0 = wall_now();
}
HostProf::~HostProf() {
  if (!active) return;
  cudaStreamSynchronize(ctx->stream);
  auto& e = g_host_prof.acc[name];
  e.first++;
  e.second += wall_now() - t0;
}

}  // namespace nsb

using namespace nsb;

struct nsb_ctx { Ctx c; int refs = 0; bool destroyed = false; cudaEvent_t ev0 = nullptr, ev1 = nullptr; };
struct nsb_net { NetBase* n; nsb_ctx* ctx; };

static int fail(Ctx* ctx, int code, const std::string& msg) {
  tl_error = msg;
  if (ctx) ctx->last_error = msg;
  return code;
}

#define NSB_TRY(ctxp) try {
#define NSB_CATCH(ctxp)                                                        \
  }                                                                            \
  catch (const nsb::Error& e) { return fail((ctxp), e.code, e.what()); }       \
  catch (const std::bad_alloc&) { return fail((ctxp), NSB_ENOMEM, "host allocation failed"); } \
  catch (const std::exception& e) { return fail((ctxp), NSB_EINTERNAL, e.what()); } \
  catch (...) { return fail((ctxp), NSB_EINTERNAL, "unknown error"); }         \
  return NSB_OK;

extern "C" {
#pragma GCC visibility push(default)

const char* nsb_version(void) { return "nsb200 0.1 (sm_100a)"; }

const char* nsb_last_error(nsb_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : tl_error.c_str(); }

int nsb_ctx_create(int device, nsb_ctx** out) {
  if (!out) return fail(nullptr, NSB_EINVAL, "nsb_ctx_create: null output");
  *out = nullptr;
  nsb_ctx* h = nullptr;
  NSB_TRY(nullptr)
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(NSB_ECUDA, "no CUDA device available: the B200 path has no CPU fallback");
  }
  NSB_REQUIRE(device >= 0 && device < ndev, NSB_EINVAL, "bad device ordinal");
  NSB_CUDA(cudaSetDevice(device));
  h = new nsb_ctx();
  Ctx& c = h->c;
  c.device = device;
  cudaDeviceProp prop;
  NSB_CUDA(cudaGetDeviceProperties(&prop, device));
  c.num_sms = prop.multiProcessorCount;
  c.smem_optin = prop.sharedMemPerBlockOptin;
  NSB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  NSB_CUDA(cudaDeviceGetDefaultMemPool(&c.pool, device));
  uint64_t thresh = UINT64_MAX;
  NSB_CUDA(cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  NSB_CUDA(cudaMalloc(&c.d_scratch, sizeof(double) * Ctx::SCRATCH_DOUBLES));
  NSB_CUDA(cudaMallocHost(&c.h_pinned, sizeof(double) * Ctx::SCRATCH_DOUBLES));
  *out = h;
  NSB_CATCH(nullptr)
}

static void ctx_really_destroy(nsb_ctx* ctx);

int nsb_ctx_destroy(nsb_ctx* ctx) {
  if (!ctx) return NSB_OK;
  if (ctx->refs > 0) { ctx->destroyed = true; return NSB_OK; }   // networks still alive: defer
  ctx_really_destroy(ctx);
  return NSB_OK;
}

static void ctx_really_destroy(nsb_ctx* ctx) {
  cudaSetDevice(ctx->c.device);
  if (ctx->c.nccl_comm) { try { nccl_api().CommDestroy((ncclComm_t)ctx->c.nccl_comm); } catch (...) {} ctx->c.nccl_comm = nullptr; }
  ctx->c.helpers.clear();
  ctx->c.flush_big_cache();
  cudaStreamSynchronize(ctx->c.stream);
  for (auto& g : ctx->c.gemm_prof) { cudaEventDestroy(g.e0); cudaEventDestroy(g.e1); }
  if (ctx->c.d_scratch) cudaFree(ctx->c.d_scratch);
  if (ctx->c.h_pinned) cudaFreeHost(ctx->c.h_pinned);
  cudaStreamDestroy(ctx->c.stream);
  delete ctx;
}

int nsb_ctx_set_option(nsb_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return fail(nullptr, NSB_EINVAL, "null argument");
  NSB_TRY(&ctx->c)
  std::string k(key);
  if (k == "gemm_impl") { NSB_REQUIRE(value >= 0 && value <= 3, NSB_EINVAL, "gemm_impl must be 0..3"); ctx->c.gemm_impl = (int)value; }
  else if (k == "gemm_naive_max_work") { NSB_REQUIRE(value >= 0, NSB_EINVAL, "gemm_naive_max_work >= 0"); ctx->c.opt.gemm_naive_max_work = value; }
  else if (k == "jacobi_block_min_n") { NSB_REQUIRE(value >= 0, NSB_EINVAL, "jacobi_block_min_n must be >= 0"); ctx->c.opt.jacobi_block_min_n = (int)value; }
  else if (k == "jacobi_precondition") { ctx->c.opt.jacobi_precondition = value != 0; }
  else if (k == "shard_fused") { ctx->c.shard_fused = value != 0; }
  else if (k == "jacobi_pivot") { ctx->c.opt.jacobi_pivot = value != 0; }
  else if (k == "jacobi_inner_cap") { NSB_REQUIRE(value >= 1, NSB_EINVAL, "jacobi_inner_cap >= 1"); ctx->c.opt.jacobi_inner_cap = (int)value; }
  else if (k == "jacobi_precondition_min_n") { ctx->c.opt.jacobi_precondition_min_n = (int)value; }
  else if (k == "jacobi_dsmem_spc") { NSB_REQUIRE(value >= 1 && value <= 16, NSB_EINVAL, "jacobi_dsmem_spc 1..16"); ctx->c.opt.jacobi_dsmem_spc = (int)value; }
  else if (k == "jacobi_dsmem_min_n") { NSB_REQUIRE(value >= 0, NSB_EINVAL, "jacobi_dsmem_min_n >= 0"); ctx->c.opt.jacobi_dsmem_min_n = (int)value; }
  else if (k == "jacobi_dsmem_max_n") { NSB_REQUIRE(value >= 0 && value <= 256, NSB_EINVAL, "jacobi_dsmem_max_n 0..256"); ctx->c.opt.jacobi_dsmem_max_n = (int)value; }
  else if (k == "jacobi_cluster_max_n") { NSB_REQUIRE(value >= 0 && value <= 4096, NSB_EINVAL, "jacobi_cluster_max_n 0..4096"); ctx->c.opt.jacobi_cluster_max_n = (int)value; }
  else if (k == "big_cache_gib") { NSB_REQUIRE(value >= 0, NSB_EINVAL, "big_cache_gib >= 0"); ctx->c.big_cache_cap = (size_t)value << 30; if (value == 0) ctx->c.flush_big_cache(); }
  else if (k == "sbr_staged") { ctx->c.opt.sbr_staged = value != 0; }
  else if (k == "skip_identity") { ctx->c.opt.skip_identity = value != 0; }
  else if (k == "skip_identity_sharded") { ctx->c.opt.skip_identity_sharded = value != 0; }
  else if (k == "merge_site_ops") { ctx->c.opt.merge_site_ops = value != 0; }
  else if (k == "eigh_min_n") { ctx->c.opt.eigh_min_n = (int)value; }
  else if (k == "eigh_coop") { ctx->c.opt.eigh_coop = value != 0; }
  else if (k == "eigh_coop_ctas") { NSB_REQUIRE(value >= 1 && value <= 8, NSB_EINVAL, "eigh_coop_ctas 1..8"); ctx->c.opt.eigh_coop_ctas = (int)value; }
  else if (k == "eigh_direct_min_n") ctx->c.opt.eigh_direct_min_n = (int)value;
  else if (k == "eigh_sym") ctx->c.opt.eigh_sym = value ? 1 : 0;
  else if (k == "eigh_l2_persist") ctx->c.opt.eigh_l2_persist = value ? 1 : 0;
  else if (k == "eigh_sym_tc") { NSB_REQUIRE(value == 0 || value == 16 || value == 32 || value == 64 || value == 128, NSB_EINVAL, "eigh_sym_tc must be 0, 16, 32, 64 or 128"); ctx->c.opt.eigh_sym_tc = (int)value; }
  else if (k == "eigh_split") { NSB_REQUIRE(value >= 1 && value <= 16, NSB_EINVAL, "eigh_split 1..16"); ctx->c.opt.eigh_split = (int)value; }
  else if (k == "eigh_wb") { NSB_REQUIRE(value >= 2 && value <= 256 && value % 2 == 0, NSB_EINVAL, "eigh_wb must be even, 2..256"); ctx->c.opt.eigh_wb = (int)value; }
  else if (k == "eigh_nb") { NSB_REQUIRE(value >= 2 && value <= 128 && value % 2 == 0, NSB_EINVAL, "eigh_nb must be even, 2..128"); ctx->c.opt.eigh_nb = (int)value; }
  else if (k == "qr_smem") { ctx->c.opt.qr_smem = value != 0; }
  else if (k == "qr_block_min") { NSB_REQUIRE(value >= 0, NSB_EINVAL, "qr_block_min >= 0"); ctx->c.opt.qr_block_min = (int)value; }
  else if (k == "qn_block_sparse") { ctx->c.opt.qn_block_sparse = value != 0; }
  else if (k == "nccl_sync") { ctx->c.opt.nccl_sync = value != 0; }
  else if (k == "shard_envs") { ctx->c.opt.shard_envs = value != 0; }
  else throw Error(NSB_EINVAL, "unknown option " + k);
  NSB_CATCH(&ctx->c)
}

int nsb_ctx_counters(nsb_ctx* ctx, nsb_counters* out) {
  if (!ctx || !out) return fail(nullptr, NSB_EINVAL, "null argument");
  const Counters& c = ctx->c.cnt;
  out->kernel_launches = c.kernel_launches; out->gemm_calls = c.gemm_calls; out->gemm_flops = c.gemm_flops;
  out->permute_bytes = c.permute_bytes; out->matvecs = c.matvecs; out->env_builds = c.env_builds;
  out->qr_calls = c.qr_calls; out->svd_calls = c.svd_calls; out->jacobi_sweeps = c.jacobi_sweeps;
  return NSB_OK;
}
int nsb_ctx_counters_reset(nsb_ctx* ctx) { if (!ctx) return NSB_EINVAL; ctx->c.cnt = Counters(); return NSB_OK; }
int nsb_ctx_synchronize(nsb_ctx* ctx) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c) ctx->c.sync(); NSB_CATCH(&ctx->c)
}
int nsb_event_tic(nsb_ctx* ctx) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  if (!ctx->ev0) { NSB_CUDA(cudaEventCreate(&ctx->ev0)); NSB_CUDA(cudaEventCreate(&ctx->ev1)); }
  NSB_CUDA(cudaEventRecord(ctx->ev0, ctx->c.stream));
  NSB_CATCH(&ctx->c)
}
int nsb_event_toc(nsb_ctx* ctx, double* ms_out) {
  if (!ctx || !ms_out) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_REQUIRE(ctx->ev0, NSB_EINVAL, "nsb_event_toc without nsb_event_tic");
  NSB_CUDA(cudaEventRecord(ctx->ev1, ctx->c.stream));
  NSB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0;
  NSB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_out = ms;
  NSB_CATCH(&ctx->c)
}
int nsb_timers_enable(nsb_ctx* ctx, int on) { (void)ctx; g_timers_enabled = on != 0; return NSB_OK; }
int nsb_timers_get(nsb_ctx* ctx, double* ms_out) {
  if (!ctx || !ms_out) return NSB_EINVAL;
  for (int i = 0; i < NSB_NUM_TIMERS; ++i) ms_out[i] = ctx->c.timers_ms[i];
  return NSB_OK;
}
int nsb_timers_reset(nsb_ctx* ctx) { if (!ctx) return NSB_EINVAL; for (auto& t : ctx->c.timers_ms) t = 0; return NSB_OK; }
int nsb_gemm_profile_enable(nsb_ctx* ctx, int32_t on) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  ctx->c.sync();
  for (auto& g : ctx->c.gemm_prof) { cudaEventDestroy(g.e0); cudaEventDestroy(g.e1); }
  ctx->c.gemm_prof.clear();
  ctx->c.gemm_profile = on != 0;
  NSB_CATCH(&ctx->c)
}
int nsb_gemm_profile_read(nsb_ctx* ctx, int64_t cap, double* ms_out, double* flops_out, int64_t* mnkb_out, int64_t* count_out) {
  if (!ctx || !count_out) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  ctx->c.sync();
  const int64_t n = (int64_t)ctx->c.gemm_prof.size();
  *count_out = n;
  for (int64_t i = 0; i < n && i < cap; ++i) {
    auto& g = ctx->c.gemm_prof[i];
    float ms = 0.f;
    NSB_CUDA(cudaEventElapsedTime(&ms, g.e0, g.e1));
    if (ms_out) ms_out[i] = ms;
    if (flops_out) flops_out[i] = g.flops;
    if (mnkb_out) { mnkb_out[4 * i] = g.M; mnkb_out[4 * i + 1] = g.N; mnkb_out[4 * i + 2] = g.K; mnkb_out[4 * i + 3] = g.batch; }
  }
  NSB_CATCH(&ctx->c)
}
int nsb_profiler(nsb_ctx* ctx, int32_t on) {   // cudaProfilerStart / Stop: `ncu --profile-from-start off` then sees only the marked region
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  ctx->c.sync();
  if (on) NSB_CUDA(cudaProfilerStart()); else NSB_CUDA(cudaProfilerStop());
  NSB_CATCH(&ctx->c)
}
int nsb_mem_info(nsb_ctx* ctx, int64_t* free_bytes, int64_t* total_bytes, int64_t* pool_used_bytes) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  size_t f, t;
  NSB_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  if (pool_used_bytes) { uint64_t u = 0; NSB_CUDA(cudaMemPoolGetAttribute(ctx->c.pool, cudaMemPoolAttrUsedMemCurrent, &u)); *pool_used_bytes = (int64_t)u; }
  NSB_CATCH(&ctx->c)
}

// ---- NCCL plumbing -------------------------------------------------------------------------------
int nsb_comm_unique_id(char id_out[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  NSB_TRY(nullptr)
  ncclUniqueId id;
  if (nccl_api().GetUniqueId(&id) != ncclSuccess) throw Error(NSB_ENCCL, "ncclGetUniqueId failed");
  memcpy(id_out, &id, 128);
  NSB_CATCH(nullptr)
}
int nsb_comm_init(nsb_ctx* ctx, const char id[128], int rank, int nranks) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclComm_t comm;
  ncclResult_t r = nccl_api().CommInitRank(&comm, nranks, uid, rank);
  if (r != ncclSuccess) throw Error(NSB_ENCCL, std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(r));
  ctx->c.nccl_comm = comm;
  ctx->c.rank = rank;
  ctx->c.nranks = nranks;
  // bring every channel of the three collectives the region step uses up now (NCCL connects lazily inside the first call of each
  // kind): the first H_eff application then runs on established connections
  {
    const size_t n = (size_t)nranks * 512;
    DevBuf a(&ctx->c, sizeof(double) * n), b(&ctx->c, sizeof(double) * n);
    NSB_CUDA(cudaMemsetAsync(a.ptr, 0, sizeof(double) * n, ctx->c.stream));
    auto chk = [&](ncclResult_t rr, const char* what) { if (rr != ncclSuccess) throw Error(NSB_ENCCL, std::string(what) + ": " + nccl_api().GetErrorString(rr)); };
    chk(nccl_api().AllReduce(a.ptr, a.ptr, n, ncclDouble, ncclSum, comm, ctx->c.stream), "ncclAllReduce(warm-up)");
    chk(nccl_api().ReduceScatter(a.ptr, b.ptr, 512, ncclDouble, ncclSum, comm, ctx->c.stream), "ncclReduceScatter(warm-up)");
    chk(nccl_api().AllGather(b.ptr, a.ptr, 512, ncclDouble, comm, ctx->c.stream), "ncclAllGather(warm-up)");
    ctx->c.sync();
  }
  NSB_CATCH(&ctx->c)
}
int nsb_comm_destroy(nsb_ctx* ctx) {
  if (!ctx) return NSB_EINVAL;
  if (ctx->c.nccl_comm) { try { nccl_api().CommDestroy((ncclComm_t)ctx->c.nccl_comm); } catch (...) {} ctx->c.nccl_comm = nullptr; }
  ctx->c.rank = 0; ctx->c.nranks = 1;
  return NSB_OK;
}

// ---- peer-memory windows (fused GEMM + reduce-scatter epilogue) ------------------------------------
int nsb_peer_window_create(nsb_ctx* ctx, int64_t bytes, char handle_out[64]) {
  if (!ctx || !handle_out) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  if (ctx->c.win_local && ctx->c.win_bytes < (size_t)bytes) { NSB_CUDA(cudaFree(ctx->c.win_local)); ctx->c.win_local = nullptr; }
  if (!ctx->c.win_local) { NSB_CUDA(cudaMalloc(&ctx->c.win_local, (size_t)bytes)); ctx->c.win_bytes = (size_t)bytes; }
  cudaIpcMemHandle_t h;
  NSB_CUDA(cudaIpcGetMemHandle(&h, ctx->c.win_local));
  memcpy(handle_out, &h, 64);
  ctx->c.win_peer[ctx->c.rank] = ctx->c.win_local;
  NSB_CATCH(&ctx->c)
}
int nsb_peer_window_open(nsb_ctx* ctx, int32_t peer_rank, const char handle[64]) {
  if (!ctx || !handle) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NSB_REQUIRE(peer_rank >= 0 && peer_rank < 8 && peer_rank < ctx->c.nranks, NSB_EINVAL, "bad peer rank");
  if (peer_rank == ctx->c.rank) { ctx->c.win_peer[peer_rank] = ctx->c.win_local; }
  else if (ctx->c.win_peer[peer_rank] && memcmp(ctx->c.win_peer_handle[peer_rank], handle, 64) == 0) {
    // same allocation already mapped
  } else {
    if (ctx->c.win_peer[peer_rank]) { cudaIpcCloseMemHandle(ctx->c.win_peer[peer_rank]); ctx->c.win_peer[peer_rank] = nullptr; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    NSB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->c.win_peer[peer_rank] = p;
    memcpy(ctx->c.win_peer_handle[peer_rank], handle, 64);
  }
  NSB_CATCH(&ctx->c)
}

// ---- network -------------------------------------------------------------------------------------
int nsb_network_create(nsb_ctx* ctx, int32_t nverts, const int32_t* edges, int32_t nedges, const int64_t* site_dims,
                       int32_t dtype, nsb_net** out) {
  if (!ctx || !out || !site_dims || (nedges > 0 && !edges)) return fail(ctx ? &ctx->c : nullptr, NSB_EINVAL, "null argument");
  *out = nullptr;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NetBase* n = nullptr;
  if (dtype == NSB_F64) n = new Net<double>(&ctx->c, nverts, edges, nedges, site_dims);
  else if (dtype == NSB_C128) n = new Net<cdouble>(&ctx->c, nverts, edges, nedges, site_dims);
  else throw Error(NSB_EINVAL, "dtype must be NSB_F64 or NSB_C128");
  nsb_net* h = new nsb_net();
  h->n = n; h->ctx = ctx;
  ctx->refs++;
  *out = h;
  NSB_CATCH(&ctx->c)
}
int nsb_network_destroy(nsb_net* net) {
  if (!net) return NSB_OK;
  nsb_ctx* ctx = net->ctx;
  cudaSetDevice(ctx->c.device);
  delete net->n;
  delete net;
  if (--ctx->refs == 0 && ctx->destroyed) ctx_really_destroy(ctx);
  return NSB_OK;
}

#define NET_CALL(net, body)                              \
  if (!(net)) return fail(nullptr, NSB_EINVAL, "null network"); \
  Ctx* cx__ = &(net)->ctx->c;                            \
  NSB_TRY(cx__)                                          \
  NSB_CUDA(cudaSetDevice(cx__->device));                 \
  body;                                                  \
  NSB_CATCH(cx__)

int nsb_site_upload(nsb_net* net, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, const void* host) {
  NET_CALL(net, NSB_REQUIRE(legs && dims && host && rank >= 1 && rank <= MAX_RANK, NSB_EINVAL, "bad arguments"); net->n->site_upload(v, rank, legs, dims, host))
}
int nsb_site_info(nsb_net* net, int32_t v, int32_t* rank, int32_t* legs, int64_t* dims) {
  NET_CALL(net, NSB_REQUIRE(rank, NSB_EINVAL, "null rank"); net->n->site_info(v, rank, legs, dims))
}
int nsb_site_download(nsb_net* net, int32_t v, void* host) {
  NET_CALL(net, NSB_REQUIRE(host, NSB_EINVAL, "null buffer"); net->n->site_download(v, host))
}
int nsb_site_fill_random(nsb_net* net, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, uint64_t seed, double scale) {
  NET_CALL(net, NSB_REQUIRE(legs && dims && rank >= 1 && rank <= MAX_RANK, NSB_EINVAL, "bad arguments"); net->n->site_fill_random(v, rank, legs, dims, seed, scale))
}
int nsb_mpo_upload(nsb_net* net, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, const void* host) {
  NET_CALL(net, NSB_REQUIRE(legs && dims && host && rank >= 2 && rank <= MAX_RANK, NSB_EINVAL, "bad arguments"); net->n->mpo_upload(v, rank, legs, dims, host))
}
int nsb_set_ortho_region(nsb_net* net, const int32_t* verts, int32_t n) {
  NET_CALL(net, NSB_REQUIRE(n >= 0 && (n == 0 || verts), NSB_EINVAL, "bad arguments"); net->n->set_ortho_region(verts, n))
}
int nsb_get_ortho_region(nsb_net* net, int32_t* verts, int32_t* n) {
  NET_CALL(net, NSB_REQUIRE(n, NSB_EINVAL, "null n"); net->n->get_ortho_region(verts, n))
}
int nsb_linkdim(nsb_net* net, int32_t u, int32_t v, int64_t* dim) {
  NET_CALL(net, NSB_REQUIRE(dim, NSB_EINVAL, "null dim"); *dim = net->n->linkdim(u, v))
}
int nsb_maxlinkdim(nsb_net* net, int64_t* dim) {
  NET_CALL(net, NSB_REQUIRE(dim, NSB_EINVAL, "null dim"); *dim = net->n->maxlinkdim())
}
int nsb_env_drop_all(nsb_net* net) { NET_CALL(net, net->n->env_drop_all()) }
int nsb_env_count(nsb_net* net, int32_t* n) { NET_CALL(net, NSB_REQUIRE(n, NSB_EINVAL, "null n"); *n = net->n->env_count()) }

int nsb_extract(nsb_net* net, const int32_t* region, int32_t nreg, const nsb_trunc* trunc, const nsb_expand* expand, nsb_extract_info* info) {
  NET_CALL(net, NSB_REQUIRE(region, NSB_EINVAL, "null region"); net->n->extract(region, nreg, trunc, expand, info))
}
int nsb_update_eigsolve(nsb_net* net, const nsb_krylov* params, double* eigval, nsb_solve_info* info) {
  NET_CALL(net, net->n->update_eigsolve(params, eigval, info))
}
int nsb_update_exp(nsb_net* net, double t_re, double t_im, int32_t solver, const nsb_krylov* params, int32_t nsites,
                   int32_t next_vertex, nsb_solve_info* info) {
  NET_CALL(net, net->n->update_exp(t_re, t_im, solver, params, nsites, next_vertex, info))
}
int nsb_fit_target_upload(nsb_net* net, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, const void* host) {
  NET_CALL(net, NSB_REQUIRE(legs && dims && host && rank >= 1 && rank <= MAX_RANK, NSB_EINVAL, "bad arguments"); net->n->fit_target_upload(v, rank, legs, dims, host))
}
int nsb_update_fit(nsb_net* net, double* overlap) {
  NET_CALL(net, NSB_REQUIRE(overlap, NSB_EINVAL, "null overlap"); *overlap = net->n->update_fit())
}
int nsb_insert(nsb_net* net, const nsb_trunc* trunc, int32_t normalize, int32_t set_ortho, nsb_insert_info* info) {
  NET_CALL(net, net->n->insert(trunc, normalize, set_ortho, info))
}
int nsb_local_info(nsb_net* net, int32_t* rank, int32_t* legs, int64_t* dims) {
  NET_CALL(net, NSB_REQUIRE(rank, NSB_EINVAL, "null rank"); net->n->local_info(rank, legs, dims))
}
int nsb_local_download(nsb_net* net, void* host) { NET_CALL(net, NSB_REQUIRE(host, NSB_EINVAL, "null buffer"); net->n->local_download(host)) }
int nsb_local_sync(nsb_net* net) { NET_CALL(net, net->n->local_sync()) }
int nsb_local_upload(nsb_net* net, const void* host) { NET_CALL(net, NSB_REQUIRE(host, NSB_EINVAL, "null buffer"); net->n->local_upload(host)) }
int nsb_matvec_host(nsb_net* net, const void* host_in, void* host_out) {
  NET_CALL(net, NSB_REQUIRE(host_in && host_out, NSB_EINVAL, "null buffer"); net->n->matvec_host(host_in, host_out))
}
int nsb_matvec_host_slab(nsb_net* net, const void* host_in, void* host_out) {
  NET_CALL(net, NSB_REQUIRE(host_in && host_out, NSB_EINVAL, "null buffer"); net->n->matvec_host_slab(host_in, host_out))
}
int nsb_env_bytes(nsb_net* net, int64_t* resident, int64_t* replicated) { NET_CALL(net, net->n->env_bytes(resident, replicated)) }
int nsb_shard_range(nsb_net* net, int64_t* lo, int64_t* hi, int64_t* last_dim) { NET_CALL(net, net->n->shard_range(lo, hi, last_dim)) }
int nsb_matvec_device(nsb_net* net, int32_t reps, void* host_out) { NET_CALL(net, net->n->matvec_device(reps, host_out)) }
int nsb_matvec_flops(nsb_net* net, double* flops) { NET_CALL(net, NSB_REQUIRE(flops, NSB_EINVAL, "null"); *flops = net->n->matvec_flops()) }
int nsb_matvec_flops_executed(nsb_net* net, double* flops) {
  NET_CALL(net, NSB_REQUIRE(flops, NSB_EINVAL, "null"); *flops = net->n->matvec_flops_executed())
}
int nsb_net_set_shard(nsb_net* net, int32_t enable, int32_t* active) {
  NET_CALL(net, int a = net->n->set_shard(enable); if (active) *active = a)
}
int nsb_shard_emulate(nsb_net* net, int32_t nranks, void* host_out, int32_t* mode_out) {
  NET_CALL(net, NSB_REQUIRE(host_out, NSB_EINVAL, "null buffer"); net->n->shard_emulate(nranks, host_out, mode_out))
}
int nsb_qn_enable(nsb_net* net, int32_t nq, const int32_t* total) {
  NET_CALL(net, NSB_REQUIRE(total, NSB_EINVAL, "null"); net->n->qn_enable(nq, total))
}
int nsb_qn_set_site(nsb_net* net, int32_t v, const int32_t* charges) {
  NET_CALL(net, NSB_REQUIRE(charges, NSB_EINVAL, "null"); net->n->qn_set_site(v, charges))
}
int nsb_qn_set_link(nsb_net* net, int32_t u, int32_t v, const int32_t* charges) {
  NET_CALL(net, NSB_REQUIRE(charges, NSB_EINVAL, "null"); net->n->qn_set_link(u, v, charges))
}
int nsb_qn_get_link(nsb_net* net, int32_t u, int32_t v, int32_t* charges_out) {
  NET_CALL(net, NSB_REQUIRE(charges_out, NSB_EINVAL, "null"); net->n->qn_get_link(u, v, charges_out))
}
int nsb_qn_project(nsb_net* net, int32_t v) { NET_CALL(net, net->n->qn_project(v)) }
int nsb_norm(nsb_net* net, double* out) { NET_CALL(net, NSB_REQUIRE(out, NSB_EINVAL, "null"); *out = net->n->norm()) }

#pragma GCC visibility pop
}  // extern "C"

// ---- dense helpers -------------------------------------------------------------------------------
template <typename T>
static void gemm_host_impl(Ctx* c, int opa, int opb, int64_t m, int64_t n, int64_t k, const void* A, int64_t lda,
                           const void* B, int64_t ldb, void* C, int64_t ldc, int impl) {
  int64_t acols = (opa == OP_N || opa == OP_CONJ) ? k : m, bcols = (opb == OP_N || opb == OP_CONJ) ? n : k;
  DevBuf dA(c, sizeof(T) * lda * acols), dB(c, sizeof(T) * ldb * bcols), dC(c, sizeof(T) * ldc * n);
  NSB_CUDA(cudaMemcpyAsync(dA.ptr, A, sizeof(T) * lda * acols, cudaMemcpyHostToDevice, c->stream));
  NSB_CUDA(cudaMemcpyAsync(dB.ptr, B, sizeof(T) * ldb * bcols, cudaMemcpyHostToDevice, c->stream));
  NSB_CUDA(cudaMemsetAsync(dC.ptr, 0xff, sizeof(T) * ldc * n, c->stream));   // NaN-fill: catches unwritten tiles
  gemm<T>(c, opa, opb, m, n, k, from_complex<T>(1.0, 0.0), (const T*)dA.ptr, lda, 0, (const T*)dB.ptr, ldb, 0, zero_<T>(),
          (T*)dC.ptr, ldc, 0, 1, impl);
  NSB_CUDA(cudaMemcpyAsync(C, dC.ptr, sizeof(T) * ldc * n, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
}
extern "C" __attribute__((visibility("default"))) int nsb_gemm_host(nsb_ctx* ctx, int32_t dtype, int32_t opa, int32_t opb, int64_t m, int64_t n, int64_t k, const void* A,
                  int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int32_t impl) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NSB_REQUIRE(A && B && C && m > 0 && n > 0 && k > 0, NSB_EINVAL, "bad arguments");
  if (dtype == NSB_F64) gemm_host_impl<double>(&ctx->c, opa, opb, m, n, k, A, lda, B, ldb, C, ldc, impl);
  else gemm_host_impl<cdouble>(&ctx->c, opa, opb, m, n, k, A, lda, B, ldb, C, ldc, impl);
  NSB_CATCH(&ctx->c)
}

template <typename T>
static double gemm_bench_impl(Ctx* c, int opa, int opb, int64_t m, int64_t n, int64_t k, int impl, int reps) {
  int64_t lda = (opa == OP_N || opa == OP_CONJ) ? m : k, acols = (opa == OP_N || opa == OP_CONJ) ? k : m;
  int64_t ldb = (opb == OP_N || opb == OP_CONJ) ? k : n, bcols = (opb == OP_N || opb == OP_CONJ) ? n : k;
  DevBuf dA(c, sizeof(T) * lda * acols), dB(c, sizeof(T) * ldb * bcols), dC(c, sizeof(T) * m * n);
  fill_normal<T>(c, (T*)dA.ptr, lda * acols, 11, 1.0);
  fill_normal<T>(c, (T*)dB.ptr, ldb * bcols, 12, 1.0);
  auto run = [&]() {
    gemm<T>(c, opa, opb, m, n, k, from_complex<T>(1.0, 0.0), (const T*)dA.ptr, lda, 0, (const T*)dB.ptr, ldb, 0, zero_<T>(),
            (T*)dC.ptr, m, 0, 1, impl);
  };
  for (int i = 0; i < 2; ++i) run();
  cudaEvent_t e0, e1;
  NSB_CUDA(cudaEventCreate(&e0)); NSB_CUDA(cudaEventCreate(&e1));
  NSB_CUDA(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < reps; ++i) run();
  NSB_CUDA(cudaEventRecord(e1, c->stream));
  NSB_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  NSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return ms / reps;
}
extern "C" __attribute__((visibility("default"))) int nsb_gemm_bench(nsb_ctx* ctx, int32_t dtype, int32_t opa, int32_t opb, int64_t m, int64_t n, int64_t k, int32_t impl,
                   int32_t reps, double* ms_out) {
  if (!ctx || !ms_out) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NSB_REQUIRE(m > 0 && n > 0 && k > 0 && reps > 0, NSB_EINVAL, "bad arguments");
  *ms_out = dtype == NSB_F64 ? gemm_bench_impl<double>(&ctx->c, opa, opb, m, n, k, impl, reps)
                             : gemm_bench_impl<cdouble>(&ctx->c, opa, opb, m, n, k, impl, reps);
  NSB_CATCH(&ctx->c)
}

extern "C" __attribute__((visibility("default"))) int nsb_dmma_peak(nsb_ctx* ctx, double* tflops_out) {
  if (!ctx || !tflops_out) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  *tflops_out = dmma_peak_tflops(&ctx->c);
  NSB_CATCH(&ctx->c)
}

template <typename T>
static void factorize_host_impl(Ctx* c, int64_t rows, int64_t cols, const void* M, const nsb_trunc* trunc, void* U, void* C,
                                double* spectrum, nsb_insert_info* info) {
  nsb_trunc tr = trunc ? *trunc : nsb_trunc{0.0, 1, INT64_MAX};
  DevBuf dM(c, sizeof(T) * rows * cols), Ub, Cb;
  NSB_CUDA(cudaMemcpyAsync(dM.ptr, M, sizeof(T) * rows * cols, cudaMemcpyHostToDevice, c->stream));
  std::vector<double> spec;
  FactorInfo fi = factorize_left<T>(c, (const T*)dM.ptr, rows, cols, rows, false, tr.cutoff, tr.mindim, tr.maxdim, false, Ub, Cb, spec);
  if (U) NSB_CUDA(cudaMemcpyAsync(U, Ub.ptr, sizeof(T) * rows * fi.newdim, cudaMemcpyDeviceToHost, c->stream));
  if (C) NSB_CUDA(cudaMemcpyAsync(C, Cb.ptr, sizeof(T) * fi.newdim * cols, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  if (spectrum) for (size_t i = 0; i < spec.size(); ++i) spectrum[i] = spec[i];
  if (info) { info->newdim = fi.newdim; info->truncerr = fi.truncerr; info->decomp = fi.decomp; info->jacobi_sweeps = fi.sweeps; }
}
extern "C" __attribute__((visibility("default"))) int nsb_factorize_host(nsb_ctx* ctx, int32_t dtype, int64_t rows, int64_t cols, const void* M, const nsb_trunc* trunc, void* U,
                       void* C, double* spectrum, nsb_insert_info* info) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NSB_REQUIRE(M && rows > 0 && cols > 0, NSB_EINVAL, "bad arguments");
  if (dtype == NSB_F64) factorize_host_impl<double>(&ctx->c, rows, cols, M, trunc, U, C, spectrum, info);
  else factorize_host_impl<cdouble>(&ctx->c, rows, cols, M, trunc, U, C, spectrum, info);
  NSB_CATCH(&ctx->c)
}

template <typename T>
static void eigh_host_impl(Ctx* c, int64_t n, const void* A, double* w, void* U) {
  DevBuf dA(c, sizeof(T) * n * n), dU(c, sizeof(T) * n * n);
  NSB_CUDA(cudaMemcpyAsync(dA.ptr, A, sizeof(T) * n * n, cudaMemcpyHostToDevice, c->stream));
  Eigh<T> eg;
  eg.factor(c, (T*)dA.ptr, n, n);
  std::vector<int32_t> order(n);
  for (int64_t i = 0; i < n; ++i) order[i] = (int32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return eg.w[a] < eg.w[b]; });
  for (int64_t i = 0; i < n; ++i) w[i] = eg.w[order[i]];
  if (U) {
    eg.vectors(order.data(), n, (T*)dU.ptr, n);
    NSB_CUDA(cudaMemcpyAsync(U, dU.ptr, sizeof(T) * n * n, cudaMemcpyDeviceToHost, c->stream));
  }
  c->sync();
}
extern "C" __attribute__((visibility("default"))) int nsb_eigh_host(nsb_ctx* ctx, int32_t dtype, int64_t n, const void* A, double* w, void* U) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NSB_REQUIRE(A && w && n > 0, NSB_EINVAL, "bad arguments");
  if (dtype == NSB_F64) eigh_host_impl<double>(&ctx->c, n, A, w, U);
  else eigh_host_impl<cdouble>(&ctx->c, n, A, w, U);
  NSB_CATCH(&ctx->c)
}

namespace nsb { void sbr_chase_device(Ctx* ctx, int64_t n, int b, double* ab, int64_t ld, double* V2, double* tau2, int64_t ldtau); }
extern "C" __attribute__((visibility("default"))) int nsb_sbr_chase_host(nsb_ctx* ctx, int64_t n, int32_t b, double* ab, int64_t ld, double* V2, double* tau2,
                                                                          int64_t ldtau) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NSB_REQUIRE(ab && V2 && tau2 && n > 0 && b > 0 && ld >= 2 * (int64_t)b + 1 && ldtau >= 1, NSB_EINVAL, "bad arguments");
  Ctx* c = &ctx->c;
  DevBuf dab(c, sizeof(double) * (size_t)ld * n), dV(c, sizeof(double) * (size_t)n * n), dt(c, sizeof(double) * (size_t)ldtau * n);
  NSB_CUDA(cudaMemcpyAsync(dab.ptr, ab, sizeof(double) * (size_t)ld * n, cudaMemcpyHostToDevice, c->stream));
  NSB_CUDA(cudaMemsetAsync(dV.ptr, 0, sizeof(double) * (size_t)n * n, c->stream));
  NSB_CUDA(cudaMemsetAsync(dt.ptr, 0, sizeof(double) * (size_t)ldtau * n, c->stream));
  nsb::sbr_chase_device(c, n, b, (double*)dab.ptr, ld, (double*)dV.ptr, (double*)dt.ptr, ldtau);
  NSB_CUDA(cudaMemcpyAsync(ab, dab.ptr, sizeof(double) * (size_t)ld * n, cudaMemcpyDeviceToHost, c->stream));
  NSB_CUDA(cudaMemcpyAsync(V2, dV.ptr, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, c->stream));
  NSB_CUDA(cudaMemcpyAsync(tau2, dt.ptr, sizeof(double) * (size_t)ldtau * n, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  NSB_CATCH(&ctx->c)
}

template <typename T>
static void qr_host_impl(Ctx* c, int64_t rows, int64_t cols, const void* M, void* Q, void* R) {
  int64_t k = std::min(rows, cols);
  DevBuf dM(c, sizeof(T) * rows * cols), dQ(c, sizeof(T) * rows * k), dR(c, sizeof(T) * k * cols);
  NSB_CUDA(cudaMemcpyAsync(dM.ptr, M, sizeof(T) * rows * cols, cudaMemcpyHostToDevice, c->stream));
  qr_thin<T>(c, (T*)dM.ptr, rows, cols, rows, (T*)dQ.ptr, rows, (T*)dR.ptr, k);
  NSB_CUDA(cudaMemcpyAsync(Q, dQ.ptr, sizeof(T) * rows * k, cudaMemcpyDeviceToHost, c->stream));
  NSB_CUDA(cudaMemcpyAsync(R, dR.ptr, sizeof(T) * k * cols, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
}
extern "C" __attribute__((visibility("default"))) int nsb_qr_host(nsb_ctx* ctx, int32_t dtype, int64_t rows, int64_t cols, const void* M, void* Q, void* R) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NSB_REQUIRE(M && Q && R && rows > 0 && cols > 0, NSB_EINVAL, "bad arguments");
  if (dtype == NSB_F64) qr_host_impl<double>(&ctx->c, rows, cols, M, Q, R);
  else qr_host_impl<cdouble>(&ctx->c, rows, cols, M, Q, R);
  NSB_CATCH(&ctx->c)
}

template <typename T>
static double qr_bench_impl(Ctx* c, int64_t rows, int64_t cols, int reps) {
  const int64_t k = std::min(rows, cols);
  DevBuf dM(c, sizeof(T) * rows * cols), dW(c, sizeof(T) * rows * cols), dQ(c, sizeof(T) * rows * k), dR(c, sizeof(T) * k * cols);
  fill_normal<T>(c, (T*)dM.ptr, rows * cols, 21, 1.0);
  cudaEvent_t e0, e1;
  NSB_CUDA(cudaEventCreate(&e0)); NSB_CUDA(cudaEventCreate(&e1));
  double total = 0.0;
  for (int i = 0; i < reps + 1; ++i) {
    vec_copy<T>(c, rows * cols, (const T*)dM.ptr, (T*)dW.ptr);
    NSB_CUDA(cudaEventRecord(e0, c->stream));
    qr_thin<T>(c, (T*)dW.ptr, rows, cols, rows, (T*)dQ.ptr, rows, (T*)dR.ptr, k);
    NSB_CUDA(cudaEventRecord(e1, c->stream));
    NSB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    NSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (i > 0) total += ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return total / reps;
}
extern "C" __attribute__((visibility("default"))) int nsb_qr_bench(nsb_ctx* ctx, int32_t dtype, int64_t rows, int64_t cols, int32_t reps, double* ms_out) {
  if (!ctx || !ms_out) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NSB_REQUIRE(rows > 0 && cols > 0 && reps > 0, NSB_EINVAL, "bad arguments");
  *ms_out = dtype == NSB_F64 ? qr_bench_impl<double>(&ctx->c, rows, cols, reps) : qr_bench_impl<cdouble>(&ctx->c, rows, cols, reps);
  NSB_CATCH(&ctx->c)
}

// Randomised range finder of a host matrix A (m x n): linear map = one GEMM per probe panel (rangefinder.cu).
template <typename T>
static int64_t range_finder_impl(Ctx* c, int64_t m, int64_t n, const void* A, const void* probes, int64_t max_rank, int oversample,
                                 int north_pass, double thr, double cutoff, uint64_t seed, void* Qh) {
  if (max_rank <= 0) return 0;
  const int64_t sketch = std::min(std::min(max_rank, std::min(m, n)) + oversample, std::min(m, n));
  DevBuf dA(c, sizeof(T) * m * n), dQ(c, sizeof(T) * m * std::max<int64_t>(sketch, 1)), dP(c, probes ? sizeof(T) * n * sketch : 0);
  NSB_CUDA(cudaMemcpyAsync(dA.ptr, A, sizeof(T) * m * n, cudaMemcpyHostToDevice, c->stream));
  if (probes) NSB_CUDA(cudaMemcpyAsync(dP.ptr, probes, sizeof(T) * n * sketch, cudaMemcpyHostToDevice, c->stream));
  RangeMap<T> map = [&](const T* Om, T* Y, int64_t p) {
    gemm<T>(c, OP_N, OP_N, m, p, n, from_complex<T>(1.0, 0.0), (const T*)dA.ptr, m, 0, Om, n, 0, zero_<T>(), Y, m, 0, 1);
  };
  const int64_t have = range_finder_blocked<T>(c, m, n, map, probes ? (const T*)dP.ptr : nullptr, seed, max_rank, oversample, north_pass, thr,
                                               cutoff, (T*)dQ.ptr);
  if (have > 0) NSB_CUDA(cudaMemcpyAsync(Qh, dQ.ptr, sizeof(T) * m * have, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  return have;
}
extern "C" __attribute__((visibility("default"))) int nsb_range_finder_host(nsb_ctx* ctx, int32_t dtype, int64_t m, int64_t n, const void* A, const void* probes,
                          int64_t max_rank, int32_t oversample, int32_t north_pass, double orthogonal_threshold, double cutoff, uint64_t seed,
                          void* Q, int64_t* rank_out) {
  if (!ctx) return NSB_EINVAL;
  NSB_TRY(&ctx->c)
  NSB_CUDA(cudaSetDevice(ctx->c.device));
  NSB_REQUIRE(A && Q && rank_out && m > 0 && n > 0, NSB_EINVAL, "bad arguments");
  *rank_out = dtype == NSB_F64 ? range_finder_impl<double>(&ctx->c, m, n, A, probes, max_rank, oversample, north_pass, orthogonal_threshold, cutoff, seed, Q)
                               : range_finder_impl<cdouble>(&ctx->c, m, n, A, probes, max_rank, oversample, north_pass, orthogonal_threshold, cutoff, seed, Q);
  NSB_CATCH(&ctx->c)
}
extern "C" __attribute__((visibility("default"))) int nsb_range_finder_heff(nsb_net* net, const void* probes, uint64_t seed, int64_t max_rank, int32_t oversample,
                          int32_t north_pass, double orthogonal_threshold, double cutoff, void* Q, int64_t* rank_out) {
  NET_CALL(net, NSB_REQUIRE(Q && rank_out, NSB_EINVAL, "null argument");
           *rank_out = net->n->range_finder_heff(probes, seed, max_rank, oversample, north_pass, orthogonal_threshold, cutoff, Q))
}
extern "C" __attribute__((visibility("default"))) int nsb_expand_set_probe(nsb_net* net, int64_t rows, int64_t cols, const void* host) {
  NET_CALL(net, NSB_REQUIRE(host && rows > 0 && cols > 0, NSB_EINVAL, "bad arguments"); net->n->expand_set_probe(rows, cols, host))
}
