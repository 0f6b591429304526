// Single-process multi-device entry points ("Multi-GPU: single process, one NCCL communicator inside the context").
//
// A host such as the Julia shim is one process with one thread of control; the multi-GPU partition of libnsb200 is one rank
// per GPU running the same region step in lock step (net.cu).  nsb_multi_* bridges the two: an nsb_multi owns one context and
// one replica network per device plus the NCCL communicator joining them, and every nsb_multi_* hook fans the call out to one
// host thread per device (each thread binds its device, so the collectives inside a hook are issued concurrently, as NCCL
// requires for several devices of one process) and returns when all replicas are done.  Scalars are those of device 0 (all
// replicas compute identical values).  The per-device handles stay reachable (nsb_multi_ctx / nsb_multi_net) for everything
// that is not a collective: counters, options, downloads.
#include <thread>

#include "nccl_dyn.h"
#include "net.h"

using namespace nsb;

struct nsb_multi {
  int ndev = 0;
  std::vector<nsb_ctx*> ctxs;
  std::vector<nsb_net*> nets;
  std::string last_error;
};

template <class F>
static int fan_out(nsb_multi* m, F f) {
  std::vector<int> rc(m->ndev, NSB_OK);
  if (m->ndev == 1) { rc[0] = f(0); }
  else {
    std::vector<std::thread> th;
    for (int r = 0; r < m->ndev; ++r) th.emplace_back([&, r] { rc[r] = f(r); });
    for (auto& t : th) t.join();
  }
  for (int r = 0; r < m->ndev; ++r)
    if (rc[r] != NSB_OK) {
      m->last_error = std::string("device ") + std::to_string(r) + ": " + nsb_last_error(m->ctxs[r]);
      return rc[r];
    }
  return NSB_OK;
}

extern "C" {
#pragma GCC visibility push(default)

int nsb_multi_create(const int32_t* devices, int32_t ndev, nsb_multi** out) {
  if (!devices || !out || ndev < 1 || ndev > 8) return NSB_EINVAL;
  *out = nullptr;
  nsb_multi* m = new nsb_multi();
  m->ndev = ndev;
  m->ctxs.assign(ndev, nullptr);
  m->nets.assign(ndev, nullptr);
  for (int r = 0; r < ndev; ++r) {
    int rc = nsb_ctx_create(devices[r], &m->ctxs[r]);
    if (rc != NSB_OK) { for (int q = 0; q < r; ++q) nsb_ctx_destroy(m->ctxs[q]); delete m; return rc; }
  }
  if (ndev > 1) {
    char id[128];
    int rc = nsb_comm_unique_id(id);
    if (rc == NSB_OK) rc = fan_out(m, [&](int r) { return nsb_comm_init(m->ctxs[r], id, r, ndev); });
    if (rc != NSB_OK) { for (auto c : m->ctxs) nsb_ctx_destroy(c); delete m; return rc; }
  }
  *out = m;
  return NSB_OK;
}

int nsb_multi_destroy(nsb_multi* m) {
  if (!m) return NSB_OK;
  for (auto n : m->nets) if (n) nsb_network_destroy(n);
  if (m->ndev > 1) fan_out(m, [&](int r) { return nsb_comm_destroy(m->ctxs[r]); });
  for (auto c : m->ctxs) nsb_ctx_destroy(c);
  delete m;
  return NSB_OK;
}

const char* nsb_multi_last_error(nsb_multi* m) { return m ? m->last_error.c_str() : "null nsb_multi"; }
int nsb_multi_ndev(nsb_multi* m, int32_t* ndev) { if (!m || !ndev) return NSB_EINVAL; *ndev = m->ndev; return NSB_OK; }
int nsb_multi_ctx(nsb_multi* m, int32_t r, nsb_ctx** out) { if (!m || !out || r < 0 || r >= m->ndev) return NSB_EINVAL; *out = m->ctxs[r]; return NSB_OK; }
int nsb_multi_net(nsb_multi* m, int32_t r, nsb_net** out) { if (!m || !out || r < 0 || r >= m->ndev || !m->nets[r]) return NSB_EINVAL; *out = m->nets[r]; return NSB_OK; }

int nsb_multi_network_create(nsb_multi* m, int32_t nverts, const int32_t* edges, int32_t nedges, const int64_t* site_dims, int32_t dtype) {
  if (!m) return NSB_EINVAL;
  for (auto& n : m->nets) if (n) { nsb_network_destroy(n); n = nullptr; }
  return fan_out(m, [&](int r) { return nsb_network_create(m->ctxs[r], nverts, edges, nedges, site_dims, dtype, &m->nets[r]); });
}
#define MULTI_REQUIRE_NET(m) if (!(m) || (m)->nets.empty() || !(m)->nets[0]) return NSB_EINVAL
int nsb_multi_site_upload(nsb_multi* m, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, const void* host) {
  MULTI_REQUIRE_NET(m);
  return fan_out(m, [&](int r) { return nsb_site_upload(m->nets[r], v, rank, legs, dims, host); });
}
int nsb_multi_mpo_upload(nsb_multi* m, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, const void* host) {
  MULTI_REQUIRE_NET(m);
  return fan_out(m, [&](int r) { return nsb_mpo_upload(m->nets[r], v, rank, legs, dims, host); });
}
int nsb_multi_site_fill_random(nsb_multi* m, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, uint64_t seed, double scale) {
  MULTI_REQUIRE_NET(m);
  return fan_out(m, [&](int r) { return nsb_site_fill_random(m->nets[r], v, rank, legs, dims, seed, scale); });
}
int nsb_multi_set_ortho_region(nsb_multi* m, const int32_t* verts, int32_t n) {
  MULTI_REQUIRE_NET(m);
  return fan_out(m, [&](int r) { return nsb_set_ortho_region(m->nets[r], verts, n); });
}
int nsb_multi_set_shard(nsb_multi* m, int32_t enable, int32_t* active) {
  MULTI_REQUIRE_NET(m);
  std::vector<int32_t> a(m->ndev, 0);
  int rc = fan_out(m, [&](int r) { return nsb_net_set_shard(m->nets[r], enable, &a[r]); });
  if (active) *active = a[0];
  return rc;
}
int nsb_multi_extract(nsb_multi* m, const int32_t* region, int32_t nreg, const nsb_trunc* trunc, const nsb_expand* expand, nsb_extract_info* info) {
  MULTI_REQUIRE_NET(m);
  std::vector<nsb_extract_info> inf(m->ndev);
  int rc = fan_out(m, [&](int r) { return nsb_extract(m->nets[r], region, nreg, trunc, expand, &inf[r]); });
  if (info) *info = inf[0];
  return rc;
}
int nsb_multi_update_eigsolve(nsb_multi* m, const nsb_krylov* params, double* eigval, nsb_solve_info* info) {
  MULTI_REQUIRE_NET(m);
  std::vector<double> ev(m->ndev, 0.0);
  std::vector<nsb_solve_info> inf(m->ndev);
  int rc = fan_out(m, [&](int r) { return nsb_update_eigsolve(m->nets[r], params, &ev[r], &inf[r]); });
  if (eigval) *eigval = ev[0];
  if (info) *info = inf[0];
  return rc;
}
int nsb_multi_update_exp(nsb_multi* m, double t_re, double t_im, int32_t solver, const nsb_krylov* params, int32_t nsites, int32_t next_vertex,
                         nsb_solve_info* info) {
  MULTI_REQUIRE_NET(m);
  std::vector<nsb_solve_info> inf(m->ndev);
  int rc = fan_out(m, [&](int r) { return nsb_update_exp(m->nets[r], t_re, t_im, solver, params, nsites, next_vertex, &inf[r]); });
  if (info) *info = inf[0];
  return rc;
}
int nsb_multi_insert(nsb_multi* m, const nsb_trunc* trunc, int32_t normalize, int32_t set_ortho, nsb_insert_info* info) {
  MULTI_REQUIRE_NET(m);
  std::vector<nsb_insert_info> inf(m->ndev);
  int rc = fan_out(m, [&](int r) { return nsb_insert(m->nets[r], trunc, normalize, set_ortho, &inf[r]); });
  if (info) *info = inf[0];
  return rc;
}
int nsb_multi_matvec_device(nsb_multi* m, int32_t reps) {
  MULTI_REQUIRE_NET(m);
  return fan_out(m, [&](int r) { return nsb_matvec_device(m->nets[r], reps, nullptr); });
}
// local tensor of device 0 on the host (a sharded local tensor is completed by a collective on every device first)
int nsb_multi_local_download(nsb_multi* m, void* host) {
  MULTI_REQUIRE_NET(m);
  int rc = fan_out(m, [&](int r) { return nsb_local_sync(m->nets[r]); });
  if (rc != NSB_OK) return rc;
  return nsb_local_download(m->nets[0], host);
}
int nsb_multi_synchronize(nsb_multi* m) {
  if (!m) return NSB_EINVAL;
  return fan_out(m, [&](int r) { return nsb_ctx_synchronize(m->ctxs[r]); });
}

#pragma GCC visibility pop
}
