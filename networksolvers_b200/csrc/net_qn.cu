// Abelian quantum-number conservation (dense storage, block-wise factorisations).
//
// Counterpart of running the reference on QN-conserving ITensors (`siteinds(...; conserve_qns=true)`): the three
// factorisations on the path -- the inserter's truncating factorisation, the expansion's `eigen`, the gauge `qr` --
// never mix symmetry sectors and the truncation acts on the merged spectrum (NDTensors rule).
// Tensors stay dense: entries forbidden by symmetry are exact zeros, and contractions of symmetric tensors keep them
// exact, so the matvec / environment kernels are unchanged.  Charge convention: see oracle/qn.py.
#include "net.h"

#include <algorithm>
#include <atomic>
#include <functional>
#include <numeric>
#include <thread>

namespace nsb {

#define LAUNCH_CHECK(ctx) do { (ctx)->cnt.kernel_launches++; NSB_CUDA(cudaGetLastError()); } while (0)

template <typename T>
__global__ void gather_block_kernel(const T* __restrict__ M, int64_t ld, const int32_t* __restrict__ rows, int64_t nr,
                                    const int32_t* __restrict__ cols, int64_t nc, T* __restrict__ out) {
  int64_t total = nr * nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i % nr, c = i / nr;
    out[i] = M[(int64_t)rows[r] + (int64_t)cols[c] * ld];
  }
}
// out[rows[r] + (c0 + c) * ldo] = in[r + c * ldi]   (block rows scattered, consecutive output columns)
template <typename T>
__global__ void scatter_rows_kernel(const T* __restrict__ in, int64_t ldi, const int32_t* __restrict__ rows, int64_t nr,
                                    int64_t nc, T* __restrict__ out, int64_t ldo, int64_t c0) {
  int64_t total = nr * nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i % nr, c = i / nr;
    out[(int64_t)rows[r] + (c0 + c) * ldo] = in[r + c * ldi];
  }
}
// out[(r0 + r) + cols[c] * ldo] = in[r + c * ldi]   (consecutive output rows, block columns scattered)
template <typename T>
__global__ void scatter_cols_kernel(const T* __restrict__ in, int64_t ldi, int64_t nr, const int32_t* __restrict__ cols,
                                    int64_t nc, T* __restrict__ out, int64_t ldo, int64_t r0) {
  int64_t total = nr * nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i % nr, c = i / nr;
    out[(r0 + r) + (int64_t)cols[c] * ldo] = in[r + c * ldi];
  }
}

static inline int grid1d(Ctx* ctx, int64_t n) {
  int64_t b = (n + 255) / 256, cap = (int64_t)ctx->num_sms * 8;
  return (int)std::max<int64_t>(1, std::min(b, cap));
}

namespace {
struct Block { int64_t key; std::vector<int32_t> rows, cols; };

std::vector<Block> make_blocks(const std::vector<int64_t>& rk, const std::vector<int64_t>& ck) {
  std::map<int64_t, Block> m;
  std::vector<int64_t> order;
  for (size_t i = 0; i < rk.size(); ++i) {
    auto it = m.find(rk[i]);
    if (it == m.end()) { order.push_back(rk[i]); it = m.emplace(rk[i], Block{rk[i], {}, {}}).first; }
    it->second.rows.push_back((int32_t)i);
  }
  for (size_t j = 0; j < ck.size(); ++j) {
    auto it = m.find(ck[j]);
    if (it != m.end()) it->second.cols.push_back((int32_t)j);
  }
  std::vector<Block> out;
  for (int64_t k : order) { Block& b = m[k]; if (!b.cols.empty()) out.push_back(b); }
  return out;
}

// NDTensors block-sparse truncation: merge the spectra, truncate globally, every block keeps the values above docut.
std::vector<int64_t> truncate_merged(const std::vector<std::vector<double>>& P, double cutoff, int64_t mindim, int64_t maxdim,
                                     double* truncerr) {
  std::vector<double> all;
  for (auto& p : P) all.insert(all.end(), p.begin(), p.end());
  std::sort(all.begin(), all.end(), std::greater<double>());
  std::vector<int64_t> keep(P.size(), 0);
  if (truncerr) *truncerr = 0.0;
  if (all.empty()) return keep;
  int64_t n = truncate_spectrum(all, cutoff, mindim, maxdim, truncerr);
  double docut = -1.0;
  if (n < (int64_t)all.size()) {
    docut = (all[n - 1] + all[n]) / 2.0;
    if (std::fabs(all[n - 1] - all[n]) < 1e-3 * all[n - 1]) docut += 1e-3 * all[n - 1];
  }
  int64_t tot = 0;
  for (size_t b = 0; b < P.size(); ++b) {
    for (double x : P[b]) if (std::max(x, 0.0) > docut) keep[b]++;
    tot += keep[b];
  }
  if (tot == 0) {
    size_t best = 0; double mx = -1;
    for (size_t b = 0; b < P.size(); ++b) if (!P[b].empty() && P[b][0] > mx) { mx = P[b][0]; best = b; }
    keep[best] = 1;
  }
  return keep;
}
}  // namespace

static inline int64_t pack_key(const int64_t* q, int nq) {
  int64_t k = 0;
  for (int c = 0; c < nq; ++c) k |= ((q[c] + 32768) & 0xffff) << (16 * c);
  return k;
}
static inline void unpack_key(int64_t k, int nq, int64_t* q) {
  for (int c = 0; c < nq; ++c) q[c] = ((k >> (16 * c)) & 0xffff) - 32768;
}

template <typename T>
void Net<T>::qn_enable(int nq_, const int32_t* total) {
  NSB_REQUIRE(nq_ >= 1 && nq_ <= 4, NSB_EINVAL, "qn_enable: 1 <= nq <= 4");
  nq = nq_;
  qn_total.assign(total, total + nq_);
  qn_site.assign(nverts, {});
  qn_link.assign(edges.size(), {});
  qn_side.assign(edges.size(), -1);
  qn_on = true;
}
template <typename T>
void Net<T>::qn_set_site(int v, const int32_t* charges) {
  NSB_REQUIRE(qn_on && v >= 0 && v < nverts, NSB_EINVAL, "qn_set_site: bad vertex or QN not enabled");
  qn_site[v].assign(charges, charges + site_dims[v] * nq);
}
template <typename T>
void Net<T>::qn_set_link(int u, int v, const int32_t* charges) {
  NSB_REQUIRE(qn_on && eid.count({u, v}) && psi[u].valid(), NSB_EINVAL, "qn_set_link: bad edge or QN not enabled");
  int e = eid.at({u, v});
  int64_t dim = psi[u].dim_of(llink(u, v));
  qn_link[e].assign(charges, charges + dim * nq);
  qn_side[e] = u;
  if ((int)link_mode.size() > e) link_mode[e] = nullptr;
}
template <typename T>
void Net<T>::qn_get_link(int u, int v, int32_t* out) {
  NSB_REQUIRE(qn_on && eid.count({u, v}), NSB_EINVAL, "qn_get_link: bad edge or QN not enabled");
  std::vector<int64_t> c = side_charge(v, u);   // u's side
  for (size_t i = 0; i < c.size(); ++i) out[i] = (int32_t)c[i];
}

template <typename T>
std::vector<int64_t> Net<T>::side_charge(int v, int n) const {
  int e = eid.at({v, n});
  NSB_REQUIRE(qn_side[e] >= 0, NSB_EINVAL, "QN charges of a link have not been set");
  std::vector<int64_t> c = qn_link[e];
  if (qn_side[e] != n)
    for (size_t i = 0; i < c.size(); ++i) c[i] = qn_total[i % nq] - c[i];
  return c;
}

template <typename T>
std::vector<int64_t> Net<T>::leg_charges(int owner, Label l) const {
  if (label_kind(l) == LK_SITE) {
    NSB_REQUIRE(!qn_site[label_id(l)].empty(), NSB_EINVAL, "QN site charges have not been set");
    return qn_site[label_id(l)];
  }
  NSB_REQUIRE(label_kind(l) == LK_LINK, NSB_EINTERNAL, "leg_charges: unexpected label");
  auto e = edges[label_id(l)];
  int other = (e.first == owner) ? e.second : e.first;
  return side_charge(owner, other);
}

// Packed charge key of every value of the column-major multi-index over `labels` (first label fastest): the sum of
// the leg charges, or total - sum when `complement`.
template <typename T>
std::vector<int64_t> Net<T>::multi_keys(int owner, const std::vector<Label>& labels, const std::vector<int64_t>& dims,
                                        bool complement) const {
  int64_t n = 1;
  for (auto d : dims) n *= d;
  std::vector<int64_t> sum((size_t)n * nq, 0);
  int64_t stride = 1;
  for (size_t k = 0; k < labels.size(); ++k) {
    std::vector<int64_t> ch = leg_charges(owner, labels[k]);
    NSB_REQUIRE((int64_t)ch.size() == dims[k] * nq, NSB_EINTERNAL, "QN charge table does not match the index dimension");
    for (int64_t i = 0; i < n; ++i) {
      int64_t idx = (i / stride) % dims[k];
      for (int c = 0; c < nq; ++c) sum[i * nq + c] += ch[idx * nq + c];
    }
    stride *= dims[k];
  }
  std::vector<int64_t> keys(n);
  std::vector<int64_t> q(nq);
  for (int64_t i = 0; i < n; ++i) {
    for (int c = 0; c < nq; ++c) q[c] = complement ? qn_total[c] - sum[i * nq + c] : sum[i * nq + c];
    keys[i] = pack_key(q.data(), nq);
  }
  return keys;
}

template <typename T>
void Net<T>::qn_store_link(int v, int n, const std::vector<int64_t>& keys) {
  int e = eid.at({v, n});
  if ((int)link_mode.size() > e) link_mode[e] = nullptr;
  qn_link[e].assign(keys.size() * nq, 0);
  for (size_t i = 0; i < keys.size(); ++i) unpack_key(keys[i], nq, &qn_link[e][i * nq]);
  qn_side[e] = v;
}

// Block-wise truncated factorisation M = U C with the merged-spectrum truncation.  M is rows x cols, column-major,
// contiguous.  U (rows x k) and C (k x cols) are dense with zeros outside the symmetry blocks.
template <typename T>
FactorInfo Net<T>::factorize_qn(const T* M, int64_t rows, int64_t cols, const std::vector<int64_t>& rk,
                                const std::vector<int64_t>& ck, double cutoff, int64_t mindim, int64_t maxdim,
                                bool sqrt_spectrum, DevBuf& U, DevBuf& C, std::vector<int64_t>& new_keys) {
  std::vector<Block> blocks = make_blocks(rk, ck);
  NSB_REQUIRE(!blocks.empty(), NSB_EINVAL, "factorize_qn: tensor has no symmetry-allowed block");
  FactorInfo info;
  info.decomp = (cutoff <= 1e-12) ? 1 : 2;
  maxdim = std::min<int64_t>(maxdim, std::min(rows, cols));
  struct Res { DevBuf U, C; std::vector<double> spec; int64_t r, c; DevBuf ridx, cidx; };
  std::vector<Res> res(blocks.size());
  std::vector<std::vector<double>> P(blocks.size());
  // The sectors are independent small factorisations, each a latency-bound chain of launches with host look-ups: run them
  // concurrently, one host thread + helper context (own stream) per lane, largest blocks first.
  std::vector<size_t> order(blocks.size());
  std::iota(order.begin(), order.end(), (size_t)0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    return blocks[a].rows.size() * blocks[a].cols.size() > blocks[b].rows.size() * blocks[b].cols.size(); });
  double work = 0.0;
  for (auto& B : blocks) work += (double)B.rows.size() * (double)B.cols.size() * (double)std::min(B.rows.size(), B.cols.size());
  const int lanes = (blocks.size() >= 4 && work > 5e7) ? (int)std::min<size_t>(8, blocks.size()) : 1;
  std::vector<int> sweeps(blocks.size(), 0);
  auto do_block = [&](Ctx* c, size_t b) {
    Block& B = blocks[b];
    Res& R = res[b];
    R.r = (int64_t)B.rows.size(); R.c = (int64_t)B.cols.size();
    R.ridx = DevBuf(c, sizeof(int32_t) * R.r); R.cidx = DevBuf(c, sizeof(int32_t) * R.c);
    NSB_CUDA(cudaMemcpyAsync(R.ridx.ptr, B.rows.data(), sizeof(int32_t) * R.r, cudaMemcpyHostToDevice, c->stream));
    NSB_CUDA(cudaMemcpyAsync(R.cidx.ptr, B.cols.data(), sizeof(int32_t) * R.c, cudaMemcpyHostToDevice, c->stream));
    DevBuf Bm(c, sizeof(T) * R.r * R.c);
    gather_block_kernel<T><<<grid1d(c, R.r * R.c), 256, 0, c->stream>>>(M, rows, (const int32_t*)R.ridx.ptr, R.r,
                                                                        (const int32_t*)R.cidx.ptr, R.c, (T*)Bm.ptr);
    LAUNCH_CHECK(c);
    int64_t kb = std::min(R.r, R.c);
    FactorInfo fi = factorize_left<T>(c, (const T*)Bm.ptr, R.r, R.c, R.r, false, 0.0, kb, kb, sqrt_spectrum, R.U, R.C, R.spec);
    sweeps[b] = fi.sweeps;
    P[b] = R.spec;
    c->sync();
  };
  if (lanes <= 1) {
    for (size_t b : order) do_block(ctx, b);
  } else {
    ctx->sync();                                   // M is complete before the helper streams read it
    std::vector<Ctx*> hc(lanes);
    for (int l = 0; l < lanes; ++l) hc[l] = ctx->helper(l);
    std::atomic<size_t> next{0};
    std::vector<std::string> errs(lanes);
    std::vector<int> codes(lanes, 0);
    std::vector<std::thread> th;
    for (int l = 0; l < lanes; ++l)
      th.emplace_back([&, l] {
        cudaSetDevice(ctx->device);
        try {
          for (size_t i = next++; i < order.size(); i = next++) do_block(hc[l], order[i]);
        } catch (const Error& e) { codes[l] = e.code; errs[l] = e.what(); }
        catch (const std::exception& e) { codes[l] = NSB_EINTERNAL; errs[l] = e.what(); }
      });
    for (auto& t : th) t.join();
    for (int l = 0; l < lanes; ++l) {
      ctx->cnt.kernel_launches += hc[l]->cnt.kernel_launches; ctx->cnt.gemm_calls += hc[l]->cnt.gemm_calls;
      ctx->cnt.gemm_flops += hc[l]->cnt.gemm_flops; ctx->cnt.svd_calls += hc[l]->cnt.svd_calls;
      ctx->cnt.jacobi_sweeps += hc[l]->cnt.jacobi_sweeps; ctx->cnt.qr_calls += hc[l]->cnt.qr_calls;
      hc[l]->cnt = Counters();
      if (codes[l] != 0) throw Error(codes[l], errs[l]);
    }
  }
  for (int sw : sweeps) info.sweeps = std::max(info.sweeps, sw);
  double terr = 0.0;
  std::vector<int64_t> keep = truncate_merged(P, cutoff, mindim, maxdim, &terr);
  int64_t ktot = std::accumulate(keep.begin(), keep.end(), (int64_t)0);
  info.newdim = ktot;
  info.truncerr = terr;
  U = DevBuf(ctx, sizeof(T) * rows * ktot);
  C = DevBuf(ctx, sizeof(T) * ktot * cols);
  vec_zero<T>(ctx, rows * ktot, (T*)U.ptr);
  vec_zero<T>(ctx, ktot * cols, (T*)C.ptr);
  new_keys.clear();
  int64_t pos = 0;
  for (size_t b = 0; b < blocks.size(); ++b) {
    int64_t nk = keep[b];
    if (nk == 0) continue;
    Res& R = res[b];
    int64_t kb = std::min(R.r, R.c);
    scatter_rows_kernel<T><<<grid1d(ctx, R.r * nk), 256, 0, ctx->stream>>>((const T*)R.U.ptr, R.r, (const int32_t*)R.ridx.ptr, R.r, nk,
                                                                           (T*)U.ptr, rows, pos);
    LAUNCH_CHECK(ctx);
    scatter_cols_kernel<T><<<grid1d(ctx, nk * R.c), 256, 0, ctx->stream>>>((const T*)R.C.ptr, kb, nk, (const int32_t*)R.cidx.ptr, R.c,
                                                                           (T*)C.ptr, ktot, pos);
    LAUNCH_CHECK(ctx);
    for (int64_t i = 0; i < nk; ++i) new_keys.push_back(blocks[b].key);
    pos += nk;
  }
  ctx->sync();
  return info;
}

template <typename T>
void Net<T>::qr_qn(const T* M, int64_t rows, int64_t cols, const std::vector<int64_t>& rk, const std::vector<int64_t>& ck,
                   DevBuf& Q, DevBuf& Rm, int64_t* kout, std::vector<int64_t>& new_keys) {
  std::vector<Block> blocks = make_blocks(rk, ck);
  int64_t ktot = 0;
  for (auto& B : blocks) ktot += std::min<int64_t>(B.rows.size(), B.cols.size());
  NSB_REQUIRE(ktot > 0, NSB_EINVAL, "qr_qn: tensor has no symmetry-allowed block");
  Q = DevBuf(ctx, sizeof(T) * rows * ktot);
  Rm = DevBuf(ctx, sizeof(T) * ktot * cols);
  vec_zero<T>(ctx, rows * ktot, (T*)Q.ptr);
  vec_zero<T>(ctx, ktot * cols, (T*)Rm.ptr);
  new_keys.clear();
  int64_t pos = 0;
  for (auto& B : blocks) {
    int64_t r = (int64_t)B.rows.size(), c = (int64_t)B.cols.size(), kb = std::min(r, c);
    DevBuf ridx(ctx, sizeof(int32_t) * r), cidx(ctx, sizeof(int32_t) * c), Bm(ctx, sizeof(T) * r * c), Qb(ctx, sizeof(T) * r * kb),
        Rb(ctx, sizeof(T) * kb * c);
    NSB_CUDA(cudaMemcpyAsync(ridx.ptr, B.rows.data(), sizeof(int32_t) * r, cudaMemcpyHostToDevice, ctx->stream));
    NSB_CUDA(cudaMemcpyAsync(cidx.ptr, B.cols.data(), sizeof(int32_t) * c, cudaMemcpyHostToDevice, ctx->stream));
    gather_block_kernel<T><<<grid1d(ctx, r * c), 256, 0, ctx->stream>>>(M, rows, (const int32_t*)ridx.ptr, r, (const int32_t*)cidx.ptr, c,
                                                                        (T*)Bm.ptr);
    LAUNCH_CHECK(ctx);
    qr_thin<T>(ctx, (T*)Bm.ptr, r, c, r, (T*)Qb.ptr, r, (T*)Rb.ptr, kb);
    scatter_rows_kernel<T><<<grid1d(ctx, r * kb), 256, 0, ctx->stream>>>((const T*)Qb.ptr, r, (const int32_t*)ridx.ptr, r, kb, (T*)Q.ptr,
                                                                         rows, pos);
    LAUNCH_CHECK(ctx);
    scatter_cols_kernel<T><<<grid1d(ctx, kb * c), 256, 0, ctx->stream>>>((const T*)Rb.ptr, kb, kb, (const int32_t*)cidx.ptr, c, (T*)Rm.ptr,
                                                                         ktot, pos);
    LAUNCH_CHECK(ctx);
    for (int64_t i = 0; i < kb; ++i) new_keys.push_back(B.key);
    pos += kb;
    ctx->sync();   // index vectors / temporaries go out of scope
  }
  *kout = ktot;
}

// ------------------------------------------------------------------------------------------------
// block-sparse engine: modes, conversions, environment build and H_eff application on symmetry blocks
// ------------------------------------------------------------------------------------------------
template <typename T>
std::shared_ptr<BMode> Net<T>::mode_for(Label l, int64_t dim) {
  const int kind = label_kind(l), id = label_id(l);
  if (kind == LK_SITE) {
    if ((int)site_mode.size() != nverts) site_mode.assign(nverts, nullptr);
    if (!site_mode[id] || site_mode[id]->dim != dim) site_mode[id] = make_mode_small(dim);
    return site_mode[id];
  }
  if (kind == LK_OP) {
    if (op_mode.size() != edges.size()) op_mode.assign(edges.size(), nullptr);
    if (!op_mode[id] || op_mode[id]->dim != dim) op_mode[id] = make_mode_small(dim);
    return op_mode[id];
  }
  NSB_REQUIRE(kind == LK_LINK, NSB_EUNSUPPORTED, "block-sparse engine: auxiliary labels are not sectorised");
  if (link_mode.size() != edges.size()) link_mode.assign(edges.size(), nullptr);
  NSB_REQUIRE((int64_t)qn_link[id].size() == dim * nq, NSB_EINTERNAL, "QN charge table does not match the link dimension");
  if (!link_mode[id] || link_mode[id]->dim != dim) {
    std::vector<int64_t> keys(dim);
    for (int64_t i = 0; i < dim; ++i) keys[i] = pack_key(&qn_link[id][i * nq], nq);
    link_mode[id] = make_mode_from_keys(keys);
  }
  return link_mode[id];
}

template <typename T>
std::vector<std::shared_ptr<BMode>> Net<T>::modes_for(const std::vector<Label>& labels, const std::vector<int64_t>& dims) {
  std::vector<std::shared_ptr<BMode>> m;
  for (size_t i = 0; i < labels.size(); ++i) m.push_back(mode_for(labels[i], dims[i]));
  return m;
}

template <typename T>
BTensor<T> Net<T>::bt_of(const DTensor<T>& t, std::shared_ptr<BStruct> st) {
  if (!bcache) bcache = make_bcache();
  return from_dense<T>(ctx, t, modes_for(t.labels, t.dims), st);
}

template <typename T>
const std::vector<T>& Net<T>::w_host(int v) {
  if ((int)Whost.size() != nverts) Whost.assign(nverts, std::vector<T>());
  if (Whost[v].empty()) {
    Whost[v].resize(W[v].numel());
    NSB_CUDA(cudaMemcpyAsync(Whost[v].data(), W[v].data(), sizeof(T) * W[v].numel(), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
  }
  return Whost[v];
}

// All blocks of a local tensor that charge conservation allows: the charges of the outer subtrees (one per link leg) and of
// the sites add up to the total charge.
template <typename T>
std::shared_ptr<BStruct> Net<T>::allowed_struct(const DTensor<T>& th) { return allowed_struct_in(th, region); }

// Zero every entry of psi[v] that charge conservation forbids (gather the allowed blocks, scatter them back): turns a
// randomly filled tensor into a random symmetric one (synthetic QN states of the benchmarks).
template <typename T>
void Net<T>::qn_project(int v) {
  NSB_REQUIRE(qn_on && v >= 0 && v < nverts && psi[v].valid(), NSB_EINVAL, "qn_project: bad vertex or QN not enabled");
  std::vector<int> inside{v};
  BTensor<T> b = bt_of(psi[v], allowed_struct_in(psi[v], inside));
  psi[v] = to_dense<T>(ctx, b);
  ver[v]++;
}

template <typename T>
std::shared_ptr<BStruct> Net<T>::allowed_struct_in(const DTensor<T>& th, const std::vector<int>& region) {
  auto st = std::make_shared<BStruct>();
  st->modes = modes_for(th.labels, th.dims);
  const int r = th.rank();
  // charge vector of every sector of every mode, as seen from the region
  std::vector<std::vector<std::vector<int64_t>>> sc(r);
  for (int m = 0; m < r; ++m) {
    const Label l = th.labels[m];
    std::vector<int64_t> ch;                                 // per dense state
    if (label_kind(l) == LK_SITE) ch = qn_site[label_id(l)];
    else {
      auto e = edges[label_id(l)];
      const bool first_in = std::find(region.begin(), region.end(), e.first) != region.end();
      const int inside = first_in ? e.first : e.second, outside = first_in ? e.second : e.first;
      ch = side_charge(inside, outside);                     // subtree on the outer vertex's side
    }
    const BMode& md = *st->modes[m];
    sc[m].assign(md.nsec(), std::vector<int64_t>(nq, 0));
    std::vector<char> seen(md.nsec(), 0);
    for (int64_t i = 0; i < md.dim; ++i) {
      const int s = md.state_sector[i];
      if (!seen[s]) { seen[s] = 1; for (int c = 0; c < nq; ++c) sc[m][s][c] = ch[i * nq + c]; }
    }
  }
  std::vector<int32_t> cur(r, 0);
  std::vector<int64_t> sum(nq, 0);
  std::function<void(int)> rec = [&](int m) {
    if (m == r) {
      for (int c = 0; c < nq; ++c) if (sum[c] != qn_total[c]) return;
      st->add_block(cur);
      return;
    }
    for (int s = 0; s < st->modes[m]->nsec(); ++s) {
      cur[m] = s;
      for (int c = 0; c < nq; ++c) sum[c] += sc[m][s][c];
      rec(m + 1);
      for (int c = 0; c < nq; ++c) sum[c] -= sc[m][s][c];
    }
  };
  // mode 0 is the outermost loop here; the block order follows it (any fixed order is fine)
  rec(0);
  st->finalize();
  return st;
}

template <typename T>
const DTensor<T>& Net<T>::env_dense(int u, int v) {
  Env& e = envs.at({u, v});
  env_promote(e);
  if (!e.t.valid() && e.bt.valid()) {
    DTensor<T> d = to_dense<T>(ctx, e.bt);
    e.t.buf = d.buf;
  }
  return e.t;
}

template <typename T>
bool Net<T>::make_env_bt(int u, int v, const std::vector<int>& others, Env* out) {
  if (!bcache) bcache = make_bcache();
  for (int n : others) if (!envs.at({n, u}).bt.valid()) return false;
  BTensor<T> A = bt_of(psi[u]);
  BTensor<T> X = A;
  size_t start = 0;
  if (!others.empty()) {
    X = bcontract<T>(ctx, *bcache, X, envs.at({others[0], u}).bt, false, false, 1);
    if (!X.valid()) return false;
    start = 1;
  }
  {
    std::vector<int> reg{u};
    DTensor<T> fx = fake_dense(X);
    X = bapply_small<T>(ctx, *bcache, X, w_host(u), W[u].labels, W[u].dims, w_out_labels(fx, W[u], u, reg), (uint64_t)(u + 1));
  }
  for (size_t i = start; i < others.size(); ++i) {
    X = bcontract<T>(ctx, *bcache, X, envs.at({others[i], u}).bt, false, false, 1);
    if (!X.valid()) return false;
  }
  BTensor<T> bra = A.primed();
  const std::vector<Label> want{llink(u, v, 0), lop(u, v), llink(u, v, 1)};
  BTensor<T> E = bcontract<T>(ctx, *bcache, bra, X, true, false, 1);
  if (!E.valid() || E.labels != want) E = bcontract<T>(ctx, *bcache, X, bra, false, true, 1);
  if (!E.valid() || E.labels != want) return false;
  out->bt = E;
  out->t = DTensor<T>();
  out->t.labels = want;
  out->t.dims = E.dims();
  return true;
}

template <typename T>
bool Net<T>::apply_heff_bt(const BTensor<T>& x, BTensor<T>* y) {
  const double f0 = ctx->cnt.gemm_flops;
  BTensor<T> X = x;
  for (size_t i = 0; i < plan.size(); ++i) {
    auto& s = plan[i];
    if (s.type == 0) {
      const Env& e = envs.at({s.u, s.v});
      if (!e.bt.valid()) return false;
      X = bcontract<T>(ctx, *bcache, X, e.bt, false, false, 1);
      if (!X.valid()) return false;
    } else if (s.type == 1) {
      DTensor<T> fx = fake_dense(X);
      X = bapply_small<T>(ctx, *bcache, X, w_host(s.v), W[s.v].labels, W[s.v].dims, w_out_labels(fx, W[s.v], s.v, pos), (uint64_t)(s.v + 1));
    } else {
      if (s.Wm_host.empty()) {
        s.Wm_host.resize(s.Wm.numel());
        NSB_CUDA(cudaMemcpyAsync(s.Wm_host.data(), s.Wm.data(), sizeof(T) * s.Wm.numel(), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->sync();
      }
      DTensor<T> fx = fake_dense(X);
      X = bapply_small<T>(ctx, *bcache, X, s.Wm_host, s.Wm.labels, s.Wm.dims, merged_out_labels(fx, s.u, s.v),
                          (uint64_t)(1000000 + (uint64_t)s.u * (uint64_t)nverts + (uint64_t)s.v));
    }
  }
  X = X.noprime();
  if (X.labels != x.labels) return false;
  *y = conform<T>(ctx, *bcache, X, x.st, x.labels);
  ctx->cnt.matvecs++;
  bt_last_apply_flops = ctx->cnt.gemm_flops - f0;
  return true;
}

#define INST(T)                                                                                                          \
  template std::shared_ptr<BMode> Net<T>::mode_for(Label, int64_t);                                                      \
  template std::vector<std::shared_ptr<BMode>> Net<T>::modes_for(const std::vector<Label>&, const std::vector<int64_t>&); \
  template BTensor<T> Net<T>::bt_of(const DTensor<T>&, std::shared_ptr<BStruct>);                                        \
  template const std::vector<T>& Net<T>::w_host(int);                                                                    \
  template std::shared_ptr<BStruct> Net<T>::allowed_struct(const DTensor<T>&);                                           \
  template std::shared_ptr<BStruct> Net<T>::allowed_struct_in(const DTensor<T>&, const std::vector<int>&);               \
  template void Net<T>::qn_project(int);                                                                                  \
  template const DTensor<T>& Net<T>::env_dense(int, int);                                                                \
  template bool Net<T>::make_env_bt(int, int, const std::vector<int>&, Env*);                                            \
  template bool Net<T>::apply_heff_bt(const BTensor<T>&, BTensor<T>*);                                                   \
  template void Net<T>::qn_enable(int, const int32_t*);                                                                  \
  template void Net<T>::qn_set_site(int, const int32_t*);                                                                \
  template void Net<T>::qn_set_link(int, int, const int32_t*);                                                           \
  template void Net<T>::qn_get_link(int, int, int32_t*);                                                                 \
  template std::vector<int64_t> Net<T>::side_charge(int, int) const;                                                     \
  template std::vector<int64_t> Net<T>::leg_charges(int, Label) const;                                                   \
  template std::vector<int64_t> Net<T>::multi_keys(int, const std::vector<Label>&, const std::vector<int64_t>&, bool) const; \
  template void Net<T>::qn_store_link(int, int, const std::vector<int64_t>&);                                            \
  template FactorInfo Net<T>::factorize_qn(const T*, int64_t, int64_t, const std::vector<int64_t>&, const std::vector<int64_t>&, \
                                           double, int64_t, int64_t, bool, DevBuf&, DevBuf&, std::vector<int64_t>&);     \
  template void Net<T>::qr_qn(const T*, int64_t, int64_t, const std::vector<int64_t>&, const std::vector<int64_t>&, DevBuf&, DevBuf&, \
                              int64_t*, std::vector<int64_t>&);
INST(double)
INST(cdouble)

}  // namespace nsb
