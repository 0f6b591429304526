// Abelian quantum-number conservation (dense storage, block-wise factorisations).
//
// Counterpart of running the reference on QN-conserving ITensors (`siteinds(...; conserve_qns=true)`): the three
// factorisations on the path -- the inserter's truncating factorisation, the expansion's `eigen`, the gauge `qr` --
// never mix symmetry sectors and the truncation acts on the merged spectrum (NDTensors rule).
// Tensors stay dense: entries forbidden by symmetry are exact zeros, and contractions of symmetric tensors keep them
// exact, so the matvec / environment kernels are unchanged.  Charge convention: see oracle/qn.py.
#include "net.h"

#include <algorithm>
#include <numeric>

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
  for (size_t b = 0; b < blocks.size(); ++b) {
    Block& B = blocks[b];
    Res& R = res[b];
    R.r = (int64_t)B.rows.size(); R.c = (int64_t)B.cols.size();
    R.ridx = DevBuf(ctx, sizeof(int32_t) * R.r); R.cidx = DevBuf(ctx, sizeof(int32_t) * R.c);
    NSB_CUDA(cudaMemcpyAsync(R.ridx.ptr, B.rows.data(), sizeof(int32_t) * R.r, cudaMemcpyHostToDevice, ctx->stream));
    NSB_CUDA(cudaMemcpyAsync(R.cidx.ptr, B.cols.data(), sizeof(int32_t) * R.c, cudaMemcpyHostToDevice, ctx->stream));
    DevBuf Bm(ctx, sizeof(T) * R.r * R.c);
    gather_block_kernel<T><<<grid1d(ctx, R.r * R.c), 256, 0, ctx->stream>>>(M, rows, (const int32_t*)R.ridx.ptr, R.r,
                                                                            (const int32_t*)R.cidx.ptr, R.c, (T*)Bm.ptr);
    LAUNCH_CHECK(ctx);
    int64_t kb = std::min(R.r, R.c);
    FactorInfo fi = factorize_left<T>(ctx, (const T*)Bm.ptr, R.r, R.c, R.r, false, 0.0, kb, kb, sqrt_spectrum, R.U, R.C, R.spec);
    info.sweeps = std::max(info.sweeps, fi.sweeps);
    P[b] = R.spec;
  }
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

#define INST(T)                                                                                                          \
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
