// Block-sparse tensors and the sector-batched contraction engine (see bsparse.h).
#include "bsparse.h"

#include <algorithm>
#include <atomic>
#include <tuple>

namespace nsb {

#define LAUNCH_CHECK(ctx) do { (ctx)->cnt.kernel_launches++; NSB_CUDA(cudaGetLastError()); } while (0)

static std::atomic<uint64_t> g_next_id{1};

std::shared_ptr<BMode> make_mode_small(int64_t dim) {
  auto m = std::make_shared<BMode>();
  m->dim = dim; m->small = true;
  m->key.resize(dim); m->sdim.assign(dim, 1); m->soff.resize(dim); m->state_sector.resize(dim); m->state_pos.assign(dim, 0);
  for (int64_t i = 0; i < dim; ++i) { m->key[i] = i; m->soff[i] = i; m->state_sector[i] = (int32_t)i; }
  m->id = g_next_id++;
  return m;
}

std::shared_ptr<BMode> make_mode_from_keys(const std::vector<int64_t>& sk) {
  auto m = std::make_shared<BMode>();
  m->dim = (int64_t)sk.size();
  m->state_sector.resize(sk.size()); m->state_pos.resize(sk.size());
  std::map<int64_t, int> where;
  for (size_t i = 0; i < sk.size(); ++i) {
    auto it = where.find(sk[i]);
    if (it == where.end()) { it = where.emplace(sk[i], (int)m->key.size()).first; m->key.push_back(sk[i]); m->sdim.push_back(0); }
    m->state_sector[i] = it->second;
    m->state_pos[i] = (int32_t)m->sdim[it->second]++;
  }
  m->soff.resize(m->key.size());
  int64_t o = 0;
  for (size_t s = 0; s < m->key.size(); ++s) { m->soff[s] = o; o += m->sdim[s]; }
  m->id = g_next_id++;
  return m;
}

static bool same_mode(const BMode& a, const BMode& b) {
  if (a.id == b.id) return true;
  return a.dim == b.dim && a.key == b.key && a.sdim == b.sdim && a.state_sector == b.state_sector;
}

void BStruct::add_block(const std::vector<int32_t>& s) {
  if (index.count(s)) return;
  Blk b; b.s = s;
  index[s] = (int)blocks.size();
  blocks.push_back(b);
}
void BStruct::finalize() {
  int64_t o = 0;
  for (auto& b : blocks) {
    int64_t n = 1;
    for (size_t m = 0; m < modes.size(); ++m) n *= modes[m]->sdim[b.s[m]];
    b.off = o; b.numel = n;
    o += (n + 1) & ~(int64_t)1;     // even offsets: 16-byte aligned real blocks whenever their leading dimension allows it
  }
  total = o;
  id = g_next_id++;
}

// ------------------------------------------------------------------------------------------------
// gather / scatter between dense storage and blocks
// ------------------------------------------------------------------------------------------------
struct BMap {
  int rank;
  int64_t dim[MAX_RANK], cstride[MAX_RANK];
  int64_t off_sector[MAX_RANK], off_pos[MAX_RANK], off_sdim[MAX_RANK], off_cand;
};

static BMap build_tables(Ctx* ctx, BStruct& st) {
  BMap m{};
  m.rank = st.rank();
  NSB_REQUIRE(m.rank <= MAX_RANK, NSB_EINTERNAL, "block tensor: rank too large");
  std::vector<int64_t> tab;
  int64_t ncand = 1;
  for (int k = 0; k < m.rank; ++k) {
    const BMode& md = *st.modes[k];
    m.dim[k] = md.dim;
    m.cstride[k] = ncand;
    ncand *= md.nsec();
    NSB_REQUIRE(ncand < (1ll << 26), NSB_EUNSUPPORTED, "block tensor: too many sector combinations");
    m.off_sector[k] = (int64_t)tab.size();
    for (auto s : md.state_sector) tab.push_back(s);
    m.off_pos[k] = (int64_t)tab.size();
    for (auto p : md.state_pos) tab.push_back(p);
    m.off_sdim[k] = (int64_t)tab.size();
    for (auto d : md.sdim) tab.push_back(d);
  }
  m.off_cand = (int64_t)tab.size();
  tab.resize(tab.size() + ncand, -1);
  for (auto& b : st.blocks) {
    int64_t c = 0;
    for (int k = 0; k < m.rank; ++k) c += (int64_t)b.s[k] * m.cstride[k];
    tab[m.off_cand + c] = b.off;
  }
  st.ncand = ncand;
  st.d_tables = DevBuf(ctx, sizeof(int64_t) * tab.size());
  NSB_CUDA(cudaMemcpyAsync(st.d_tables.ptr, tab.data(), sizeof(int64_t) * tab.size(), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
  st.tables_ready = true;
  return m;
}

static BMap map_of(Ctx* ctx, BStruct& st) {
  // (re)derive the small header; tables are uploaded once per structure
  if (!st.tables_ready) return build_tables(ctx, st);
  BMap m{};
  m.rank = st.rank();
  int64_t pos = 0, ncand = 1;
  for (int k = 0; k < m.rank; ++k) {
    const BMode& md = *st.modes[k];
    m.dim[k] = md.dim; m.cstride[k] = ncand; ncand *= md.nsec();
    m.off_sector[k] = pos; pos += md.dim;
    m.off_pos[k] = pos; pos += md.dim;
    m.off_sdim[k] = pos; pos += md.nsec();
  }
  m.off_cand = pos;
  return m;
}

// mode 0: flat[block] = dense;  1: dense = flat[block] or 0;  2: flags[cand] = 1 where dense != 0
template <typename T, int MODE>
__global__ void __launch_bounds__(256) bmap_kernel(BMap m, const int64_t* __restrict__ tab, T* __restrict__ dense, T* __restrict__ flat,
                                                   int32_t* __restrict__ flags, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (MODE == 2) {
      const T v = dense[i];
      if (re(v) == 0.0 && im(v) == 0.0) continue;
    }
    int64_t r = i, cand = 0, local = 0, lstride = 1;
    for (int k = 0; k < m.rank; ++k) {
      const int64_t idx = r % m.dim[k];
      r /= m.dim[k];
      const int64_t sec = tab[m.off_sector[k] + idx];
      cand += sec * m.cstride[k];
      local += tab[m.off_pos[k] + idx] * lstride;
      lstride *= tab[m.off_sdim[k] + sec];
    }
    if (MODE == 2) { flags[cand] = 1; continue; }
    const int64_t boff = tab[m.off_cand + cand];
    if (MODE == 0) { if (boff >= 0) flat[boff + local] = dense[i]; }
    else dense[i] = (boff >= 0) ? flat[boff + local] : zero_<T>();
  }
}

static int grid_for_n(Ctx* ctx, int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8)); }

template <typename T>
BTensor<T> from_dense(Ctx* ctx, const DTensor<T>& t, const std::vector<std::shared_ptr<BMode>>& modes, std::shared_ptr<BStruct> st) {
  NSB_REQUIRE((int)modes.size() == t.rank(), NSB_EINTERNAL, "from_dense: rank mismatch");
  for (int k = 0; k < t.rank(); ++k) NSB_REQUIRE(modes[k]->dim == t.dims[k], NSB_EINTERNAL, "from_dense: mode dimension mismatch");
  const int64_t total = t.numel();
  if (!st) {
    // detect the non-vanishing blocks
    BStruct probe;
    probe.modes = modes;
    probe.finalize();
    BMap m = build_tables(ctx, probe);
    DevBuf flags(ctx, sizeof(int32_t) * probe.ncand);
    NSB_CUDA(cudaMemsetAsync(flags.ptr, 0, sizeof(int32_t) * probe.ncand, ctx->stream));
    bmap_kernel<T, 2><<<grid_for_n(ctx, total), 256, 0, ctx->stream>>>(m, (const int64_t*)probe.d_tables.ptr, t.data(), nullptr, (int32_t*)flags.ptr, total);
    LAUNCH_CHECK(ctx);
    std::vector<int32_t> hf(probe.ncand);
    NSB_CUDA(cudaMemcpyAsync(hf.data(), flags.ptr, sizeof(int32_t) * probe.ncand, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    st = std::make_shared<BStruct>();
    st->modes = modes;
    for (int64_t c = 0; c < probe.ncand; ++c) {
      if (!hf[c]) continue;
      std::vector<int32_t> s(modes.size());
      int64_t r = c;
      for (size_t k = 0; k < modes.size(); ++k) { s[k] = (int32_t)(r % modes[k]->nsec()); r /= modes[k]->nsec(); }
      st->add_block(s);
    }
    st->finalize();
  }
  BTensor<T> out;
  out.st = st;
  out.labels = t.labels;
  out.buf = std::make_shared<DevBuf>(ctx, sizeof(T) * (size_t)std::max<int64_t>(st->total, 1));
  vec_zero<T>(ctx, st->total, out.data());      // (padding between odd-sized blocks stays zero: flat dot products see no garbage)
  BMap m = map_of(ctx, *st);
  bmap_kernel<T, 0><<<grid_for_n(ctx, total), 256, 0, ctx->stream>>>(m, (const int64_t*)st->d_tables.ptr, t.data(), out.data(), nullptr, total);
  LAUNCH_CHECK(ctx);
  return out;
}

template <typename T>
DTensor<T> to_dense(Ctx* ctx, const BTensor<T>& b) {
  DTensor<T> out(ctx, b.dims(), b.labels);
  BMap m = map_of(ctx, *b.st);
  const int64_t total = out.numel();
  bmap_kernel<T, 1><<<grid_for_n(ctx, total), 256, 0, ctx->stream>>>(m, (const int64_t*)b.st->d_tables.ptr, out.data(), b.data(), nullptr, total);
  LAUNCH_CHECK(ctx);
  return out;
}

// ------------------------------------------------------------------------------------------------
// grouped block linear combinations: dst_i = sum_s coef_s * src_s   (bapply_small, conform)
// ------------------------------------------------------------------------------------------------
struct LItem { int64_t dst_off, n; int32_t s0, ns; };
struct LSrc { int64_t src_off; double cr, ci; };

template <typename T>
__global__ void __launch_bounds__(256) block_lincomb_kernel(const LItem* __restrict__ items, int nitems, const LSrc* __restrict__ srcs,
                                                            const T* __restrict__ sbase, T* __restrict__ dbase) {
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
    const LItem I = items[it];
    T* dst = dbase + I.dst_off;
    for (int64_t e = threadIdx.x; e < I.n; e += blockDim.x) {
      T acc = zero_<T>();
      for (int s = I.s0; s < I.s0 + I.ns; ++s) {
        const LSrc S = srcs[s];
        fma_(acc, from_complex<T>(S.cr, S.ci), sbase[S.src_off + e]);
      }
      dst[e] = acc;
    }
  }
}

struct LincombPlan {
  std::shared_ptr<BStruct> out;
  std::vector<Label> out_labels;
  DevBuf items, srcs;
  int nitems = 0;
  double flops = 0.0;
};

static void upload_lincomb(Ctx* ctx, LincombPlan& p, const std::vector<LItem>& items0, const std::vector<LSrc>& srcs0) {
  // chunk large blocks so that one block does not serialise on one CTA
  const int64_t CH = 16384;
  std::vector<LItem> items;
  std::vector<LSrc> srcs;
  for (auto& I : items0) {
    for (int64_t c = 0; c < std::max<int64_t>(I.n, 1); c += CH) {
      LItem J; J.dst_off = I.dst_off + c; J.n = std::min(CH, I.n - c); J.s0 = (int32_t)srcs.size(); J.ns = I.ns;
      for (int s = I.s0; s < I.s0 + I.ns; ++s) { LSrc S = srcs0[s]; S.src_off += c; srcs.push_back(S); }
      if (J.n > 0) items.push_back(J);
    }
  }
  p.nitems = (int)items.size();
  p.items = DevBuf(ctx, sizeof(LItem) * std::max<size_t>(items.size(), 1));
  p.srcs = DevBuf(ctx, sizeof(LSrc) * std::max<size_t>(srcs.size(), 1));
  if (!items.empty()) NSB_CUDA(cudaMemcpyAsync(p.items.ptr, items.data(), sizeof(LItem) * items.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (!srcs.empty()) NSB_CUDA(cudaMemcpyAsync(p.srcs.ptr, srcs.data(), sizeof(LSrc) * srcs.size(), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
}

template <typename T>
static BTensor<T> run_lincomb(Ctx* ctx, const LincombPlan& p, const BTensor<T>& X) {
  BTensor<T> out;
  out.st = p.out;
  out.labels = p.out_labels;
  out.buf = std::make_shared<DevBuf>(ctx, sizeof(T) * (size_t)std::max<int64_t>(p.out->total, 1));
  vec_zero<T>(ctx, p.out->total, out.data());
  if (p.nitems > 0) {
    const int grid = std::min(p.nitems, ctx->num_sms * 8);
    block_lincomb_kernel<T><<<grid, 256, 0, ctx->stream>>>((const LItem*)p.items.ptr, p.nitems, (const LSrc*)p.srcs.ptr, X.data(), out.data());
    LAUNCH_CHECK(ctx);
  }
  return out;
}

// ------------------------------------------------------------------------------------------------
// plan cache
// ------------------------------------------------------------------------------------------------
struct ContractGroup {
  int opa = 0, opb = 0;
  bool a_from_x = true;     // GEMM operand A comes from X (the operand that keeps its mode order), B from Y; else swapped
  DevBuf probs, segs;
  int nprob = 0;
  int64_t total_tiles = 0;
  double flops = 0.0;
};
struct ContractPlanB {
  bool direct = false, x_is_a = true;
  std::shared_ptr<BStruct> out;
  std::vector<Label> out_labels;
  ContractGroup grp[2];
};
typedef std::tuple<uint64_t, uint64_t, std::vector<Label>, std::vector<Label>, int> CKey;
typedef std::tuple<uint64_t, uint64_t, std::vector<Label>> LKey;
struct BCache {
  std::map<CKey, std::shared_ptr<ContractPlanB>> contract;
  std::map<LKey, std::shared_ptr<LincombPlan>> lincomb;
};
std::shared_ptr<BCache> make_bcache() { return std::make_shared<BCache>(); }

template <typename T>
BTensor<T> conform(Ctx* ctx, BCache& cache, const BTensor<T>& x, std::shared_ptr<BStruct> st, const std::vector<Label>& labels) {
  NSB_REQUIRE(x.labels == labels, NSB_EINTERNAL, "conform: label order differs");
  if (x.st->id == st->id) return x;
  LKey key(x.st->id, st->id | (1ull << 63), labels);
  auto itc = cache.lincomb.find(key);
  if (itc != cache.lincomb.end()) return run_lincomb<T>(ctx, *itc->second, x);
  NSB_REQUIRE(x.st->rank() == st->rank(), NSB_EINTERNAL, "conform: rank mismatch");
  for (int k = 0; k < st->rank(); ++k) NSB_REQUIRE(same_mode(*x.st->modes[k], *st->modes[k]), NSB_EINTERNAL, "conform: sectorisation differs");
  auto pp = std::make_shared<LincombPlan>();
  LincombPlan& p = *pp;
  p.out = st; p.out_labels = labels;
  std::vector<LItem> items;
  std::vector<LSrc> srcs;
  for (auto& b : st->blocks) {
    auto it = x.st->index.find(b.s);
    if (it == x.st->index.end()) continue;
    LItem I; I.dst_off = b.off; I.n = b.numel; I.s0 = (int32_t)srcs.size(); I.ns = 1;
    srcs.push_back(LSrc{x.st->blocks[it->second].off, 1.0, 0.0});
    items.push_back(I);
  }
  upload_lincomb(ctx, p, items, srcs);
  if (cache.lincomb.size() > 4096) cache.lincomb.clear();
  cache.lincomb[key] = pp;
  return run_lincomb<T>(ctx, p, x);
}

// ------------------------------------------------------------------------------------------------
// bcontract
// ------------------------------------------------------------------------------------------------
namespace {
struct LabelInfo {
  std::vector<int> posX, posY;
  bool x_contig = false, same_order = false, y_prefix = false, y_suffix = false;
  int x_start = 0;
};
LabelInfo analyze_labels(const std::vector<Label>& X, const std::vector<Label>& Y) {
  LabelInfo b;
  for (int i = 0; i < (int)X.size(); ++i)
    for (int j = 0; j < (int)Y.size(); ++j)
      if (X[i] == Y[j]) { b.posX.push_back(i); b.posY.push_back(j); }
  const int k = (int)b.posX.size();
  if (k == 0) { b.x_contig = b.same_order = b.y_prefix = true; b.x_start = (int)X.size(); return b; }
  b.x_contig = true;
  for (int i = 1; i < k; ++i) if (b.posX[i] != b.posX[i - 1] + 1) b.x_contig = false;
  b.x_start = b.posX[0];
  b.same_order = true;
  for (int i = 1; i < k; ++i) if (b.posY[i] != b.posY[i - 1] + 1) b.same_order = false;
  b.y_prefix = b.same_order && b.posY[0] == 0;
  b.y_suffix = b.same_order && b.posY[k - 1] == (int)Y.size() - 1;
  return b;
}
bool direct_ok(const LabelInfo& b) { return b.x_contig && b.same_order && (b.y_prefix || b.y_suffix); }
}  // namespace

template <typename T>
static std::shared_ptr<ContractPlanB> build_contract_plan(Ctx* ctx, const BTensor<T>& A, const BTensor<T>& B, bool conjA, bool conjB, int prefer_x) {
  auto plan = std::make_shared<ContractPlanB>();
  // operand roles as in contract(): X keeps its mode order, the shared block of Y is a prefix or suffix
  bool x_is_a = true, found = false;
  LabelInfo li;
  for (int attempt = 0; attempt < 2 && !found; ++attempt) {
    const bool try_a = (prefer_x == 2) ? (attempt == 1) : (attempt == 0);
    li = try_a ? analyze_labels(A.labels, B.labels) : analyze_labels(B.labels, A.labels);
    if (direct_ok(li)) { x_is_a = try_a; found = true; }
  }
  if (!found) return plan;     // direct == false
  plan->direct = true;
  plan->x_is_a = x_is_a;
  const BTensor<T>& X = x_is_a ? A : B;
  const BTensor<T>& Y = x_is_a ? B : A;
  const bool conjX = x_is_a ? conjA : conjB, conjY = x_is_a ? conjB : conjA;
  const int k = (int)li.posX.size(), rx = X.rank(), ry = Y.rank();
  for (int i = 0; i < k; ++i)
    NSB_REQUIRE(same_mode(*X.st->modes[li.posX[i]], *Y.st->modes[li.posY[i]]), NSB_EINTERNAL, "bcontract: sectorisations of a shared link differ");
  const bool yprefix = (k == 0) ? true : li.y_prefix;
  // output modes: X before the block, Y free modes, X after the block
  std::vector<int> yfree;
  for (int j = 0; j < ry; ++j) if (std::find(li.posY.begin(), li.posY.end(), j) == li.posY.end()) yfree.push_back(j);
  auto out = std::make_shared<BStruct>();
  for (int i = 0; i < li.x_start; ++i) { out->modes.push_back(X.st->modes[i]); plan->out_labels.push_back(X.labels[i]); }
  for (int j : yfree) { out->modes.push_back(Y.st->modes[j]); plan->out_labels.push_back(Y.labels[j]); }
  for (int i = li.x_start + k; i < rx; ++i) { out->modes.push_back(X.st->modes[i]); plan->out_labels.push_back(X.labels[i]); }
  // Y blocks by their shared-sector tuple
  std::map<std::vector<int32_t>, std::vector<int>> ybys;
  for (int by = 0; by < (int)Y.st->blocks.size(); ++by) {
    std::vector<int32_t> key(k);
    for (int i = 0; i < k; ++i) key[i] = Y.st->blocks[by].s[li.posY[i]];
    ybys[key].push_back(by);
  }
  struct Pair { int bx, by; };
  std::map<std::vector<int32_t>, std::vector<Pair>> byout;
  std::vector<std::vector<int32_t>> order;      // output blocks in order of first appearance (deterministic)
  for (int bx = 0; bx < (int)X.st->blocks.size(); ++bx) {
    std::vector<int32_t> key(k);
    for (int i = 0; i < k; ++i) key[i] = X.st->blocks[bx].s[li.posX[i]];
    auto it = ybys.find(key);
    if (it == ybys.end()) continue;
    for (int by : it->second) {
      std::vector<int32_t> os;
      for (int i = 0; i < li.x_start; ++i) os.push_back(X.st->blocks[bx].s[i]);
      for (int j : yfree) os.push_back(Y.st->blocks[by].s[j]);
      for (int i = li.x_start + k; i < rx; ++i) os.push_back(X.st->blocks[bx].s[i]);
      auto f = byout.find(os);
      if (f == byout.end()) { order.push_back(os); f = byout.emplace(os, std::vector<Pair>()).first; }
      f->second.push_back(Pair{bx, by});
    }
  }
  for (auto& os : order) out->add_block(os);
  out->finalize();
  plan->out = out;
  // GEMM tables: group 0 = "X as A" form (P > 1 or Q == 1), group 1 = "Y as A" form (P == 1)
  const bool cplx = ScalarTraits<T>::is_complex;
  const int BM = gemm_tile_m(cplx), BN = gemm_tile_n(cplx);
  std::vector<GroupedProblem> probs[2];
  std::vector<GroupedSegment> segs[2];
  int64_t tiles[2] = {0, 0};
  double flops[2] = {0.0, 0.0};
  const double fpm = cplx ? 8.0 : 2.0;
  auto aligned = [&](int64_t off, int64_t ld) { return cplx || ((off % 2 == 0) && (ld % 2 == 0)); };
  for (auto& os : order) {
    const BStruct::Blk& ob = out->blocks[out->index.at(os)];
    auto& pairs = byout.at(os);
    // block extents (identical for all pairs of this output block except Kc)
    const BStruct::Blk& x0 = X.st->blocks[pairs[0].bx];
    const BStruct::Blk& y0 = Y.st->blocks[pairs[0].by];
    int64_t P = 1, Q = 1, N = 1;
    for (int i = 0; i < li.x_start; ++i) P *= X.st->modes[i]->sdim[x0.s[i]];
    for (int i = li.x_start + k; i < rx; ++i) Q *= X.st->modes[i]->sdim[x0.s[i]];
    for (int j : yfree) N *= Y.st->modes[j]->sdim[y0.s[j]];
    const int gsel = (P == 1 && Q > 1) ? 1 : 0;
    const int64_t nq = (gsel == 0) ? Q : 1;
    for (int64_t q = 0; q < nq; ++q) {
      GroupedProblem pr{};
      pr.seg0 = (int32_t)segs[gsel].size();
      if (gsel == 0) { pr.M = P; pr.N = N; pr.ldc = P; pr.c_off = ob.off + q * P * N; }
      else { pr.M = N; pr.N = Q; pr.ldc = N; pr.c_off = ob.off; }
      for (auto& pp : pairs) {
        const BStruct::Blk& xb = X.st->blocks[pp.bx];
        const BStruct::Blk& yb = Y.st->blocks[pp.by];
        int64_t Kc = 1;
        for (int i = 0; i < k; ++i) Kc *= X.st->modes[li.posX[i]]->sdim[xb.s[li.posX[i]]];
        GroupedSegment sg{};
        sg.K = Kc;
        if (gsel == 0) {        // Out_q[P, N] = X_q[P, Kc] Ymat
          sg.a_off = xb.off + q * P * Kc; sg.lda = P;
          sg.b_off = yb.off; sg.ldb = yprefix ? Kc : N;
        } else {                // Out[N, Q] = Ymat^T X[Kc, Q]
          sg.a_off = yb.off; sg.lda = yprefix ? Kc : N;
          sg.b_off = xb.off; sg.ldb = Kc;
        }
        sg.alignedA = aligned(sg.a_off, sg.lda) ? 1 : 0;
        sg.alignedB = aligned(sg.b_off, sg.ldb) ? 1 : 0;
        segs[gsel].push_back(sg);
        flops[gsel] += fpm * (double)pr.M * (double)pr.N * (double)Kc;
      }
      pr.nseg = (int32_t)segs[gsel].size() - pr.seg0;
      pr.tiles_m = (pr.M + BM - 1) / BM;
      pr.tile0 = tiles[gsel];
      tiles[gsel] += pr.tiles_m * ((pr.N + BN - 1) / BN);
      probs[gsel].push_back(pr);
    }
  }
  for (int gsel = 0; gsel < 2; ++gsel) {
    ContractGroup& G = plan->grp[gsel];
    G.nprob = (int)probs[gsel].size();
    G.total_tiles = tiles[gsel];
    G.flops = flops[gsel];
    if (gsel == 0) {
      G.a_from_x = true;
      G.opa = conjX ? OP_CONJ : OP_N;
      G.opb = yprefix ? (conjY ? OP_CONJ : OP_N) : (conjY ? OP_C : OP_T);
    } else {
      G.a_from_x = false;
      G.opa = yprefix ? (conjY ? OP_C : OP_T) : (conjY ? OP_CONJ : OP_N);
      G.opb = conjX ? OP_CONJ : OP_N;
    }
    if (G.nprob == 0) continue;
    G.probs = DevBuf(ctx, sizeof(GroupedProblem) * probs[gsel].size());
    G.segs = DevBuf(ctx, sizeof(GroupedSegment) * segs[gsel].size());
    NSB_CUDA(cudaMemcpyAsync(G.probs.ptr, probs[gsel].data(), sizeof(GroupedProblem) * probs[gsel].size(), cudaMemcpyHostToDevice, ctx->stream));
    NSB_CUDA(cudaMemcpyAsync(G.segs.ptr, segs[gsel].data(), sizeof(GroupedSegment) * segs[gsel].size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->sync();
  return plan;
}

template <typename T>
BTensor<T> bcontract(Ctx* ctx, BCache& cache, const BTensor<T>& A, const BTensor<T>& B, bool conjA, bool conjB, int prefer_x) {
  CKey key(A.st->id, B.st->id, A.labels, B.labels, (conjA ? 1 : 0) | (conjB ? 2 : 0) | (prefer_x << 2));
  auto it = cache.contract.find(key);
  std::shared_ptr<ContractPlanB> plan;
  if (it != cache.contract.end()) plan = it->second;
  else {
    if (cache.contract.size() > 4096) cache.contract.clear();
    plan = build_contract_plan<T>(ctx, A, B, conjA, conjB, prefer_x);
    cache.contract[key] = plan;
  }
  BTensor<T> out;
  if (!plan->direct) return out;
  out.st = plan->out;
  out.labels = plan->out_labels;
  out.buf = std::make_shared<DevBuf>(ctx, sizeof(T) * (size_t)std::max<int64_t>(plan->out->total, 1));
  vec_zero<T>(ctx, plan->out->total, out.data());     // padding elements stay zero
  const BTensor<T>& X = plan->x_is_a ? A : B;
  const BTensor<T>& Y = plan->x_is_a ? B : A;
  for (int gsel = 0; gsel < 2; ++gsel) {
    const ContractGroup& G = plan->grp[gsel];
    if (G.nprob == 0) continue;
    gemm_grouped<T>(ctx, G.opa, G.opb, (const GroupedProblem*)G.probs.ptr, G.nprob, (const GroupedSegment*)G.segs.ptr, G.total_tiles,
                    G.a_from_x ? X.data() : Y.data(), G.a_from_x ? Y.data() : X.data(), out.data(), G.flops);
  }
  return out;
}

// ------------------------------------------------------------------------------------------------
// bapply_small
// ------------------------------------------------------------------------------------------------
template <typename T>
BTensor<T> bapply_small(Ctx* ctx, BCache& cache, const BTensor<T>& X, const std::vector<T>& Wh, const std::vector<Label>& wlabels,
                        const std::vector<int64_t>& wdims, const std::vector<Label>& out_labels, uint64_t op_id) {
  LKey key(X.st->id, op_id, out_labels);
  auto it = cache.lincomb.find(key);
  std::shared_ptr<LincombPlan> plan;
  if (it != cache.lincomb.end()) plan = it->second;
  else {
    if (cache.lincomb.size() > 4096) cache.lincomb.clear();
    plan = std::make_shared<LincombPlan>();
    const int rx = X.rank(), rw = (int)wlabels.size();
    std::vector<int> wpos_of_x(rx, -1);                 // contracted modes of X -> position in W
    for (int i = 0; i < rx; ++i)
      for (int j = 0; j < rw; ++j)
        if (X.labels[i] == wlabels[j]) wpos_of_x[i] = j;
    std::vector<int64_t> wstride(rw);
    { int64_t s = 1; for (int j = 0; j < rw; ++j) { wstride[j] = s; s *= wdims[j]; } }
    // output modes
    auto out = std::make_shared<BStruct>();
    std::vector<int> src_of_out(out_labels.size(), -1), w_of_out(out_labels.size(), -1);
    int last_kept = -1;
    for (size_t o = 0; o < out_labels.size(); ++o) {
      int ix = X.find(out_labels[o]);
      if (ix >= 0) {
        NSB_REQUIRE(wpos_of_x[ix] < 0, NSB_EINTERNAL, "bapply_small: kept label also in the operator");
        if (!X.st->modes[ix]->small) { NSB_REQUIRE(ix > last_kept, NSB_EINTERNAL, "bapply_small: big modes must keep their order"); last_kept = ix; }
        src_of_out[o] = ix;
        out->modes.push_back(X.st->modes[ix]);
      } else {
        int jw = -1;
        for (int j = 0; j < rw; ++j) if (wlabels[j] == out_labels[o]) jw = j;
        NSB_REQUIRE(jw >= 0, NSB_EINTERNAL, "bapply_small: unknown output label");
        w_of_out[o] = jw;
        out->modes.push_back(make_mode_small(wdims[jw]));
      }
    }
    std::vector<int> newmodes;     // W positions of the new modes
    for (size_t o = 0; o < out_labels.size(); ++o) if (w_of_out[o] >= 0) newmodes.push_back(w_of_out[o]);
    for (int i = 0; i < rx; ++i) if (wpos_of_x[i] >= 0) NSB_REQUIRE(X.st->modes[i]->small, NSB_EUNSUPPORTED, "bapply_small: contracted mode is not small");
    int64_t nnew = 1;
    for (int j : newmodes) nnew *= wdims[j];
    struct Contrib { int bx; T coef; };
    std::map<std::vector<int32_t>, std::vector<Contrib>> byout;
    std::vector<std::vector<int32_t>> order;
    for (int bx = 0; bx < (int)X.st->blocks.size(); ++bx) {
      const BStruct::Blk& xb = X.st->blocks[bx];
      int64_t wbase = 0;
      for (int i = 0; i < rx; ++i) if (wpos_of_x[i] >= 0) wbase += (int64_t)xb.s[i] * wstride[wpos_of_x[i]];   // small mode: sector == state
      for (int64_t nn = 0; nn < nnew; ++nn) {
        int64_t r = nn, woff = wbase;
        std::vector<int64_t> nstate(rw, 0);
        for (int j : newmodes) { nstate[j] = r % wdims[j]; r /= wdims[j]; woff += nstate[j] * wstride[j]; }
        const T wv = Wh[woff];
        if (re(wv) == 0.0 && im(wv) == 0.0) continue;
        std::vector<int32_t> os(out_labels.size());
        for (size_t o = 0; o < out_labels.size(); ++o) os[o] = (src_of_out[o] >= 0) ? xb.s[src_of_out[o]] : (int32_t)nstate[w_of_out[o]];
        auto f = byout.find(os);
        if (f == byout.end()) { order.push_back(os); f = byout.emplace(os, std::vector<Contrib>()).first; }
        f->second.push_back(Contrib{bx, wv});
      }
    }
    for (auto& os : order) out->add_block(os);
    out->finalize();
    plan->out = out;
    plan->out_labels = out_labels;
    std::vector<LItem> items;
    std::vector<LSrc> srcs;
    for (auto& os : order) {
      const BStruct::Blk& ob = out->blocks[out->index.at(os)];
      LItem I; I.dst_off = ob.off; I.n = ob.numel; I.s0 = (int32_t)srcs.size();
      for (auto& c : byout.at(os)) {
        NSB_REQUIRE(X.st->blocks[c.bx].numel == ob.numel, NSB_EINTERNAL, "bapply_small: block shapes differ");
        srcs.push_back(LSrc{X.st->blocks[c.bx].off, re(c.coef), im(c.coef)});
        plan->flops += (ScalarTraits<T>::is_complex ? 8.0 : 2.0) * (double)ob.numel;
      }
      I.ns = (int32_t)srcs.size() - I.s0;
      items.push_back(I);
    }
    upload_lincomb(ctx, *plan, items, srcs);
    cache.lincomb[key] = plan;
  }
  return run_lincomb<T>(ctx, *plan, X);
}

#define INST(T)                                                                                                                        \
  template BTensor<T> from_dense<T>(Ctx*, const DTensor<T>&, const std::vector<std::shared_ptr<BMode>>&, std::shared_ptr<BStruct>);     \
  template DTensor<T> to_dense<T>(Ctx*, const BTensor<T>&);                                                                            \
  template BTensor<T> conform<T>(Ctx*, BCache&, const BTensor<T>&, std::shared_ptr<BStruct>, const std::vector<Label>&);                          \
  template BTensor<T> bcontract<T>(Ctx*, BCache&, const BTensor<T>&, const BTensor<T>&, bool, bool, int);                                \
  template BTensor<T> bapply_small<T>(Ctx*, BCache&, const BTensor<T>&, const std::vector<T>&, const std::vector<Label>&,               \
                                      const std::vector<int64_t>&, const std::vector<Label>&, uint64_t);
INST(double)
INST(cdouble)

}  // namespace nsb
