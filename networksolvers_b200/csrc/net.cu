// Network state, projected-operator environments and the three region hooks (see net.h).
//
// Reference behaviour restated here (paths relative to the reference repo; UPSTREAM = ITensorNetworks /
// ITensors / KrylovKit rules):
//   Net::extract          src/extracter.jl:3-17
//   Net::orthogonalize    UPSTREAM itn.orthogonalize (App. A.3)
//   Net::position/make_env UPSTREAM itn.position / make_environment (App. A.2)
//   Net::apply_heff       src/operator_map.jl:15-42 (fixed order: one environment on the first site, site
//                         operators as early as possible, remaining environments)
//   Net::update_eigsolve  src/eigsolve.jl:14-28 + KrylovKit.eigsolve Lanczos (App. A.6)
//   Net::update_exp       src/applyexp.jl:18-48 + src/local_solvers/runge_kutta.jl:2-25 / KrylovKit.exponentiate (A.7)
//   Net::insert           src/inserter.jl:3-33 + ITensors.factorize (App. A.4, A.5)
//   Net::expand_densitymatrix  src/subspace/densitymatrix.jl:5-74, src/subspace/subspace.jl:28-48
#include "net.h"

#include "nccl_dyn.h"
#include "rangefinder.h"

#include <algorithm>
#include <cmath>
#include <functional>

namespace nsb {


// ------------------------------------------------------------------------------------------------
// small host dense helpers
// ------------------------------------------------------------------------------------------------
void host_sym_eig(int n, std::vector<double>& A, std::vector<double>& evals, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[i + (size_t)i * n] = 1.0;
  auto a = [&](int i, int j) -> double& { return A[i + (size_t)j * n]; };
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) { diag += a(i, i) * a(i, i); for (int j = i + 1; j < n; ++j) off += a(i, j) * a(i, j); }
    if (off <= 1e-60 || off <= 1e-34 * diag) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double apq = a(p, q);
        if (apq == 0.0) continue;
        double zeta = (a(q, q) - a(p, p)) / (2.0 * apq);
        double t = std::copysign(1.0, zeta) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < n; ++k) { double akp = a(k, p), akq = a(k, q); a(k, p) = c * akp - s * akq; a(k, q) = s * akp + c * akq; }
        for (int k = 0; k < n; ++k) { double apk = a(p, k), aqk = a(q, k); a(p, k) = c * apk - s * aqk; a(q, k) = s * apk + c * aqk; }
        for (int k = 0; k < n; ++k) { double vkp = V[k + (size_t)p * n], vkq = V[k + (size_t)q * n]; V[k + (size_t)p * n] = c * vkp - s * vkq; V[k + (size_t)q * n] = s * vkp + c * vkq; }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return a(x, x) < a(y, y); });
  evals.resize(n);
  std::vector<double> Vs((size_t)n * n);
  for (int j = 0; j < n; ++j) { evals[j] = a(order[j], order[j]); for (int k = 0; k < n; ++k) Vs[k + (size_t)j * n] = V[k + (size_t)order[j] * n]; }
  V.swap(Vs);
}

void host_expm_complex(int n, std::vector<std::complex<double>>& A) {
  typedef std::complex<double> C;
  double nrm = 0.0;
  for (int j = 0; j < n; ++j) { double s = 0.0; for (int i = 0; i < n; ++i) s += std::abs(A[i + (size_t)j * n]); nrm = std::max(nrm, s); }
  int sq = 0;
  while (nrm > 0.25) { nrm *= 0.5; ++sq; }
  double scale = std::ldexp(1.0, -sq);
  std::vector<C> X((size_t)n * n), term((size_t)n * n, C(0)), E((size_t)n * n, C(0)), tmp((size_t)n * n);
  for (size_t i = 0; i < X.size(); ++i) X[i] = A[i] * scale;
  for (int i = 0; i < n; ++i) { term[i + (size_t)i * n] = 1.0; E[i + (size_t)i * n] = 1.0; }
  auto matmul = [&](const std::vector<C>& P, const std::vector<C>& Q, std::vector<C>& R) {
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        C s(0);
        for (int k = 0; k < n; ++k) s += P[i + (size_t)k * n] * Q[k + (size_t)j * n];
        R[i + (size_t)j * n] = s;
      }
  };
  for (int k = 1; k <= 24; ++k) {
    matmul(term, X, tmp);
    for (size_t i = 0; i < tmp.size(); ++i) { term[i] = tmp[i] / (double)k; E[i] += term[i]; }
  }
  for (int s = 0; s < sq; ++s) { matmul(E, E, tmp); E.swap(tmp); }
  A.swap(E);
}

// ------------------------------------------------------------------------------------------------
// construction, labels, I/O
// ------------------------------------------------------------------------------------------------
template <typename T>
Net<T>::Net(Ctx* c, int nv, const int32_t* e, int ne, const int64_t* sd) {
  ctx = c;
  dtype = ScalarTraits<T>::dtype;
  NSB_REQUIRE(nv >= 1 && ne == nv - 1, NSB_EINVAL, "network must be a tree (nedges == nverts - 1)");
  nverts = nv;
  adj.resize(nv);
  site_dims.assign(sd, sd + nv);
  for (int i = 0; i < ne; ++i) {
    int u = e[2 * i], v = e[2 * i + 1];
    NSB_REQUIRE(u >= 0 && u < nv && v >= 0 && v < nv && u != v, NSB_EINVAL, "bad edge");
    NSB_REQUIRE(!eid.count({u, v}), NSB_EINVAL, "duplicate edge");
    edges.push_back({u, v});
    eid[{u, v}] = i;
    eid[{v, u}] = i;
    adj[u].push_back(v);
    adj[v].push_back(u);
  }
  std::vector<int> seen;
  if (nv > 1) { subtree(0, -1, seen); NSB_REQUIRE((int)seen.size() == nv, NSB_EINVAL, "graph is not connected"); }
  psi.resize(nv);
  W.resize(nv);
  ver.assign(nv, 0);
  for (int v = 0; v < nv; ++v) ortho.push_back(v);
}

template <typename T>
std::vector<Label> Net<T>::canonical_labels(int v) const {
  std::vector<Label> out;
  if (!adj[v].empty()) out.push_back(llink(v, adj[v][0]));
  out.push_back(lsite(v));
  for (size_t i = 1; i < adj[v].size(); ++i) out.push_back(llink(v, adj[v][i]));
  return out;
}

template <typename T>
std::vector<Label> Net<T>::decode_legs(int rank, const int32_t* legs, bool is_operator) const {
  std::vector<Label> out;
  for (int i = 0; i < rank; ++i) {
    int a = legs[2 * i], b = legs[2 * i + 1];
    NSB_REQUIRE(a >= 0 && a < nverts, NSB_EINVAL, "leg: bad vertex");
    if (b == NSB_SITE) out.push_back(lsite(a, 0));
    else if (b == NSB_SITE_OUT) out.push_back(lsite(a, 1));
    else {
      NSB_REQUIRE(b >= 0 && b < nverts && eid.count({a, b}), NSB_EINVAL, "leg: not an edge of the tree");
      out.push_back(is_operator ? lop(a, b) : llink(a, b));
    }
  }
  return out;
}

template <typename T>
void Net<T>::encode_legs(const std::vector<Label>& labels, int32_t* legs) const {
  for (size_t i = 0; i < labels.size(); ++i) {
    Label l = labels[i];
    int kind = label_kind(l), id = label_id(l);
    if (kind == LK_SITE) { legs[2 * i] = id; legs[2 * i + 1] = label_plev(l) == 0 ? NSB_SITE : NSB_SITE_OUT; }
    else if (kind == LK_LINK || kind == LK_OP) { legs[2 * i] = edges[id].first; legs[2 * i + 1] = edges[id].second; }
    else { legs[2 * i] = -100 - id; legs[2 * i + 1] = -3; }
  }
}

template <typename T>
void Net<T>::canonicalize(int v) {
  std::vector<Label> can = canonical_labels(v);
  if (psi[v].labels != can) psi[v] = permuted(ctx, psi[v], can);
}

template <typename T>
void Net<T>::site_upload(int v, int rank, const int32_t* legs, const int64_t* dims, const void* host) {
  NSB_REQUIRE(v >= 0 && v < nverts, NSB_EINVAL, "site_upload: bad vertex");
  std::vector<Label> labels = decode_legs(rank, legs, false);
  std::vector<Label> can = canonical_labels(v);
  {
    std::vector<Label> a = labels, b = can;
    std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
    NSB_REQUIRE(a == b, NSB_EINVAL, "site_upload: legs must be the site index and one link per neighbour");
  }
  std::vector<int64_t> d(dims, dims + rank);
  for (int i = 0; i < rank; ++i) {
    NSB_REQUIRE(d[i] >= 1, NSB_EINVAL, "site_upload: bad dimension");
    if (label_kind(labels[i]) == LK_SITE) NSB_REQUIRE(d[i] == site_dims[v], NSB_EINVAL, "site_upload: site dimension mismatch");
  }
  DTensor<T> t(ctx, d, labels);
  NSB_CUDA(cudaMemcpyAsync(t.data(), host, sizeof(T) * t.numel(), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
  uint64_t pb = ctx->cnt.permute_bytes;
  psi[v] = t;
  canonicalize(v);
  ctx->cnt.permute_bytes = pb;
  ver[v]++;
}

template <typename T>
void Net<T>::site_fill_random(int v, int rank, const int32_t* legs, const int64_t* dims, uint64_t seed, double scale) {
  NSB_REQUIRE(v >= 0 && v < nverts, NSB_EINVAL, "site_fill_random: bad vertex");
  std::vector<Label> labels = decode_legs(rank, legs, false);
  std::vector<int64_t> d(dims, dims + rank);
  DTensor<T> t(ctx, d, labels);
  fill_normal<T>(ctx, t.data(), t.numel(), seed, scale);
  uint64_t pb = ctx->cnt.permute_bytes;
  psi[v] = t;
  canonicalize(v);
  ctx->cnt.permute_bytes = pb;
  ver[v]++;
}

template <typename T>
void Net<T>::site_info(int v, int32_t* rank, int32_t* legs, int64_t* dims) {
  NSB_REQUIRE(v >= 0 && v < nverts && psi[v].valid(), NSB_EINVAL, "site_info: no tensor");
  *rank = psi[v].rank();
  if (legs) encode_legs(psi[v].labels, legs);
  if (legs)  // orient link legs as (v, neighbour)
    for (int i = 0; i < psi[v].rank(); ++i)
      if (legs[2 * i + 1] >= 0 && legs[2 * i] != v) std::swap(legs[2 * i], legs[2 * i + 1]);
  if (dims) for (int i = 0; i < psi[v].rank(); ++i) dims[i] = psi[v].dims[i];
}

template <typename T>
void Net<T>::site_download(int v, void* host) {
  NSB_REQUIRE(v >= 0 && v < nverts && psi[v].valid(), NSB_EINVAL, "site_download: no tensor");
  NSB_CUDA(cudaMemcpyAsync(host, psi[v].data(), sizeof(T) * psi[v].numel(), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
}

template <typename T>
void Net<T>::mpo_upload(int v, int rank, const int32_t* legs, const int64_t* dims, const void* host) {
  NSB_REQUIRE(v >= 0 && v < nverts, NSB_EINVAL, "mpo_upload: bad vertex");
  std::vector<Label> labels = decode_legs(rank, legs, true);
  NSB_REQUIRE(rank == (int)adj[v].size() + 2, NSB_EINVAL, "mpo_upload: need one operator link per neighbour plus site in/out");
  std::vector<int64_t> d(dims, dims + rank);
  DTensor<T> t(ctx, d, labels);
  NSB_REQUIRE(t.find(lsite(v, 0)) >= 0 && t.find(lsite(v, 1)) >= 0, NSB_EINVAL, "mpo_upload: site in/out legs missing");
  NSB_CUDA(cudaMemcpyAsync(t.data(), host, sizeof(T) * t.numel(), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
  W[v] = t;
  if ((int)Whost.size() > v) Whost[v].clear();
  envs.clear();
  plan.clear();
}

template <typename T>
void Net<T>::set_ortho_region(const int32_t* verts, int n) {
  ortho.assign(verts, verts + n);
}
template <typename T>
void Net<T>::get_ortho_region(int32_t* verts, int32_t* n) {
  *n = (int)ortho.size();
  if (verts) for (size_t i = 0; i < ortho.size(); ++i) verts[i] = ortho[i];
}
template <typename T>
int64_t Net<T>::linkdim(int u, int v) {
  NSB_REQUIRE(eid.count({u, v}) && psi[u].valid(), NSB_EINVAL, "linkdim: bad edge");
  return psi[u].dim_of(llink(u, v));
}
template <typename T>
int64_t Net<T>::maxlinkdim() {
  int64_t m = 1;
  for (auto& e : edges) m = std::max(m, linkdim(e.first, e.second));
  return m;
}

// ------------------------------------------------------------------------------------------------
// graph helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
void Net<T>::subtree(int u, int v, std::vector<int>& out) const {
  std::vector<int> stack{u};
  std::vector<char> seen(nverts, 0);
  seen[u] = 1;
  if (v >= 0) seen[v] = 1;
  while (!stack.empty()) {
    int x = stack.back(); stack.pop_back();
    out.push_back(x);
    for (int n : adj[x]) if (!seen[n]) { seen[n] = 1; stack.push_back(n); }
  }
}

template <typename T>
std::vector<int> Net<T>::path(int a, int b) const {
  std::vector<int> parent(nverts, -2), stack{a};
  parent[a] = -1;
  while (!stack.empty()) {
    int x = stack.back(); stack.pop_back();
    for (int n : adj[x]) if (parent[n] == -2) { parent[n] = x; stack.push_back(n); }
  }
  std::vector<int> p{b};
  while (p.back() != a) p.push_back(parent[p.back()]);
  std::reverse(p.begin(), p.end());
  return p;
}

// ------------------------------------------------------------------------------------------------
// gauge moves
// ------------------------------------------------------------------------------------------------
template <typename T>
void Net<T>::qr_step(int a, int b) {
  Label l = llink(a, b);
  DTensor<T> A = psi[a];
  std::vector<Label> order;
  for (Label x : A.labels) if (x != l) order.push_back(x);
  order.push_back(l);
  DTensor<T> Ap = permuted(ctx, A, order);
  if (Ap.data() == A.data()) Ap = clone(ctx, A);   // qr_thin destroys its input
  int64_t cols = Ap.dims.back(), rows = Ap.numel() / cols, k = std::min(rows, cols);
  std::vector<int64_t> qd(Ap.dims.begin(), Ap.dims.end() - 1);
  qd.push_back(k);
  Label aux = make_label(LK_AUX, 1);
  DTensor<T> Q, R;
  if (qn_on) {
    std::vector<Label> others(order.begin(), order.end() - 1);
    std::vector<int64_t> odims(Ap.dims.begin(), Ap.dims.end() - 1);
    std::vector<int64_t> rk = multi_keys(a, others, odims, false);                 // charge of a's side of {a, b}
    std::vector<int64_t> ck = multi_keys(b, {l}, {cols}, false);                   // the same, as labelled on the link
    DevBuf Qb, Rb;
    std::vector<int64_t> newk;
    qr_qn(Ap.data(), rows, cols, rk, ck, Qb, Rb, &k, newk);
    qd.back() = k;
    Q.buf = std::make_shared<DevBuf>(std::move(Qb)); Q.dims = qd; Q.labels = order;
    R.buf = std::make_shared<DevBuf>(std::move(Rb)); R.dims = {k, cols}; R.labels = {aux, l};
    qn_store_link(a, b, newk);
  } else {
    Q = DTensor<T>(ctx, qd, order);
    R = DTensor<T>(ctx, {k, cols}, {aux, l});
    qr_thin<T>(ctx, Ap.data(), rows, cols, rows, Q.data(), rows, R.data(), k);
  }
  psi[a] = Q;
  canonicalize(a);
  ver[a]++;
  DTensor<T> nb = contract(ctx, psi[b], R, false, false, 1);   // psi[b] with l replaced by aux
  std::vector<Label> nl = nb.labels;
  for (auto& x : nl) if (x == aux) x = l;
  psi[b] = nb.relabeled(nl);
  canonicalize(b);
  ver[b]++;
}

template <typename T>
int Net<T>::orthogonalize(const std::vector<int>& target) {
  {
    std::set<int> s1(target.begin(), target.end()), s2(ortho.begin(), ortho.end());
    if (s1 == s2) return 0;
  }
  // BFS from target[0]
  int root = target[0];
  std::vector<int> parent(nverts, -2), dist(nverts, 0), queue{root};
  parent[root] = -1;
  for (size_t qi = 0; qi < queue.size(); ++qi) {
    int x = queue[qi];
    for (int n : adj[x]) if (parent[n] == -2) { parent[n] = x; dist[n] = dist[x] + 1; queue.push_back(n); }
  }
  std::vector<char> mark(nverts, 0);
  auto mark_path = [&](int t) { while (t != -1 && !mark[t]) { mark[t] = 1; t = parent[t]; } };
  for (int t : target) mark_path(t);
  for (int t : ortho) mark_path(t);
  std::vector<int> order;
  for (int v = 0; v < nverts; ++v) if (mark[v] && v != root) order.push_back(v);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return dist[x] > dist[y]; });
  std::set<int> tset(target.begin(), target.end());
  int steps = 0;
  for (int v : order) {
    int p = parent[v];
    if (tset.count(v) && tset.count(p)) continue;
    qr_step(v, p);
    ++steps;
  }
  ortho = target;
  return steps;
}

// ------------------------------------------------------------------------------------------------
// local tensor, environments, H_eff
// ------------------------------------------------------------------------------------------------
template <typename T>
DTensor<T> Net<T>::build_theta(const std::vector<int>& reg) {
  if (reg.size() == 1) return clone(ctx, psi[reg[0]]);
  return contract(ctx, psi[reg[0]], psi[reg[1]], false, false, 1);
}

template <typename T>
std::vector<Label> Net<T>::w_out_labels(const DTensor<T>& X, const DTensor<T>& Wv, int v, const std::vector<int>& reg) const {
  std::vector<Label> out;
  auto is_new = [&](Label l) { return X.find(l) < 0; };
  auto placed = [&](Label l) { return std::find(out.begin(), out.end(), l) != out.end(); };
  for (Label lab : X.labels) {
    if (Wv.find(lab) >= 0) {  // contracted
      if (lab == lsite(v, 0)) {
        out.push_back(lsite(v, 1));
        for (Label wl : Wv.labels) {   // operator link to the next site of the region
          if (label_kind(wl) != LK_OP || !is_new(wl)) continue;
          auto e = edges[label_id(wl)];
          int other = (e.first == v) ? e.second : e.first;
          if (std::find(reg.begin(), reg.end(), other) != reg.end() && !placed(wl)) out.push_back(wl);
        }
      }
      continue;
    }
    out.push_back(lab);
    if (label_kind(lab) == LK_LINK && (label_plev(lab) == 0 || label_plev(lab) == 2)) {
      Label ol = make_label(LK_OP, label_id(lab), 0);
      if (Wv.find(ol) >= 0 && is_new(ol) && !placed(ol)) out.push_back(ol);
    }
  }
  for (Label wl : Wv.labels) if (is_new(wl) && !placed(wl)) out.push_back(wl);
  return out;
}

// Output label order of the merged two-site operator = the order the two single-site applications would produce.
template <typename T>
std::vector<Label> Net<T>::merged_out_labels(const DTensor<T>& X, int a, int b) const {
  DTensor<T> fake;                       // labels / dims only
  fake.labels = w_out_labels(X, W[a], a, pos);
  fake.dims.resize(fake.labels.size(), 1);
  fake.buf = X.buf;
  return w_out_labels(fake, W[b], b, pos);
}

template <typename T>
int Net<T>::make_env(int u, int v) {
  auto key = std::make_pair(u, v);
  if (envs.count(key)) return 0;
  int built = 0;
  std::vector<int> others;
  for (int n : adj[u]) if (n != v) others.push_back(n);
  for (int n : others) built += make_env(n, u);
  NSB_REQUIRE(psi[u].valid() && W[u].valid(), NSB_EINVAL, "make_env: state or operator tensor missing");
  NSB_REQUIRE(!fit_mode || xket[u].valid(), NSB_EINVAL, "make_env: fitting target tensor missing");
  for (int n : others) env_promote(envs.at({n, u}));     // (multi-GPU: inputs that live as slabs)
  if (qn_bs()) {     // QN network: the environment is built and kept as symmetry blocks (grouped sector GEMMs)
    Env benv;
    if (make_env_bt(u, v, others, &benv)) {
      benv.deps.push_back({u, ver[u]});
      for (int n : others) for (auto& d : envs.at({n, u}).deps) benv.deps.push_back(d);
      envs[key] = benv;
      ctx->cnt.env_builds++;
      return built + 1;
    }
    for (int n : others) env_dense(n, u);      // not permutation-free on blocks: dense engine on dense copies
  }
  DTensor<T> X = fit_mode ? xket[u] : psi[u];
  size_t start = 0;
  DTensor<T> bra = psi[u].primed();
  // multi-GPU: split the contraction over the bra index of the (single) incoming environment across the ranks -- 1/G of
  // every GEMM of the update -- and sum the partial environments with one all-reduce
  bool split = false;
  if (shard_enabled && ctx->nranks > 1 && ctx->nccl_comm && !fit_mode && !qn_on && others.size() == 1) {
    const DTensor<T>& Ein = envs.at({others[0], u}).t;
    const int G = ctx->nranks;
    if (Ein.rank() == 3 && Ein.dims[2] >= 4 * G && Ein.dims[2] % G == 0) {
      const int64_t per = Ein.dims[2] / G, lo = per * ctx->rank, hi = lo + per;
      DTensor<T> bs;
      if (slab_of(bra, Ein.labels[2], lo, hi, &bs)) {
        X = contract(ctx, X, Ein.last_mode_slab(lo, hi), false, false, 1);
        bra = bs;
        start = 1;
        split = true;
      }
    }
  }
  if (!split && !others.empty()) { X = contract(ctx, X, envs.at({others[0], u}).t, false, false, 1); start = 1; }
  {
    SmallOp<T> op;
    std::vector<int> reg{u};
    X = apply_small(ctx, op, X, W[u], w_out_labels(X, W[u], u, reg));
  }
  for (size_t i = start; i < others.size(); ++i) X = contract(ctx, X, envs.at({others[i], u}).t, false, false, 1);
  std::vector<Label> want{fit_mode ? lxlink(u, v) : llink(u, v, 0), lop(u, v), llink(u, v, 1)};
  std::vector<Label> l1, l2;
  bool d1 = contract_direct_labels(bra, X, &l1), d2 = contract_direct_labels(X, bra, &l2);
  DTensor<T> E;
  if (d1 && l1 == want) E = contract(ctx, bra, X, true, false, 1);
  else if (d2 && l2 == want) E = contract(ctx, X, bra, false, true, 1);
  else E = contract(ctx, X, bra, false, true, 0);
  if (E.labels != want) E = permuted(ctx, E, want);
  if (split) comm_allreduce(E.data(), E.numel());
  Env env;
  env.t = E;
  if (ctx->opt.skip_identity && !fit_mode && E.rank() == 3 && E.dims[0] == E.dims[2] && E.dims[1] >= 2 && E.dims[1] <= 64) {
    std::vector<double> dev(E.dims[1]);
    identity_deviation<T>(ctx, E.data(), E.dims[0], E.dims[1], dev.data());
    for (int64_t w = 0; w < E.dims[1]; ++w) if (dev[w] <= 1e-10) { env.ident = (int)w; break; }
  }
  if (qn_bs()) env.bt = bt_of(E);     // (dense fallback above) keep the block form too: later block contractions use it
  env.deps.push_back({u, ver[u]});
  for (int n : others) for (auto& d : envs.at({n, u}).deps) env.deps.push_back(d);
  envs[key] = env;
  ctx->cnt.env_builds++;
  if (env_sharding_on())      // the inputs are done with: keep only their slabs unless the current region touches them
    for (int n : others) if (!env_is_hot(n, u)) env_demote(envs.at({n, u}));
  return built + 1;
}

// ---- multi-GPU: environments sharded in HBM ---------------------------------------
template <typename T>
bool Net<T>::env_sharding_on() const {
  return shard_enabled && ctx->opt.shard_envs && ctx->nranks > 1 && ctx->nccl_comm && !fit_mode && !qn_on;
}
template <typename T>
bool Net<T>::env_is_hot(int u, int v) const {
  (void)u;
  return std::find(hot_region.begin(), hot_region.end(), v) != hot_region.end() ||
         std::find(pos.begin(), pos.end(), v) != pos.end();
}
template <typename T>
void Net<T>::env_promote(Env& e) {
  if (e.t.valid() || !e.shard.valid()) return;
  DTensor<T> full(ctx, e.t.dims, e.t.labels);
  comm_allgather(e.shard.data(), full.data(), e.shard.numel());
  e.t = full;
  e.shard = DTensor<T>();      // (cut again on the next demotion: a device copy of 1 / G of the tensor)
}
template <typename T>
void Net<T>::env_demote(Env& e) {
  if (!e.t.valid() || e.t.rank() != 3) return;
  const int G = ctx->nranks;
  const int64_t nb = e.t.dims[2];
  if (nb % G != 0 || nb < 4 * G) return;
  if (!e.shard.valid()) {
    const int64_t per = nb / G;
    DTensor<T> src = e.t.last_mode_slab(per * ctx->rank, per * (ctx->rank + 1));
    e.shard = DTensor<T>(ctx, src.dims, src.labels);
    vec_copy<T>(ctx, src.numel(), src.data(), e.shard.data());
  }
  DTensor<T> meta;             // dims / labels stay readable, the buffer goes back to the pool
  meta.dims = e.t.dims;
  meta.labels = e.t.labels;
  e.t = meta;
}
template <typename T>
void Net<T>::env_bytes(int64_t* resident, int64_t* replicated) {
  int64_t r = 0, f = 0;
  for (auto& kv : envs) {
    const Env& e = kv.second;
    int64_t n = 1;
    for (int64_t d : e.t.dims) n *= d;
    if (e.t.dims.empty()) n = 0;
    f += n * (int64_t)sizeof(T);
    if (e.t.valid()) r += n * (int64_t)sizeof(T);
    if (e.shard.valid()) r += e.shard.numel() * (int64_t)sizeof(T);
  }
  if (resident) *resident = r;
  if (replicated) *replicated = f;
}
template <typename T>
void Net<T>::env_rebalance() {
  if (!env_sharding_on()) return;
  for (auto& kv : envs) {
    if (env_is_hot(kv.first.first, kv.first.second)) env_promote(kv.second);
    else env_demote(kv.second);
  }
}

template <typename T>
void Net<T>::build_plan() {
  plan.clear();
  for (size_t i = 0; i < pos.size(); ++i) {
    int v = pos[i];
    std::vector<int> ext;
    for (int n : adj[v]) if (std::find(pos.begin(), pos.end(), n) == pos.end()) ext.push_back(n);
    size_t start = 0;
    if (i == 0 && !ext.empty()) {
      Step s; s.type = 0; s.u = ext[0]; s.v = v;
      plan.push_back(std::move(s));
      start = 1;
    }
    { Step s; s.type = 1; s.u = v; s.v = v; plan.push_back(std::move(s)); }
    for (size_t j = start; j < ext.size(); ++j) {
      Step s; s.type = 0; s.u = ext[j]; s.v = v;
      plan.push_back(std::move(s));
    }
  }
  // merge two consecutive site-operator steps (chain-like 2-site regions) into one pass over the big intermediate
  for (size_t i = 0; ctx->opt.merge_site_ops && i + 1 < plan.size(); ++i) {
    if (plan[i].type == 1 && plan[i + 1].type == 1) {
      int a = plan[i].v, b = plan[i + 1].v;
      const DTensor<T>&Wa = W[a], &Wb = W[b];
      int64_t kk = 1, nn = 1;   // contracted / new extents of the merged operator
      kk = site_dims[a] * site_dims[b];
      nn = kk;
      for (int j = 0; j < Wa.rank(); ++j) if (label_kind(Wa.labels[j]) == LK_OP && Wb.find(Wa.labels[j]) < 0) { kk *= Wa.dims[j]; nn *= Wa.dims[j]; }
      for (int j = 0; j < Wb.rank(); ++j) if (label_kind(Wb.labels[j]) == LK_OP && Wa.find(Wb.labels[j]) < 0) { kk *= Wb.dims[j]; nn *= Wb.dims[j]; }
      if (kk * nn > 64 * 64 * 16) continue;   // (upper bound on K*N; keeps the operator in shared memory)
      Step m; m.type = 2; m.u = a; m.v = b;
      uint64_t pb = ctx->cnt.permute_bytes;
      m.Wm = contract(ctx, Wa, Wb, false, false, 0);
      ctx->cnt.permute_bytes = pb;
      plan[i] = std::move(m);
      plan.erase(plan.begin() + i + 1);
    }
  }
}

template <typename T>
int Net<T>::position(const std::vector<int>& reg) {
  // drop environments built from tensors that have changed since
  for (auto it = envs.begin(); it != envs.end();) {
    bool ok = true;
    for (auto& d : it->second.deps) if (ver[d.first] != d.second) { ok = false; break; }
    if (ok) ++it; else it = envs.erase(it);
  }
  pos = reg;
  pos_on_edge = false;
  hot_region = reg;
  int built = 0;
  for (int v : reg)
    for (int n : adj[v])
      if (std::find(reg.begin(), reg.end(), n) == reg.end()) built += make_env(n, v);
  env_rebalance();
  build_plan();
  prepare_identity_skip();
  return built;
}

// First environment of the plan: when it has an identity channel, keep a copy without that channel ([ket, W - 1, bra]); apply_heff contracts theta with it and splices theta itself in as the
// missing channel.  (The last environment needs no copy: its channel is a contiguous block of the contraction index.)
template <typename T>
void Net<T>::prepare_identity_skip() {
  skipped_last_apply = -1.0;
  first_ident = -1;
  first_compact = DTensor<T>();
  if (!ctx->opt.skip_identity || fit_mode || plan.size() < 2 || plan.front().type != 0) return;
  const Env& e = envs.at({plan.front().u, plan.front().v});
  const DTensor<T>& E = e.t;
  if (e.ident < 0 || E.rank() != 3) return;
  const int64_t n = E.dims[0], Wd = E.dims[1], n2 = E.dims[2];
  if (n != n2 || Wd < 2) return;
  first_compact = DTensor<T>(ctx, {n, Wd - 1, n2}, E.labels);
  const int64_t w = e.ident;   // channels [0, w) and (w, Wd) keep their order
  if (w > 0) copy_block<T>(ctx, E.data(), n * Wd, first_compact.data(), n * (Wd - 1), n * w, n2);
  if (w < Wd - 1) copy_block<T>(ctx, E.data() + n * (w + 1), n * Wd, first_compact.data() + n * w, n * (Wd - 1), n * (Wd - 1 - w), n2);
  first_ident = e.ident;
}

template <typename T>
double Net<T>::skipped_flops(const DTensor<T>& x) const {
  // dry run of the conditions of apply_heff on the labels of x (the site-operator steps keep the big modes in place, so
  // the tensor that meets the last environment has x's link at the same end with the operator link next to it)
  if (!ctx->opt.skip_identity || (shard_active && (!ctx->opt.skip_identity_sharded || ctx->shard_fused)) || fit_mode || plan.size() < 2 || x.rank() < 2) return 0.0;
  double f = 0.0;
  const double cplx = ScalarTraits<T>::is_complex ? 4.0 : 1.0;
  if (first_ident >= 0 && first_compact.valid() && plan.front().type == 0) {
    const DTensor<T>& E = envs.at({plan.front().u, plan.front().v}).t;
    if (E.dims[0] == E.dims[2] && (x.labels[0] == E.labels[0] || x.labels.back() == E.labels[0])) f += 2.0 * (double)x.numel() * (double)E.dims[2];
  }
  if (plan.back().type == 0) {
    const Env& e = envs.at({plan.back().u, plan.back().v});
    if (e.ident >= 0 && e.t.rank() == 3 && e.t.dims[0] == e.t.dims[2] && (x.labels.back() == e.t.labels[0] || x.labels[0] == e.t.labels[0]))
      f += 2.0 * (double)x.numel() * (double)e.t.dims[2];
  }
  return f * cplx;
}

template <typename T> struct NcclType;
template <> struct NcclType<double> { static constexpr int mult = 1; };
template <> struct NcclType<cdouble> { static constexpr int mult = 2; };

// ------------------------------------------------------------------------------------------------
// Multi-GPU partition of the region step.  One process per GPU, every rank runs the same sweep in lock step
// on replicated site tensors and environments; what is split is the arithmetic:
//
//   H_eff application: the Krylov vectors are sharded along the LAST bond of theta (contiguous slabs, equal on every rank).
//     RS ("reduce-scatter") positions -- the last environment of the plan contracts that bond (sweeping right): every step
//        runs on the slab, the last contraction uses rows [lo, hi) of that environment and yields a partial full-size result,
//        one ncclReduceScatter leaves theta' sharded like theta.
//     AG ("all-gather") positions -- the first environment of the plan contracts that bond (sweeping left): one ncclAllGather
//        completes the input, the first contraction is split along the environment's bra index (a column slab of the
//        environment: a view), all later steps act on other modes, and the result is the slab of theta' -- no reduction.
//     Dots and norms of the Lanczos / Runge-Kutta vectors are local partial sums + an all-reduce of two doubles.
//     Bonds that do not divide evenly fall back to the round-1 form (AR: full vectors, all-reduce of theta').
//   Environment update: the contraction over the incoming environment's bra index is split (1/G of every GEMM), partial
//     results summed by one all-reduce.
//   Factorisation (Net::insert -> factorize_left, linalg.cu): Gram matrix by output column slabs + all-gather, back-
//     transformation of the kept eigenvectors by column slabs + all-gather, C = U^H theta by column slabs + all-gather; the
//     tridiagonalisation and the divide & conquer run replicated on identical data.
// ------------------------------------------------------------------------------------------------
template <typename T>
void Net<T>::nccl_check(int r, const char* what) {
  if (r != (int)ncclSuccess) throw Error(NSB_ENCCL, std::string(what) + ": " + nccl_api().GetErrorString((ncclResult_t)r));
}
// ctx option "nccl_sync": host-synchronise the stream before and after every collective (diagnostic / belt and braces)
#define NSB_COMM_SYNC() do { if (ctx->opt.nccl_sync) ctx->sync(); } while (0)

template <typename T>
void Net<T>::comm_allreduce(T* buf, int64_t n) {
  NSB_COMM_SYNC();
  nccl_check(nccl_api().AllReduce(buf, buf, (size_t)n * NcclType<T>::mult, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream), "ncclAllReduce");
  ctx->cnt.kernel_launches++;
  NSB_COMM_SYNC();
}
template <typename T>
void Net<T>::comm_allgather(const T* send, T* recv, int64_t n_per_rank) {
  NSB_COMM_SYNC();
  nccl_check(nccl_api().AllGather(send, recv, (size_t)n_per_rank * NcclType<T>::mult, ncclDouble, (ncclComm_t)ctx->nccl_comm, ctx->stream), "ncclAllGather");
  ctx->cnt.kernel_launches++;
  NSB_COMM_SYNC();
}
template <typename T>
void Net<T>::comm_reduce_scatter(const T* send, T* recv, int64_t n_per_rank) {
  NSB_COMM_SYNC();
  nccl_check(nccl_api().ReduceScatter(send, recv, (size_t)n_per_rank * NcclType<T>::mult, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream),
             "ncclReduceScatter");
  ctx->cnt.kernel_launches++;
  NSB_COMM_SYNC();
}

// slab [lo, hi) of mode `l` of t: a view when l is the last mode, a strided copy when it is the first one
template <typename T>
bool Net<T>::slab_of(const DTensor<T>& t, Label l, int64_t lo, int64_t hi, DTensor<T>* out) {
  const int pos = t.find(l);
  if (pos < 0) return false;
  if (pos == t.rank() - 1) { *out = t.last_mode_slab(lo, hi); return true; }
  if (pos == 0) {
    std::vector<int64_t> d = t.dims;
    d[0] = hi - lo;
    DTensor<T> o(ctx, d, t.labels);
    if (hi > lo) copy_block<T>(ctx, t.data() + lo, t.dims[0], o.data(), hi - lo, hi - lo, t.numel() / t.dims[0]);
    *out = o;
    return true;
  }
  return false;
}

template <typename T>
void Net<T>::shard_prepare() {
  shard_active = false;
  shard_mode = 0;
  theta_is_slab = false;
  theta_slab = DTensor<T>();
  if (!shard_enabled || ctx->nranks <= 1 || !ctx->nccl_comm || plan.size() < 2 || !theta.valid() || fit_mode || qn_on) return;
  const int G = ctx->nranks;
  const int64_t nb = theta.dims.back();
  const Label lb = theta.labels.back();
  if (label_kind(lb) != LK_LINK) return;
  const Step& last = plan.back();
  const Step& first = plan.front();
  auto env_of = [&](const Step& st) -> const DTensor<T>& { return envs.at({st.u, st.v}).t; };
  int uses = 0;
  for (auto& st : plan) if (st.type == 0 && env_of(st).find(lb) >= 0) ++uses;
  if (uses != 1) return;
  const bool even = (nb % G == 0) && nb >= G && !ctx->shard_fused;
  if (last.type == 0 && env_of(last).rank() == 3 && env_of(last).labels[0] == lb) {
    const DTensor<T>& E = env_of(last);
    const int64_t per = even ? nb / G : (nb + G - 1) / G;
    shard_lo = std::min<int64_t>(nb, per * ctx->rank);
    shard_hi = std::min<int64_t>(nb, shard_lo + per);
    const int64_t rows = shard_hi - shard_lo, cols = E.numel() / nb;
    std::vector<int64_t> d = E.dims;
    d[0] = rows;
    shard_env = DTensor<T>(ctx, d, E.labels);
    if (rows > 0) copy_block<T>(ctx, E.data() + shard_lo, nb, shard_env.data(), rows, rows, cols);
    shard_mode = even ? 1 : 3;
    shard_active = true;
    return;
  }
  if (even && first.type == 0 && env_of(first).rank() == 3 && env_of(first).labels[0] == lb && env_of(first).dims[2] == nb) {
    // the contraction must leave the bra link as the last mode (theta's last mode replaced by [operator link, bra link])
    std::vector<Label> pl;
    if (!contract_direct_labels(theta, env_of(first), &pl) || pl.empty() || pl.back() != env_of(first).labels[2]) return;
    const int64_t per = nb / G;
    shard_lo = per * ctx->rank;
    shard_hi = shard_lo + per;
    shard_env = DTensor<T>();
    shard_mode = 2;
    shard_active = true;
  }
}

template <typename T>
int Net<T>::set_shard(int enable) {
  shard_enabled = enable != 0;
  if (theta_is_slab) ensure_theta_full();
  if (!shard_enabled) { for (auto& kv : envs) env_promote(kv.second); }   // replicated again
  else env_rebalance();
  shard_prepare();
  return shard_active ? 1 : 0;
}

template <typename T>
void Net<T>::ensure_theta_full() {
  if (!theta_is_slab) return;
  DTensor<T> full(ctx, theta.dims, theta.labels);
  comm_allgather(theta_slab.data(), full.data(), theta_slab.numel());
  theta = full;
  theta_is_slab = false;
  theta_slab = DTensor<T>();
}

// T1 = X * L over L's ket link, where L is the first environment of the plan: the channel with L[:, w*, :] = 1 is X itself --
// contract with the W - 1 other channels (first_compact) and splice X in as the missing channel of the operator link.  Two
// layouts occur on the permutation-free chain path: the ket link is X's first mode (result [w, bra, rest of X]) or X's last
// mode (result [rest of X, w, bra]).  Returns false (X untouched) when the pattern does not apply.
// bra_lo / bra_hi (AG positions): only the bra-link slab [bra_lo, bra_hi) of the result is formed; Xs is that slab of X.
template <typename T>
bool Net<T>::skip_first_identity(DTensor<T>& X, int64_t bra_lo, int64_t bra_hi, const DTensor<T>* Xs) {
  if (!(first_ident >= 0 && first_compact.valid() && !fit_mode && plan.size() >= 2 && plan[0].type == 0)) return false;
  const DTensor<T>& E = envs.at({plan[0].u, plan[0].v}).t;
  const int64_t Wd = E.dims[1];
  const int r = X.rank();
  const bool slab = bra_hi > bra_lo;
  DTensor<T> comp = slab ? first_compact.last_mode_slab(bra_lo, bra_hi) : first_compact;
  std::vector<Label> pl;
  int64_t pre = 0;
  if (r >= 2 && E.dims[0] == E.dims[2] && contract_direct_labels(X, comp, &pl) && (int)pl.size() == r + 1) {
    if (!slab && X.labels[0] == E.labels[0] && pl[0] == E.labels[1] && pl[1] == E.labels[2]) pre = 1;
    else if (X.labels[r - 1] == E.labels[0] && pl[r - 1] == E.labels[1] && pl[r] == E.labels[2]) pre = X.numel() / X.dims[r - 1];
  }
  if (pre <= 0) return false;
  DTensor<T> Xc = contract(ctx, X, comp, false, false, 1);
  NSB_REQUIRE(Xc.labels == pl, NSB_EINTERNAL, "identity skipping: unexpected layout of the first contraction");
  const int wpos = (pre == 1) ? 0 : r - 1;
  std::vector<int64_t> fd = Xc.dims;
  fd[wpos] = Wd;
  DTensor<T> Xf(ctx, fd, Xc.labels);
  // pre = 1: all of X behind the operator link; else the extent of the bra link (its slab on AG positions)
  const int64_t post = slab ? (bra_hi - bra_lo) : X.numel() / pre;
  insert_mode<T>(ctx, Xc.data(), slab ? Xs->data() : X.data(), Xf.data(), pre, Wd - 1, first_ident, post);
  X = Xf;
  return true;
}

template <typename T>
DTensor<T> Net<T>::run_plan_steps(DTensor<T> X, size_t i0, size_t i1) {
  for (size_t i = i0; i < i1; ++i) {
    auto& s = plan[i];
    if (s.type == 0) X = contract(ctx, X, envs.at({s.u, s.v}).t, false, false, 1);
    else if (s.type == 1) X = apply_small(ctx, s.op, X, W[s.v], w_out_labels(X, W[s.v], s.v, pos));
    else X = apply_small(ctx, s.op, X, s.Wm, merged_out_labels(X, s.u, s.v));
  }
  return X;
}

// Contraction with the last environment of the plan E[(b, w), b'] over (ket link, operator link).  When the environment has an
// identity channel w*, that block of the contraction index contributes X itself: the result starts as that block of X and the
// other channels are added by at most two GEMMs (beta = 1).  Layouts: (b, w) are X's last two modes
// (out[p, b'] = X[p, K] R[K, b']) or its first two (out[b', q] = R[K, b']^T X[K, q]).
template <typename T>
DTensor<T> Net<T>::last_env_contract(const DTensor<T>& X, double* skipped) {
  const Env& e = envs.at({plan.back().u, plan.back().v});
  const DTensor<T>& E = e.t;
  const T one = from_complex<T>(1.0, 0.0);
  const int last_w = (ctx->opt.skip_identity && !fit_mode && e.ident >= 0 && E.rank() == 3 && E.dims[0] == E.dims[2]) ? e.ident : -1;
  if (last_w < 0) return contract(ctx, X, E, false, false, 1);
  const int r = X.rank();
  const int64_t nb = E.dims[0], Wd = E.dims[1], N = E.dims[2], Kc = nb * Wd;
  const int64_t k0 = nb * last_w, k1 = nb * (last_w + 1);    // identity block [k0, k1) of the contraction index
  if (r >= 3 && X.labels[r - 2] == E.labels[0] && X.labels[r - 1] == E.labels[1] && X.dims[r - 2] == nb && X.dims[r - 1] == Wd) {
    const int64_t P = X.numel() / Kc;
    std::vector<int64_t> od(X.dims.begin(), X.dims.end() - 2);
    std::vector<Label> ol(X.labels.begin(), X.labels.end() - 2);
    od.push_back(N); ol.push_back(E.labels[2]);
    DTensor<T> out(ctx, od, ol);
    vec_copy<T>(ctx, P * nb, X.data() + P * k0, out.data());
    if (k0 > 0) gemm<T>(ctx, OP_N, OP_N, P, N, k0, one, X.data(), P, 0, E.data(), Kc, 0, one, out.data(), P, 0, 1);
    if (k1 < Kc) gemm<T>(ctx, OP_N, OP_N, P, N, Kc - k1, one, X.data() + P * k1, P, 0, E.data() + k1, Kc, 0, one, out.data(), P, 0, 1);
    *skipped += 2.0 * (double)P * (double)nb * (double)N;
    return out;
  }
  if (r >= 3 && X.labels[0] == E.labels[0] && X.labels[1] == E.labels[1] && X.dims[0] == nb && X.dims[1] == Wd) {
    const int64_t Q = X.numel() / Kc;
    std::vector<int64_t> od{N};
    std::vector<Label> ol{E.labels[2]};
    od.insert(od.end(), X.dims.begin() + 2, X.dims.end());
    ol.insert(ol.end(), X.labels.begin() + 2, X.labels.end());
    DTensor<T> out(ctx, od, ol);
    copy_block<T>(ctx, X.data() + k0, Kc, out.data(), N, nb, Q);
    if (k0 > 0) gemm<T>(ctx, OP_T, OP_N, N, Q, k0, one, E.data(), Kc, 0, X.data(), Kc, 0, one, out.data(), N, 0, 1);
    if (k1 < Kc) gemm<T>(ctx, OP_T, OP_N, N, Q, Kc - k1, one, E.data() + k1, Kc, 0, X.data() + k1, Kc, 0, one, out.data(), N, 0, 1);
    *skipped += 2.0 * (double)Q * (double)nb * (double)N;
    return out;
  }
  return contract(ctx, X, E, false, false, 1);
}

// RS / AR positions: slab of theta -> partial full-size theta' of this rank (before the collective)
template <typename T>
DTensor<T> Net<T>::heff_partial_from_slab(const DTensor<T>& xs, double* skipped) {
  DTensor<T> X = xs;
  const T one = from_complex<T>(1.0, 0.0);
  std::vector<int64_t> fulld = xs.dims;
  fulld.back() = theta.dims.back();
  size_t i0 = 0;
  double full_numel = 1.0;
  for (auto d : fulld) full_numel *= (double)d;
  if (ctx->opt.skip_identity_sharded && skip_first_identity(X)) { i0 = 1; *skipped += 2.0 * full_numel * (double)envs.at({plan[0].u, plan[0].v}).t.dims[2]; }
  X = run_plan_steps(X, i0, plan.size() - 1);
  // identity channel of the last environment: rows (b in slab, w*) of the slab's contraction index meet the identity,
  // i.e. that block of X is this rank's contribution to columns [shard_lo, shard_hi) of the result
  const Env& le = envs.at({plan.back().u, plan.back().v});
  const int rr = X.rank();
  const int64_t rows = shard_hi - shard_lo, Wd = shard_env.rank() == 3 ? shard_env.dims[1] : 0, N = shard_env.rank() == 3 ? shard_env.dims[2] : 0;
  if (ctx->opt.skip_identity && ctx->opt.skip_identity_sharded && le.ident >= 0 && shard_env.rank() == 3 && le.t.dims[0] == N && rr >= 3 &&
      X.labels[rr - 2] == shard_env.labels[0] && X.labels[rr - 1] == shard_env.labels[1] && X.dims[rr - 2] == rows && X.dims[rr - 1] == Wd) {
    const int64_t Kc = rows * Wd, P = X.numel() / Kc, k0 = rows * le.ident, k1 = rows * (le.ident + 1);
    std::vector<int64_t> od(X.dims.begin(), X.dims.end() - 2);
    std::vector<Label> ol(X.labels.begin(), X.labels.end() - 2);
    od.push_back(N); ol.push_back(shard_env.labels[2]);
    DTensor<T> o2(ctx, od, ol);
    vec_zero<T>(ctx, o2.numel(), o2.data());
    vec_copy<T>(ctx, P * rows, X.data() + P * k0, o2.data() + P * shard_lo);
    if (k0 > 0) gemm<T>(ctx, OP_N, OP_N, P, N, k0, one, X.data(), P, 0, shard_env.data(), Kc, 0, one, o2.data(), P, 0, 1);
    if (k1 < Kc) gemm<T>(ctx, OP_N, OP_N, P, N, Kc - k1, one, X.data() + P * k1, P, 0, shard_env.data() + k1, Kc, 0, one, o2.data(), P, 0, 1);
    *skipped += 2.0 * (double)P * (double)N * (double)N;
    X = o2.noprime();
  } else {
    X = contract(ctx, X, shard_env, false, false, 1).noprime();
  }
  if (X.labels != xs.labels) X = permuted(ctx, X, xs.labels);
  return X;
}

// AG positions: complete input xf (xs = this rank's slab of it) -> this rank's slab of theta'; the first contraction is
// split along the first environment's bra index, no reduction is needed.
template <typename T>
DTensor<T> Net<T>::heff_slab_from_full(const DTensor<T>& xf, const DTensor<T>& xs, double* skipped) {
  const int G = ctx->nranks;
  DTensor<T> X = xf;
  const DTensor<T>& E1 = envs.at({plan[0].u, plan[0].v}).t;
  if (ctx->opt.skip_identity_sharded && skip_first_identity(X, shard_lo, shard_hi, &xs)) {
    *skipped += 2.0 * (double)xf.numel() * (double)E1.dims[2];
  } else {
    X = contract(ctx, xf, E1.last_mode_slab(shard_lo, shard_hi), false, false, 1);
  }
  X = run_plan_steps(X, 1, plan.size() - 1);
  double sk = 0.0;
  if (plan.back().type == 0) X = last_env_contract(X, &sk);
  else X = run_plan_steps(X, plan.size() - 1, plan.size());
  *skipped += sk * (double)G;   // every rank skips its share: whole-job count
  X = X.noprime();
  if (X.labels != xs.labels) X = permuted(ctx, X, xs.labels);
  NSB_REQUIRE(X.dims == xs.dims, NSB_EINTERNAL, "apply_heff_slab: unexpected result shape");
  return X;
}

// Test hook: the arithmetic of the G-rank partition on ONE device -- for every rank r the partial result (RS / AR positions)
// or the result slab (AG positions) is computed exactly as rank r would, and the collective is replaced by a local sum /
// concatenation.  Lets the single-GPU test tier check the multi-GPU arithmetic at any rank count.
template <typename T>
void Net<T>::shard_emulate(int G, void* host_out, int32_t* mode_out) {
  NSB_REQUIRE(theta.valid() && !plan.empty() && G >= 1, NSB_EINVAL, "shard_emulate: call nsb_extract first");
  const int rank0 = ctx->rank, nranks0 = ctx->nranks;
  void* comm0 = ctx->nccl_comm;
  const bool en0 = shard_enabled;
  DTensor<T> acc(ctx, theta.dims, theta.labels);
  vec_zero<T>(ctx, acc.numel(), acc.data());
  int mode = 0;
  try {
    for (int r = 0; r < G; ++r) {
      ctx->rank = r; ctx->nranks = G; ctx->nccl_comm = (void*)1; shard_enabled = true;
      shard_prepare();
      mode = shard_active ? shard_mode : 0;
      double sk = 0.0;
      if (G == 1 || !shard_active) {     // this position is not partitioned at this rank count: every rank applies H_eff on its own
        shard_active = false;
        DTensor<T> y = apply_heff(theta);
        vec_copy<T>(ctx, y.numel(), y.data(), acc.data());
        break;
      }
      if (shard_hi <= shard_lo) continue;
      DTensor<T> xs = theta.last_mode_slab(shard_lo, shard_hi);
      if (shard_mode == 2) {
        DTensor<T> ys = heff_slab_from_full(theta, xs, &sk);
        vec_copy<T>(ctx, ys.numel(), ys.data(), acc.last_mode_slab(shard_lo, shard_hi).data());
      } else {
        DTensor<T> part = heff_partial_from_slab(xs, &sk);
        vec_axpy<T>(ctx, acc.numel(), from_complex<T>(1.0, 0.0), part.data(), acc.data());
      }
    }
  } catch (...) {
    ctx->rank = rank0; ctx->nranks = nranks0; ctx->nccl_comm = comm0; shard_enabled = en0;
    shard_prepare();
    throw;
  }
  ctx->rank = rank0; ctx->nranks = nranks0; ctx->nccl_comm = comm0; shard_enabled = en0;
  shard_prepare();
  if (mode_out) *mode_out = mode;
  NSB_CUDA(cudaMemcpyAsync(host_out, acc.data(), sizeof(T) * acc.numel(), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
}

// Sharded application: slab of theta in, slab of theta' out (RS and AG positions).
template <typename T>
DTensor<T> Net<T>::apply_heff_slab(const DTensor<T>& xs) {
  NSB_REQUIRE(shard_active && (shard_mode == 1 || shard_mode == 2), NSB_EINTERNAL, "apply_heff_slab: position is not slab-sharded");
  const int G = ctx->nranks;
  double skipped = 0.0;
  DTensor<T> out;
  if (shard_mode == 1) {
    DTensor<T> part = heff_partial_from_slab(xs, &skipped);
    out = DTensor<T>(ctx, xs.dims, xs.labels);
    comm_reduce_scatter(part.data(), out.data(), out.numel());
    skipped *= 1.0;   // counted for the whole job inside heff_partial_from_slab
  } else {
    DTensor<T> xf(ctx, theta.dims, theta.labels);
    comm_allgather(xs.data(), xf.data(), xs.numel());
    out = heff_slab_from_full(xf, xs, &skipped);
  }
  ctx->cnt.matvecs++;
  skipped_last_apply = skipped * (ScalarTraits<T>::is_complex ? 4.0 : 1.0);
  return out;
}

template <typename T>
DTensor<T> Net<T>::apply_heff(const DTensor<T>& x) {
  DTensor<T> X = x;
  if (qn_bs() && theta_st && bt_apply_ok && x.labels == theta.labels && x.dims == theta.dims) {
    // QN network: dense local tensor -> symmetry blocks -> sector-batched application -> dense
    BTensor<T> xb = bt_of(x, theta_st), yb;
    if (apply_heff_bt(xb, &yb)) return to_dense<T>(ctx, yb);
    bt_apply_ok = false;
  }
  if (qn_bs()) for (auto& st : plan) if (st.type == 0) env_dense(st.u, st.v);     // dense engine needs dense environments
  if (shard_active && (shard_mode == 1 || shard_mode == 2)) {     // full vector in / out around the slab form
    DTensor<T> os = apply_heff_slab(x.last_mode_slab(shard_lo, shard_hi));
    DTensor<T> out(ctx, x.dims, x.labels);
    comm_allgather(os.data(), out.data(), os.numel());
    return out;
  }
  if (shard_active) {                                              // AR: uneven slabs or the fused epilogue
    DTensor<T> out(ctx, x.dims, x.labels);
    if (shard_hi > shard_lo) {
      X = x.last_mode_slab(shard_lo, shard_hi);
      double skipped = 0.0;    // whole-job count (summed over the ranks' slabs)
      // fused path: the last GEMM writes its tiles into the owners' staging windows over NVLink (P2P stores from the
      // epilogue), the owner sums the partial slabs, one all-gather completes theta'
      std::vector<Label> pl;
      const int G = ctx->nranks;
      const int64_t nb = x.dims.back();
      bool fused = ctx->shard_fused && G <= 8 && nb % G == 0 && ctx->win_local && ctx->win_bytes >= sizeof(T) * (size_t)x.numel();
      if (fused) {
        X = run_plan_steps(X, 0, plan.size() - 1);
        fused = contract_direct_labels(X, shard_env, &pl);
        if (fused) {
          for (auto& l : pl) l = label_setplev(l, 0);
          int shared = 0;
          for (Label l : X.labels) if (shard_env.find(l) >= 0) ++shared;
          fused = (pl == x.labels) && shared == 2 && shard_env.find(X.labels.back()) >= 0 && shard_env.find(X.labels[X.rank() - 2]) >= 0;
          for (int g = 0; g < G && fused; ++g) if (!ctx->win_peer[g]) fused = false;
        }
        if (fused) {
          int64_t Kc = X.dims[X.rank() - 1] * X.dims[X.rank() - 2], P = X.numel() / Kc, N = nb;
          PeerOut po;
          for (int g = 0; g < G; ++g) po.ptr[g] = ctx->win_peer[g];
          po.nranks = G; po.rank = ctx->rank; po.slab_cols = N / G; po.slab_elems = P * (N / G);
          gemm<T>(ctx, OP_N, OP_N, P, N, Kc, from_complex<T>(1.0, 0.0), X.data(), P, 0, shard_env.data(), Kc, 0, zero_<T>(),
                  out.data(), P, 0, 1, GEMM_AUTO, &po);
          // all ranks' stores must have landed before the owner reduces: a one-element all-reduce is the stream-ordered barrier
          nccl_check(nccl_api().AllReduce(ctx->d_scratch + 8, ctx->d_scratch + 8, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream),
                     "ncclAllReduce(barrier)");
          sum_slabs<T>(ctx, (const T*)ctx->win_local, G, po.slab_elems, out.data() + (int64_t)ctx->rank * po.slab_elems);
          comm_allgather(out.data() + (int64_t)ctx->rank * po.slab_elems, out.data(), po.slab_elems);
          ctx->cnt.matvecs++;
          skipped_last_apply = 0.0;
          return out;
        }
        X = contract(ctx, X, shard_env, false, false, 1).noprime();
        if (X.labels != x.labels) X = permuted(ctx, X, x.labels);
        out = X;
        skipped_last_apply = 0.0;
      } else {
        out = heff_partial_from_slab(X, &skipped);
        skipped_last_apply = skipped * (ScalarTraits<T>::is_complex ? 4.0 : 1.0);
      }
    } else {
      vec_zero<T>(ctx, out.numel(), out.data());
    }
    comm_allreduce(out.data(), out.numel());
    ctx->cnt.matvecs++;
    return out;
  }
  size_t i0 = 0, i1 = plan.size();
  double skipped = 0.0;
  if (skip_first_identity(X)) { i0 = 1; skipped += 2.0 * (double)x.numel() * (double)envs.at({plan[0].u, plan[0].v}).t.dims[2]; }
  const bool last_is_env = plan.size() >= 2 && plan.back().type == 0;
  if (last_is_env) i1 = plan.size() - 1;
  X = run_plan_steps(X, i0, i1);
  if (last_is_env) X = last_env_contract(X, &skipped);
  X = X.noprime();
  if (!fit_mode && X.labels != x.labels) X = permuted(ctx, X, x.labels);   // (fitting: the result lives on psi's links)
  ctx->cnt.matvecs++;
  skipped_last_apply = skipped * (ScalarTraits<T>::is_complex ? 4.0 : 1.0);
  return X;
}

// ------------------------------------------------------------------------------------------------
// fitting (src/fitting.jl)
// ------------------------------------------------------------------------------------------------
template <typename T>
void Net<T>::fit_target_upload(int v, int rank, const int32_t* legs, const int64_t* dims, const void* host) {
  NSB_REQUIRE(v >= 0 && v < nverts, NSB_EINVAL, "fit_target_upload: bad vertex");
  NSB_REQUIRE(!qn_on, NSB_EUNSUPPORTED, "fitting is not defined for QN-conserving networks");
  std::vector<Label> labels = decode_legs(rank, legs, false);
  {
    std::vector<Label> a = labels, b = canonical_labels(v);
    std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
    NSB_REQUIRE(a == b, NSB_EINVAL, "fit_target_upload: legs must be the site index and one link per neighbour");
  }
  for (auto& l : labels) if (label_kind(l) == LK_LINK) l = label_setplev(l, 2);
  std::vector<int64_t> d(dims, dims + rank);
  DTensor<T> t(ctx, d, labels);
  NSB_CUDA(cudaMemcpyAsync(t.data(), host, sizeof(T) * t.numel(), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
  if ((int)xket.size() != nverts) xket.assign(nverts, DTensor<T>());
  xket[v] = t;
  fit_mode = true;
  shard_enabled = false;
  envs.clear();
  plan.clear();
}

// Region environment of the overlap network <psi| A |x>: the target's region tensors pushed through the plan
// (environments built with x as the ket layer and conj(psi) as the bra layer, the operator in between).
template <typename T>
DTensor<T> Net<T>::fit_local() {
  NSB_REQUIRE(!pos.empty() && !pos_on_edge, NSB_EINVAL, "fitting: region must be one or two vertices");
  DTensor<T> x = xket[pos[0]];
  for (size_t i = 1; i < pos.size(); ++i) x = contract(ctx, x, xket[pos[i]], false, false, 1);
  return apply_heff(x);
}

template <typename T>
double Net<T>::update_fit() {
  NSB_REQUIRE(fit_mode && theta.valid(), NSB_EINVAL, "update_fit: upload a fitting target and call nsb_extract first");
  return vec_nrm2<T>(ctx, theta.numel(), theta.data());   // n / sqrt(n), n = <local|local> (src/fitting.jl:43-44)
}

template <typename T>
double Net<T>::matvec_flops() {
  NSB_REQUIRE(theta.valid() && !plan.empty(), NSB_EINVAL, "matvec_flops: no local problem (call nsb_extract first)");
  // dry run over labels/dims
  std::vector<Label> lab = theta.labels;
  std::vector<int64_t> dim = theta.dims;
  double flops = 0.0;
  auto numel = [&]() { double n = 1; for (auto d : dim) n *= (double)d; return n; };
  // a merged two-site operator is counted as the two single-site applications it replaces (algorithmic count)
  struct Item { const DTensor<T>* t; };
  std::vector<const DTensor<T>*> seq;
  for (auto& s : plan) {
    if (s.type == 0) seq.push_back(&envs.at({s.u, s.v}).t);
    else if (s.type == 1) seq.push_back(&W[s.v]);
    else { seq.push_back(&W[s.u]); seq.push_back(&W[s.v]); }
  }
  for (const DTensor<T>* Yp : seq) {
    const DTensor<T>& Y = *Yp;
    double kprod = 1, nprod = 1;
    std::vector<Label> nl;
    std::vector<int64_t> nd;
    for (size_t i = 0; i < lab.size(); ++i) {
      if (Y.find(lab[i]) >= 0) kprod *= (double)dim[i];
      else { nl.push_back(lab[i]); nd.push_back(dim[i]); }
    }
    for (int j = 0; j < Y.rank(); ++j)
      if (std::find(lab.begin(), lab.end(), Y.labels[j]) == lab.end()) { nprod *= (double)Y.dims[j]; nl.push_back(Y.labels[j]); nd.push_back(Y.dims[j]); }
    flops += 2.0 * numel() * nprod;   // (numel / kprod) * kprod * nprod multiply-adds
    (void)kprod;
    lab = nl; dim = nd;
  }
  return flops * (ScalarTraits<T>::is_complex ? 4.0 : 1.0);
}

template <typename T>
double Net<T>::matvec_flops_executed() {
  if (qn_bs() && bt_apply_ok && bt_last_apply_flops >= 0.0) return bt_last_apply_flops;   // sector GEMM flops of the last application
  // after an application at this position: what it really skipped; before: the dry run of the same conditions
  return matvec_flops() - (skipped_last_apply >= 0.0 ? skipped_last_apply : skipped_flops(theta));
}

// ------------------------------------------------------------------------------------------------
// subspace expansion (density matrix)
// ------------------------------------------------------------------------------------------------
static int64_t compute_expansion(int64_t current_dim, int64_t basis_size, double expansion_factor, int64_t max_expand, int64_t maxdim) {
  int64_t e = (int64_t)std::ceil(expansion_factor * (double)current_dim);
  e = std::min(max_expand, e);
  e = std::min(basis_size - current_dim, e);
  e = std::min(maxdim - current_dim, e);
  return std::max<int64_t>(0, e);
}

// "ortho" back-end (src/subspace/ortho_subspace.jl:19-77): enlarge the basis of the previous vertex by random
// directions orthogonal to it.  Y = (1 - A A^H) rand(basis, ax), ax = expand_space(basis_size) (:4), truncated SVD of
// Y (cutoff 1e-14, maxdim = compute_expansion) -> Ux, projected once more, A' = [A, Ux] along the bond, the centre
// tensor and the local tensor are multiplied by the expander A'^H A.  (The reference's dispatcher cannot reach this
// method -- it calls `subspace_expand`, the method is named `subspace_expand!` -- here "ortho" selects it.)
template <typename T>
bool Net<T>::expand_ortho(const nsb_trunc& trunc, const nsb_expand& ex) {
  if (pos.empty() || pos_on_edge) return false;
  NSB_REQUIRE(!qn_on, NSB_EUNSUPPORTED, "\"ortho\" subspace expansion is not defined for QN-conserving networks");
  std::vector<int> prev_set;
  for (int p : pos) if (std::find(region.begin(), region.end(), p) == region.end()) prev_set.push_back(p);
  if (prev_set.size() != 1) return false;
  int prev = prev_set[0];
  std::vector<int> nexts;
  for (int v : region) if (eid.count({v, prev})) nexts.push_back(v);
  if (nexts.empty()) return false;
  NSB_REQUIRE(nexts.size() == 1, NSB_EINTERNAL, "expansion: ambiguous next vertex");
  int next = nexts[0];
  const Label a = llink(prev, next);
  DTensor<T> A = psi[prev];
  const int64_t cur = A.dim_of(a), nb = A.numel() / cur;
  const int64_t kexp = compute_expansion(cur, nb, ex.expansion_factor, ex.max_expand, trunc.maxdim);
  const int64_t axd = std::max<int64_t>(nb + 1, (int64_t)std::floor(ex.expansion_factor * (double)nb));   // expand_space
  if (kexp <= 0) return false;
  const T one = from_complex<T>(1.0, 0.0), zero = zero_<T>(), mone = from_complex<T>(-1.0, 0.0);
  std::vector<Label> basis;
  for (Label l : A.labels) if (l != a) basis.push_back(l);
  std::vector<Label> aorder = basis;
  aorder.push_back(a);
  DTensor<T> Ap = permuted(ctx, A, aorder);     // nb x cur
  DevBuf Yb(ctx, sizeof(T) * nb * axd), tmpbuf(ctx, sizeof(T) * cur * std::max(axd, kexp + cur));
  T* Y = (T*)Yb.ptr;
  T* tmp = (T*)tmpbuf.ptr;
  if (expand_probe.ptr && expand_probe_rows == nb && expand_probe_cols == axd) {   // caller's random_itensor(basis_inds, ax)
    vec_copy<T>(ctx, nb * axd, (const T*)expand_probe.ptr, Y);
    ctx->sync();
    expand_probe = DevBuf(); expand_probe_rows = expand_probe_cols = 0;
  } else {
    fill_normal<T>(ctx, Y, nb * axd, expand_seed++, 1.0);
  }
  auto project_out = [&](T* X, int64_t ncols) {   // X <- X - A (A^H X)
    gemm<T>(ctx, OP_C, OP_N, cur, ncols, nb, one, Ap.data(), nb, 0, X, nb, 0, zero, tmp, cur, 0, 1);
    gemm<T>(ctx, OP_N, OP_N, nb, ncols, cur, mone, Ap.data(), nb, 0, tmp, cur, 0, one, X, nb, 0, 1);
  };
  project_out(Y, axd);
  if (vec_nrm2<T>(ctx, nb * axd, Y) <= 1e-15) return false;
  DevBuf Ub, Cb;
  std::vector<double> spec;
  FactorInfo fi = factorize_left<T>(ctx, Y, nb, axd, nb, false, 1e-14, 1, kexp, false, Ub, Cb, spec);
  const int64_t ku = fi.newdim;
  T* U = (T*)Ub.ptr;
  project_out(U, ku);
  // Ax = [A, U] along a;  expander = Ax^H A
  const int64_t nx = cur + ku;
  std::vector<int64_t> axdims;
  for (Label l : basis) axdims.push_back(A.dim_of(l));
  axdims.push_back(nx);
  DTensor<T> Ax(ctx, axdims, aorder);
  vec_copy<T>(ctx, nb * cur, Ap.data(), Ax.data());
  vec_copy<T>(ctx, nb * ku, U, Ax.data() + nb * cur);
  Label aux = make_label(LK_AUX, 2);
  DTensor<T> E(ctx, {nx, cur}, {aux, a});
  gemm<T>(ctx, OP_C, OP_N, nx, cur, nb, one, Ax.data(), nb, 0, Ap.data(), nb, 0, zero, E.data(), nx, 0, 1);
  auto relabel_aux = [&](DTensor<T> t) {
    std::vector<Label> nl = t.labels;
    for (auto& x : nl) if (x == aux) x = a;
    return t.relabeled(nl);
  };
  psi[prev] = Ax;
  canonicalize(prev);
  ver[prev]++;
  psi[next] = relabel_aux(contract(ctx, psi[next], E, false, false, 1));
  canonicalize(next);
  ver[next]++;
  theta = relabel_aux(contract(ctx, theta, E, false, false, 1));
  return true;
}

template <typename T>
bool Net<T>::expand_densitymatrix(const nsb_trunc& trunc, const nsb_expand& ex) {
  if (pos.empty() || pos_on_edge) return false;
  std::vector<int> prev_set;
  for (int p : pos) if (std::find(region.begin(), region.end(), p) == region.end()) prev_set.push_back(p);
  if (prev_set.size() != 1) return false;
  int prev = prev_set[0];
  std::vector<int> nexts;
  for (int v : region) if (eid.count({v, prev})) nexts.push_back(v);
  if (nexts.empty()) return false;
  NSB_REQUIRE(nexts.size() == 1, NSB_EINTERNAL, "expansion: ambiguous next vertex");
  int next = nexts[0];
  const Label a = llink(prev, next);
  DTensor<T> A = psi[prev];
  int64_t cur = A.dim_of(a), nb = A.numel() / cur;
  int64_t kexp = compute_expansion(cur, nb, ex.expansion_factor, ex.max_expand, trunc.maxdim);
  if (kexp <= 0) return false;
  const T one = from_complex<T>(1.0, 0.0), zero = zero_<T>(), mone = from_complex<T>(-1.0, 0.0);

  // sqrt_rho = A * (environments of the previous position that do not touch the new region) * W[prev]
  std::vector<int> ext;
  for (int n : adj[prev]) {
    if (std::find(pos.begin(), pos.end(), n) != pos.end()) continue;      // not an incident edge of the old position
    if (std::find(region.begin(), region.end(), n) != region.end()) continue;
    ext.push_back(n);
  }
  for (int n : ext) make_env(n, prev);   // present already on every plan the reference generates
  DTensor<T> X = A;
  size_t start = 0;
  if (!ext.empty()) { X = contract(ctx, X, env_dense(ext[0], prev), false, false, 1); start = 1; }
  {
    SmallOp<T> op;
    std::vector<int> reg{prev};
    X = apply_small(ctx, op, X, W[prev], w_out_labels(X, W[prev], prev, reg));
  }
  for (size_t i = start; i < ext.size(); ++i) X = contract(ctx, X, env_dense(ext[i], prev), false, false, 1);
  // Operator links toward neighbours of `prev` that were skipped stay open, exactly as in the reference.
  std::vector<Label> basis, basis_p;
  for (Label l : A.labels) if (l != a) { basis.push_back(l); basis_p.push_back(label_setplev(l, 1)); }
  std::vector<Label> sorder = basis_p;
  std::vector<Label> rest;
  for (Label l : X.labels) if (std::find(basis_p.begin(), basis_p.end(), l) == basis_p.end()) rest.push_back(l);
  sorder.insert(sorder.end(), rest.begin(), rest.end());
  DTensor<T> S = permuted(ctx, X, sorder);
  if (S.data() == X.data() && X.data() == A.data()) S = clone(ctx, S);
  int64_t ncol = S.numel() / nb;
  std::vector<Label> aorder = basis;
  aorder.push_back(a);
  DTensor<T> Ap = permuted(ctx, A, aorder);     // nb x cur
  DevBuf tmpbuf(ctx, sizeof(T) * cur * std::max(ncol, kexp + cur));
  T* tmp = (T*)tmpbuf.ptr;
  for (int pass = 0; pass < ex.north_pass; ++pass) {
    gemm<T>(ctx, OP_C, OP_N, cur, ncol, nb, one, Ap.data(), nb, 0, S.data(), nb, 0, zero, tmp, cur, 0, 1);
    gemm<T>(ctx, OP_N, OP_N, nb, ncol, cur, mone, Ap.data(), nb, 0, tmp, cur, 0, one, S.data(), nb, 0, 1);
  }
  DevBuf rho(ctx, sizeof(T) * nb * nb);
  gemm<T>(ctx, OP_N, OP_C, nb, nb, ncol, one, S.data(), nb, 0, S.data(), nb, 0, zero, (T*)rho.ptr, nb, 0, 1);
  DevBuf Ub, Cb;
  std::vector<double> spec;
  std::vector<int64_t> exp_keys;
  FactorInfo fi;
  if (qn_on) {
    std::vector<int64_t> bdims;
    for (Label l : basis) bdims.push_back(A.dim_of(l));
    std::vector<int64_t> keys = multi_keys(prev, basis, bdims, false);   // charge of prev's side of {prev, next}
    fi = factorize_qn((T*)rho.ptr, nb, nb, keys, keys, trunc.cutoff, trunc.mindim, kexp, true, Ub, Cb, exp_keys);
  } else {
    fi = factorize_left<T>(ctx, (T*)rho.ptr, nb, nb, nb, false, trunc.cutoff, trunc.mindim, kexp, true, Ub, Cb, spec);
  }
  int64_t ku = fi.newdim;
  T* U = (T*)Ub.ptr;
  for (int pass = 0; pass < ex.north_pass; ++pass) {
    gemm<T>(ctx, OP_C, OP_N, cur, ku, nb, one, Ap.data(), nb, 0, U, nb, 0, zero, tmp, cur, 0, 1);
    gemm<T>(ctx, OP_N, OP_N, nb, ku, cur, mone, Ap.data(), nb, 0, tmp, cur, 0, one, U, nb, 0, 1);
  }
  gemm<T>(ctx, OP_C, OP_N, cur, ku, nb, one, Ap.data(), nb, 0, U, nb, 0, zero, tmp, cur, 0, 1);
  double ovl = vec_nrm2<T>(ctx, cur * ku, tmp);
  if (ovl > 1e-10) {
    fprintf(stderr, "Warning: |U*A| = %.3E in subspace expansion\n", ovl);
    return false;
  }
  // Ax = [A, U] along a;  expander = Ax^H A
  int64_t nx = cur + ku;
  std::vector<int64_t> axd;
  for (Label l : basis) axd.push_back(A.dim_of(l));
  axd.push_back(nx);
  DTensor<T> Ax(ctx, axd, aorder);
  vec_copy<T>(ctx, nb * cur, Ap.data(), Ax.data());
  vec_copy<T>(ctx, nb * ku, U, Ax.data() + nb * cur);
  Label aux = make_label(LK_AUX, 2);
  DTensor<T> E(ctx, {nx, cur}, {aux, a});
  gemm<T>(ctx, OP_C, OP_N, nx, cur, nb, one, Ax.data(), nb, 0, Ap.data(), nb, 0, zero, E.data(), nx, 0, 1);
  auto relabel_aux = [&](DTensor<T> t) {
    std::vector<Label> nl = t.labels;
    for (auto& x : nl) if (x == aux) x = a;
    return t.relabeled(nl);
  };
  if (qn_on) {   // the new states of the bond carry the charges of their symmetry blocks (prev's side)
    std::vector<int64_t> oldc = side_charge(next, prev);
    std::vector<int64_t> keys(cur + ku);
    for (int64_t i = 0; i < cur; ++i) {
      int64_t k = 0;
      for (int c = 0; c < nq; ++c) k |= ((oldc[i * nq + c] + 32768) & 0xffff) << (16 * c);
      keys[i] = k;
    }
    for (int64_t i = 0; i < ku; ++i) keys[cur + i] = exp_keys[i];
    qn_store_link(prev, next, keys);
  }
  psi[prev] = Ax;
  canonicalize(prev);
  ver[prev]++;
  psi[next] = relabel_aux(contract(ctx, psi[next], E, false, false, 1));
  canonicalize(next);
  ver[next]++;
  theta = relabel_aux(contract(ctx, theta, E, false, false, 1));
  return true;
}

// ------------------------------------------------------------------------------------------------
// the hooks
// ------------------------------------------------------------------------------------------------
template <typename T>
void Net<T>::extract(const int32_t* reg, int nreg, const nsb_trunc* trunc, const nsb_expand* expand, nsb_extract_info* info) {
  NSB_REQUIRE(nreg == 1 || nreg == 2, NSB_EUNSUPPORTED, "Region of this length not currently supported");
  std::vector<int> r(reg, reg + nreg);
  for (int v : r) NSB_REQUIRE(v >= 0 && v < nverts && psi[v].valid() && W[v].valid(), NSB_EINVAL, "extract: bad vertex or missing tensor");
  if (nreg == 2) NSB_REQUIRE(eid.count({r[0], r[1]}), NSB_EINVAL, "extract: two-site region must be an edge");
  nsb_trunc tr = trunc ? *trunc : nsb_trunc{0.0, 1, INT64_MAX};
  int qr_steps, built;
  HostProf hp_all(ctx, "extract");
  {
    PhaseTimer pt(ctx, NSB_T_GAUGE);
    HostProf hp(ctx, "extract.gauge");
    qr_steps = orthogonalize(r);
  }
  region = r;
  {
    PhaseTimer pt(ctx, NSB_T_THETA);
    HostProf hp(ctx, "extract.theta");
    theta = build_theta(r);
  }
  bool expanded = false;
  if (expand && expand->algorithm == NSB_EXPAND_DENSITYMATRIX) {
    PhaseTimer pt(ctx, NSB_T_EXPAND);
    expanded = expand_densitymatrix(tr, *expand);
  } else if (expand && expand->algorithm == NSB_EXPAND_ORTHO) {
    PhaseTimer pt(ctx, NSB_T_EXPAND);
    expanded = expand_ortho(tr, *expand);
  } else if (expand && expand->algorithm != NSB_EXPAND_NONE) {
    throw Error(NSB_EUNSUPPORTED, "Subspace expansion not defined for requested subspace_algorithm");
  }
  {
    PhaseTimer pt(ctx, NSB_T_ENV);
    HostProf hp(ctx, "extract.env");
    built = position(r);
  }
  shard_prepare();
  theta_st = nullptr;
  bt_apply_ok = true;
  bt_last_apply_flops = -1.0;
  if (qn_bs()) {
    bool ok = true;
    for (Label l : theta.labels) if (label_kind(l) == LK_AUX) ok = false;
    if (ok) theta_st = allowed_struct(theta);
  }
  if (fit_mode) {
    PhaseTimer pt(ctx, NSB_T_MATVEC);
    theta = fit_local();
  }
  if (info) {
    info->expanded = expanded ? 1 : 0;
    info->env_builds = built;
    info->qr_steps = qr_steps;
    info->local_rank = theta.rank();
    info->local_numel = theta.numel();
  }
}

template <typename T>
DTensor<T> Net<T>::kvec_start() {
  if (krylov_blocks()) {     // QN network: the Krylov vectors are the flat block storage of the local tensor
    BTensor<T> b = bt_of(theta, theta_st);
    DTensor<T> f;
    f.buf = b.buf; f.dims = {theta_st->total}; f.labels = {make_label(LK_AUX, 7)};
    return f;
  }
  if (!krylov_sharded()) return clone(ctx, theta);
  if (theta_is_slab) return clone(ctx, theta_slab);
  return clone(ctx, theta.last_mode_slab(shard_lo, shard_hi));
}
template <typename T>
void Net<T>::kdot(const DTensor<T>& a, const DTensor<T>& b, double* re_out, double* im_out) {
  if (!krylov_sharded()) { vec_dot<T>(ctx, a.numel(), a.data(), b.data(), re_out, im_out); return; }
  vec_dot_slot<T>(ctx, a.numel(), a.data(), b.data(), 0);
  nccl_check(nccl_api().AllReduce(dot_slot_ptr(ctx, 0), dot_slot_ptr(ctx, 0), 2, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream),
             "ncclAllReduce(dot)");
  double h[2];
  dot_slots_fetch(ctx, 1, h);
  if (re_out) *re_out = h[0];
  if (im_out) *im_out = h[1];
}
template <typename T>
double Net<T>::knrm2(const DTensor<T>& a) {
  double r = 0.0;
  kdot(a, a, &r, nullptr);
  return std::sqrt(r > 0.0 ? r : 0.0);
}
template <typename T>
DTensor<T> Net<T>::kapply(const DTensor<T>& v) {
  if (krylov_blocks()) {
    BTensor<T> xb, yb;
    xb.st = theta_st; xb.buf = v.buf; xb.labels = theta.labels;
    if (!(bt_apply_ok && apply_heff_bt(xb, &yb))) {
      bt_apply_ok = false;
      DTensor<T> yd = apply_heff(to_dense<T>(ctx, xb));
      yb = bt_of(yd, theta_st);
    }
    DTensor<T> f;
    f.buf = yb.buf; f.dims = v.dims; f.labels = v.labels;
    return f;
  }
  return krylov_sharded() ? apply_heff_slab(v) : apply_heff(v);
}
template <typename T>
void Net<T>::kstore_theta(const DTensor<T>& x) {
  if (krylov_blocks()) {
    BTensor<T> xb;
    xb.st = theta_st; xb.buf = x.buf; xb.labels = theta.labels;
    theta = to_dense<T>(ctx, xb);
    theta_is_slab = false;
    return;
  }
  if (krylov_sharded()) { theta_slab = x; theta_is_slab = true; }
  else { theta = x; theta_is_slab = false; }
}
template <typename T>
const char* Net<T>::parallelism_note() const {
  return "Krylov vectors sharded along theta's last bond (reduce-scatter / all-gather per H_eff application, scalar all-reduce per dot); "
         "environment update split over the incoming environment's bra index + all-reduce; factorisation: Gram matrix, back-transformation "
         "and C = U^H theta by column slabs + all-gather, tridiagonalisation and divide & conquer replicated; tensors replicated in HBM";
}

template <typename T>
void Net<T>::update_eigsolve(const nsb_krylov* kp, double* eigval, nsb_solve_info* info) {
  NSB_REQUIRE(theta.valid() && !plan.empty(), NSB_EINVAL, "update_eigsolve: call nsb_extract first");
  nsb_krylov p = kp ? *kp : nsb_krylov{3, 1, 1e-14, 0, 0, 4, 0};
  NSB_REQUIRE(p.maxiter == 1, NSB_EUNSUPPORTED, "eigsolve: only maxiter == 1 (no restart) is implemented, as used by the reference");
  NSB_REQUIRE(p.krylovdim >= 1, NSB_EINVAL, "eigsolve: krylovdim must be >= 1");
  const int kmax = (int)std::min<int64_t>(p.krylovdim, theta.numel());
  std::vector<DTensor<T>> V;
  std::vector<double> alphas, betas;
  double beta = 0.0, beta_prev = 0.0;
  int nmv = 0;
  HostProf hp_all(ctx, "eigsolve");
  DTensor<T> v = kvec_start();      // full local tensor, or this rank's slab of it (multi-GPU)
  const int64_t n = v.numel();
  {
    PhaseTimer pt(ctx, NSB_T_KRYLOV);
    double nrm = knrm2(v);
    NSB_REQUIRE(nrm > 0.0, NSB_EINVAL, "eigsolve: zero initial vector");
    vec_scale<T>(ctx, n, from_complex<T>(1.0 / nrm, 0.0), v.data());
  }
  while (true) {
    V.push_back(v);
    DTensor<T> w;
    {
      PhaseTimer pt(ctx, NSB_T_MATVEC);
      HostProf hp(ctx, "eigsolve.matvec");
      w = kapply(v);
      ++nmv;
    }
    PhaseTimer pt(ctx, NSB_T_KRYLOV);
    if (w.data() == v.data()) w = clone(ctx, w);
    double ar, ai;
    kdot(v, w, &ar, &ai);
    double alpha = ar;
    vec_axpy<T>(ctx, n, from_complex<T>(-alpha, 0.0), v.data(), w.data());
    if (V.size() > 1) vec_axpy<T>(ctx, n, from_complex<T>(-beta_prev, 0.0), V[V.size() - 2].data(), w.data());
    // full re-orthogonalisation against the whole basis (second modified Gram-Schmidt pass)
    for (size_t i = 0; i < V.size(); ++i) {
      double cr, ci;
      kdot(V[i], w, &cr, &ci);
      vec_axpy<T>(ctx, n, from_complex<T>(-cr, -ci), V[i].data(), w.data());
      if (i + 1 == V.size()) alpha += cr;
    }
    alphas.push_back(alpha);
    beta = knrm2(w);
    int K = (int)V.size();
    if (K == kmax || beta <= p.tol) break;
    if (p.eager) {
      // KrylovKit `eager`: test the wanted Ritz pair after every expansion and leave only when its residual
      // |beta y_K| has converged; otherwise keep expanding up to krylovdim
      std::vector<double> Te((size_t)K * K, 0.0), ev, evec;
      for (int i = 0; i < K; ++i) Te[i + (size_t)i * K] = alphas[i];
      for (int i = 0; i + 1 < K; ++i) { Te[i + (size_t)(i + 1) * K] = betas[i]; Te[(i + 1) + (size_t)i * K] = betas[i]; }
      host_sym_eig(K, Te, ev, evec);
      const int id = (p.which == 0) ? 0 : K - 1;
      if (std::fabs(beta * evec[(K - 1) + (size_t)id * K]) <= p.tol) break;
    }
    betas.push_back(beta);
    beta_prev = beta;
    vec_scale<T>(ctx, n, from_complex<T>(1.0 / beta, 0.0), w.data());
    v = w;
  }
  PhaseTimer pt(ctx, NSB_T_KRYLOV);
  int K = (int)V.size();
  std::vector<double> Tm((size_t)K * K, 0.0), evals, evecs;
  for (int i = 0; i < K; ++i) Tm[i + (size_t)i * K] = alphas[i];
  for (int i = 0; i + 1 < K; ++i) { Tm[i + (size_t)(i + 1) * K] = betas[i]; Tm[(i + 1) + (size_t)i * K] = betas[i]; }
  host_sym_eig(K, Tm, evals, evecs);
  int idx = (p.which == 0) ? 0 : K - 1;
  std::vector<const T*> ptrs(K);
  std::vector<T> coef(K);
  for (int i = 0; i < K; ++i) { ptrs[i] = V[i].data(); coef[i] = from_complex<T>(evecs[i + (size_t)idx * K], 0.0); }
  DTensor<T> x(ctx, V[0].dims, V[0].labels);
  vec_lincomb<T>(ctx, n, K, ptrs.data(), coef.data(), x.data());
  kstore_theta(x);
  if (eigval) *eigval = evals[idx];
  if (info) {
    info->nmatvec = nmv;
    info->krylovdim = K;
    info->residual = std::fabs(beta * evecs[(K - 1) + (size_t)idx * K]);
    info->converged = info->residual <= p.tol ? 1 : 0;
    info->reserved = 0;
  }
}

template <typename T> struct CplxScalar;
template <> struct CplxScalar<double> {
  static bool ok(double, double im) { return im == 0.0; }
  static double make(double re, double) { return re; }
};
template <> struct CplxScalar<cdouble> {
  static bool ok(double, double) { return true; }
  static cdouble make(double re, double im) { return make_cuDoubleComplex(re, im); }
};

// Local exponential solvers on an arbitrary device operator H: x -> H x (same vector shape).
//   solver RK: src/local_solvers/runge_kutta.jl:2-25;  solver KRYLOV: KrylovKit.exponentiate (expintegrator, p = 1).
template <typename T>
DTensor<T> Net<T>::exp_solve(const std::function<DTensor<T>(const DTensor<T>&)>& Hraw, std::complex<double> t, const DTensor<T>& x0,
                             int solver, const nsb_krylov* kp, int* nmv_out, int* lastK_out, int* conv_out, double* err_out, bool edge_local) {
  const int64_t n = x0.numel();
  typedef std::complex<double> C;
  auto S = [&](C z) { return CplxScalar<T>::make(z.real(), z.imag()); };
  int nmv = 0;
  auto H = [&](const DTensor<T>& x) {
    PhaseTimer pt(ctx, NSB_T_MATVEC);
    DTensor<T> y = Hraw(x);
    if (y.data() == x.data()) y = clone(ctx, y);
    ++nmv;
    return y;
  };
  DTensor<T> result;
  int lastK = 0, converged = 1;
  double totalerr = 0.0;
  if (solver == NSB_SOLVER_RK) {
    int order = kp ? kp->rk_order : 4;
    if (order == 0) order = 4;
    if (order == 4) {
      DTensor<T> k1 = H(x0);
      DTensor<T> k2 = H(k1);  { PhaseTimer pt(ctx, NSB_T_KRYLOV); vec_scale<T>(ctx, n, S(t / 2.0), k2.data()); vec_axpy<T>(ctx, n, S(1.0), k1.data(), k2.data()); }
      DTensor<T> k3 = H(k2);  { PhaseTimer pt(ctx, NSB_T_KRYLOV); vec_scale<T>(ctx, n, S(t / 2.0), k3.data()); vec_axpy<T>(ctx, n, S(1.0), k1.data(), k3.data()); }
      DTensor<T> k4 = H(k3);  { PhaseTimer pt(ctx, NSB_T_KRYLOV); vec_scale<T>(ctx, n, S(t), k4.data()); vec_axpy<T>(ctx, n, S(1.0), k1.data(), k4.data()); }
      PhaseTimer pt(ctx, NSB_T_KRYLOV);
      const T* ptrs[5] = {x0.data(), k1.data(), k2.data(), k3.data(), k4.data()};
      T coef[5] = {S(1.0), S(t / 6.0), S(t / 3.0), S(t / 3.0), S(t / 6.0)};
      result = DTensor<T>(ctx, x0.dims, x0.labels);
      vec_lincomb<T>(ctx, n, 5, ptrs, coef, result.data());
    } else if (order == 2) {
      DTensor<T> h1 = H(x0);
      DTensor<T> h2 = H(h1);
      PhaseTimer pt(ctx, NSB_T_KRYLOV);
      const T* ptrs[3] = {x0.data(), h1.data(), h2.data()};
      T coef[3] = {S(1.0), S(t), S(t * t / 2.0)};
      result = DTensor<T>(ctx, x0.dims, x0.labels);
      vec_lincomb<T>(ctx, n, 3, ptrs, coef, result.data());
    } else {
      throw Error(NSB_EINVAL, "For runge_kutta_solver, must specify `order` keyword (2 or 4)");
    }
  } else {
    NSB_REQUIRE(solver == NSB_SOLVER_KRYLOV, NSB_EINVAL, "update_exp: unknown solver");
    nsb_krylov p = kp ? *kp : nsb_krylov{30, 100, 1e-12, 0, 1, 4, 0};
    const double tau = std::abs(t);
    if (tau == 0.0) {
      result = clone(ctx, x0);
    } else {
      const C sgn = t / tau;
      const double eta = p.tol / tau, gamma = 0.8;
      double tau0 = 0.0, dtau = tau;
      int numiter = 1;
      converged = 0;
      DTensor<T> w0 = clone(ctx, x0);
      DTensor<T> w1 = H(w0);
      const int kmax = (int)std::min<int64_t>(p.krylovdim, n);
      bool done = false;
      while (!done) {
        double beta = edge_local ? vec_nrm2<T>(ctx, n, w1.data()) : knrm2(w1);
        if (beta < p.tol) { converged = 1; break; }
        std::vector<DTensor<T>> V;
        std::vector<double> alphas, betas;
        DTensor<T> r;
        double resnorm = 0.0;
        { DTensor<T> v0 = clone(ctx, w1); vec_scale<T>(ctx, n, from_complex<T>(1.0 / beta, 0.0), v0.data()); V.push_back(v0); }
        auto expand = [&]() {
          DTensor<T>& v = V.back();
          DTensor<T> w = H(v);
          PhaseTimer pt(ctx, NSB_T_KRYLOV);
          double ar, ai;
          if (edge_local) vec_dot<T>(ctx, n, v.data(), w.data(), &ar, &ai); else kdot(v, w, &ar, &ai);
          double a = ar;
          vec_axpy<T>(ctx, n, from_complex<T>(-a, 0.0), v.data(), w.data());
          if (V.size() > 1) vec_axpy<T>(ctx, n, from_complex<T>(-betas.back(), 0.0), V[V.size() - 2].data(), w.data());
          for (size_t i = 0; i < V.size(); ++i) {
            double cr, ci;
            if (edge_local) vec_dot<T>(ctx, n, V[i].data(), w.data(), &cr, &ci); else kdot(V[i], w, &cr, &ci);
            vec_axpy<T>(ctx, n, from_complex<T>(-cr, -ci), V[i].data(), w.data());
            if (i + 1 == V.size()) a += cr;
          }
          alphas.push_back(a);
          r = w;
          resnorm = edge_local ? vec_nrm2<T>(ctx, n, w.data()) : knrm2(w);
        };
        expand();
        while (true) {
          int K = (int)V.size();
          lastK = K;
          bool stepped = false;
          double step = 0.0, eps = 0.0, omega = 0.0, q = K / 2.0;
          std::vector<C> E;
          auto small_exp = [&](double dt) {
            int m = K + 2;
            E.assign((size_t)m * m, C(0));
            for (int i = 0; i < K; ++i) E[i + (size_t)i * m] = sgn * dt * alphas[i];
            for (int i = 0; i + 1 < K; ++i) { E[i + (size_t)(i + 1) * m] = sgn * dt * betas[i]; E[(i + 1) + (size_t)i * m] = sgn * dt * betas[i]; }
            E[0 + (size_t)K * m] = 1.0;
            E[K + (size_t)(K + 1) * m] = 1.0;
            host_expm_complex(m, E);
            return std::abs(dt * beta * resnorm * E[(K - 1) + (size_t)(K + 1) * m]);
          };
          if (K == kmax) {
            dtau = std::min(dtau, tau - tau0);
            eps = small_exp(dtau);
            omega = eps / (dtau * eta);
            while (omega > 1.0) {
              double eps_prev = eps, dtau_prev = dtau;
              dtau *= std::pow(gamma / omega, 1.0 / (q + 1.0));
              eps = small_exp(dtau);
              omega = eps / (dtau * eta);
              if (eps <= 0.0) break;
              q = std::max(0.0, std::log(eps / eps_prev) / std::log(dtau / dtau_prev) - 1.0);
            }
            step = dtau;
            stepped = true;
          } else if (resnorm <= (tau - tau0) * eta || p.eager) {
            step = tau - tau0;
            eps = small_exp(step);
            omega = eps / (step * eta);
            if (omega < 1.0) stepped = true;
          }
          if (stepped) {
            PhaseTimer pt(ctx, NSB_T_KRYLOV);
            totalerr += eps;
            int m = K + 2;
            const C f = beta * sgn * step;
            std::vector<const T*> ptrs;
            std::vector<T> coef;
            for (int i = 0; i < K; ++i) { ptrs.push_back(V[i].data()); coef.push_back(S(f * E[i + (size_t)K * m])); }
            ptrs.push_back(r.data()); coef.push_back(S(f * E[(K - 1) + (size_t)(K + 1) * m]));
            ptrs.push_back(w0.data()); coef.push_back(S(1.0));
            DTensor<T> nw(ctx, x0.dims, x0.labels);
            vec_lincomb<T>(ctx, n, (int)ptrs.size(), ptrs.data(), coef.data(), nw.data());
            w0 = nw;
            tau0 += step;
            if (K == kmax && omega < gamma) dtau *= std::pow(gamma / std::max(omega, 1e-300), 1.0 / (q + 1.0));
          }
          if (tau0 >= tau * (1.0 - 1e-15)) { converged = 1; done = true; break; }
          if (stepped) break;
          if (K < kmax && resnorm > 0.0) {
            PhaseTimer pt(ctx, NSB_T_KRYLOV);
            betas.push_back(resnorm);
            DTensor<T> nv = r;
            vec_scale<T>(ctx, n, from_complex<T>(1.0 / resnorm, 0.0), nv.data());
            V.push_back(nv);
          } else break;
          expand();
        }
        if (done) break;
        if (numiter == p.maxiter) { converged = 0; break; }
        ++numiter;
        w1 = H(w0);
      }
      result = w0;
    }
  }
  if (nmv_out) *nmv_out += nmv;
  if (lastK_out) *lastK_out = lastK;
  if (conv_out) *conv_out = converged;
  if (err_out) *err_out += totalerr;
  return result;
}

template <typename T>
void Net<T>::update_exp(double tre, double tim, int solver, const nsb_krylov* kp, int nsites, int next_vertex, nsb_solve_info* info) {
  NSB_REQUIRE(theta.valid() && !plan.empty(), NSB_EINVAL, "update_exp: call nsb_extract first");
  NSB_REQUIRE(CplxScalar<T>::ok(tre, tim), NSB_EINVAL, "update_exp: complex exponent needs a complex128 network");
  typedef std::complex<double> C;
  const C t(tre, tim);
  int nmv = 0, lastK = 0, conv = 1;
  double err = 0.0;
  // forward step (src/applyexp.jl:28)
  {
    DTensor<T> x0 = kvec_start();
    kstore_theta(exp_solve([&](const DTensor<T>& x) { return kapply(x); }, t, x0, solver, kp, &nmv, &lastK, &conv, &err, false));
  }
  if (nsites == 1 && next_vertex >= 0) {
    ensure_theta_full();
    // src/applyexp.jl:30-42: QR-split the evolved site tensor toward the next region, move the projected
    // operator onto the edge (0-site H_eff = the two environments), evolve R backward by -t, recombine.
    NSB_REQUIRE(region.size() == 1, NSB_EINVAL, "update_exp: nsites == 1 needs a one-site region");
    const int v1 = region[0], v2 = next_vertex;
    NSB_REQUIRE(eid.count({v1, v2}), NSB_EINVAL, "update_exp: next_vertex is not a neighbour of the region");
    const Label l = llink(v1, v2);
    std::vector<Label> order;
    for (Label x : theta.labels) if (x != l) order.push_back(x);
    order.push_back(l);
    DTensor<T> Ap = permuted(ctx, theta, order);
    if (Ap.data() == theta.data()) Ap = clone(ctx, theta);
    const int64_t cols = Ap.dims.back(), rows = Ap.numel() / cols, k = std::min(rows, cols);
    std::vector<int64_t> qd(Ap.dims.begin(), Ap.dims.end() - 1);
    qd.push_back(k);
    const Label ax = make_label(LK_AUX, 3, 0);
    DTensor<T> Q, R;                                               // Q: [others..., l] with dim(l) = k;  R: [ax, l]
    std::vector<int64_t> saved_link;
    int saved_side = -1;
    {
      PhaseTimer pt(ctx, NSB_T_GAUGE);
      int64_t kk = k;
      if (qn_on) {
        int e = eid.at({v1, v2});
        saved_link = qn_link[e]; saved_side = qn_side[e];
        std::vector<Label> others(order.begin(), order.end() - 1);
        std::vector<int64_t> odims(Ap.dims.begin(), Ap.dims.end() - 1);
        std::vector<int64_t> rk = multi_keys(v1, others, odims, false), ck = multi_keys(v2, {l}, {cols}, false), newk;
        DevBuf Qb, Rb;
        qr_qn(Ap.data(), rows, cols, rk, ck, Qb, Rb, &kk, newk);
        qd.back() = kk;
        Q.buf = std::make_shared<DevBuf>(std::move(Qb)); Q.dims = qd; Q.labels = order;
        R.buf = std::make_shared<DevBuf>(std::move(Rb)); R.dims = {kk, cols}; R.labels = {ax, l};
        qn_store_link(v1, v2, newk);
      } else {
        Q = DTensor<T>(ctx, qd, order);
        R = DTensor<T>(ctx, {k, cols}, {ax, l});
        qr_thin<T>(ctx, Ap.data(), rows, cols, rows, Q.data(), rows, R.data(), k);
      }
    }
    DTensor<T> E1, E2;
    {
      PhaseTimer pt(ctx, NSB_T_ENV);
      psi[v1] = Q;                   // the inserter overwrites psi[v1] with the recombined tensor afterwards
      canonicalize(v1);
      ver[v1]++;
      envs.erase({v1, v2});
      make_env(v2, v1);              // present already (incident to the current region)
      make_env(v1, v2);
      E1 = env_dense(v1, v2);        // [l(k), op, l'(k)]  ->  relabel its link to the auxiliary QR index
      std::vector<Label> nl = E1.labels;
      for (auto& x : nl) if (label_kind(x) == LK_LINK) x = make_label(LK_AUX, 3, label_plev(x));
      E1 = E1.relabeled(nl);
      E2 = env_dense(v2, v1);        // [l, op, l']
    }
    auto Hedge = [&](const DTensor<T>& x) {      // src/operator_map.jl:17-20 (on-edge branch)
      DTensor<T> y = contract(ctx, x, E2, false, false, 1);        // [ax, op, l']
      y = contract(ctx, y, E1, false, false, 1).noprime();         // [ax', l'] -> [ax, l]
      if (y.labels != x.labels) y = permuted(ctx, y, x.labels);
      ctx->cnt.matvecs++;
      return y;
    };
    DTensor<T> Rt = exp_solve(Hedge, -t, R, solver, kp, &nmv, &lastK, &conv, &err, true);   // replicated small problem on the edge
    // local_state = psi[v1] * R_t
    std::vector<Label> ql = order;
    ql.back() = ax;
    DTensor<T> th = contract(ctx, Q.relabeled(ql), Rt, false, false, 1);
    if (th.labels != theta.labels) th = permuted(ctx, th, theta.labels);
    theta = th;
    if (qn_on) {   // bond of theta is the original one
      int e = eid.at({v1, v2});
      qn_link[e] = saved_link; qn_side[e] = saved_side;
      if ((int)link_mode.size() > e) link_mode[e] = nullptr;
    }
  }
  if (info) { info->nmatvec = nmv; info->krylovdim = lastK; info->converged = conv; info->residual = err; info->reserved = 0; }
}

template <typename T>
void Net<T>::insert(const nsb_trunc* trunc, int normalize, int set_ortho, nsb_insert_info* info) {
  NSB_REQUIRE(theta.valid() && !region.empty(), NSB_EINVAL, "insert: call nsb_extract first");
  ensure_theta_full();
  nsb_trunc tr = trunc ? *trunc : nsb_trunc{0.0, 1, INT64_MAX};
  nsb_insert_info out{0, 0.0, 0, 0};
  int last = region.back();
  if (region.size() == 1) {
    psi[last] = theta;
    canonicalize(last);
    ver[last]++;
    out.newdim = 0;
  } else {
    PhaseTimer pt(ctx, NSB_T_FACTORIZE);
    HostProf hp(ctx, "insert.two_site");
    int v1 = region[0], v2 = region[1];
    Label bond = llink(v1, v2);
    std::vector<Label> left, right;
    for (Label l : theta.labels) {
      if (psi[v1].find(l) >= 0) left.push_back(l);
      else right.push_back(l);
    }
    // matricisation without a permute when theta is [left..., right...] or [right..., left...]
    bool lr = true, rl = true;
    for (size_t i = 0; i < theta.labels.size(); ++i) {
      bool isl = std::find(left.begin(), left.end(), theta.labels[i]) != left.end();
      if (i < left.size() && !isl) lr = false;
      if (i >= left.size() && isl) lr = false;
      if (i < right.size() && isl) rl = false;
      if (i >= right.size() && !isl) rl = false;
    }
    DTensor<T> M = theta;
    bool trans = false;
    if (lr) trans = false;
    else if (rl) trans = true;
    else {
      std::vector<Label> order = left;
      order.insert(order.end(), right.begin(), right.end());
      M = permuted(ctx, theta, order);
    }
    int64_t rows = 1, cols = 1;
    for (Label l : left) rows *= theta.dim_of(l);
    for (Label l : right) cols *= theta.dim_of(l);
    DevBuf Ub, Cb;
    std::vector<double> spec;
    FactorInfo fi;
    if (qn_on) {
      if (trans) {   // block-wise path wants the contiguous [left..., right...] matricisation
        std::vector<Label> order = left;
        order.insert(order.end(), right.begin(), right.end());
        M = permuted(ctx, theta, order);
        trans = false;
      }
      std::vector<int64_t> ld_, rd_;
      for (Label l : left) ld_.push_back(theta.dim_of(l));
      for (Label l : right) rd_.push_back(theta.dim_of(l));
      std::vector<int64_t> rk = multi_keys(v1, left, ld_, false);      // charge of v1's side of the new bond
      std::vector<int64_t> ck = multi_keys(v2, right, rd_, true);      // total - (charge of v2's side)
      std::vector<int64_t> newk;
      fi = factorize_qn(M.data(), rows, cols, rk, ck, tr.cutoff, tr.mindim, tr.maxdim, false, Ub, Cb, newk);
      qn_store_link(v1, v2, newk);
    } else {
      FactorDist dist;
      const bool use_dist = shard_enabled && ctx->nranks > 1 && ctx->nccl_comm;
      if (use_dist) {
        dist.rank = ctx->rank; dist.nranks = ctx->nranks; dist.allow_c_transposed = true;
        dist.allgather_inplace = [this](void* buf, size_t bytes) {
          nccl_check(nccl_api().AllGather((const char*)buf + bytes * (size_t)ctx->rank, buf, bytes / sizeof(double), ncclDouble,
                                          (ncclComm_t)ctx->nccl_comm, ctx->stream), "ncclAllGather(factorize)");
          ctx->cnt.kernel_launches++;
        };
      }
      fi = factorize_left<T>(ctx, M.data(), rows, cols, trans ? cols : rows, trans, tr.cutoff, tr.mindim, tr.maxdim, false, Ub, Cb, spec,
                             use_dist ? &dist : nullptr);
    }
    int64_t k = fi.newdim;
    std::vector<int64_t> ud, cd;
    std::vector<Label> ul = left, cl;
    for (Label l : left) ud.push_back(theta.dim_of(l));
    ud.push_back(k); ul.push_back(bond);
    if (fi.c_transposed) {        // C^T: [right..., bond]
      for (Label l : right) { cd.push_back(theta.dim_of(l)); cl.push_back(l); }
      cd.push_back(k); cl.push_back(bond);
    } else {
      cd.push_back(k); cl.push_back(bond);
      for (Label l : right) { cd.push_back(theta.dim_of(l)); cl.push_back(l); }
    }
    DTensor<T> Ut, Ct;
    Ut.buf = std::make_shared<DevBuf>(std::move(Ub)); Ut.dims = ud; Ut.labels = ul;
    Ct.buf = std::make_shared<DevBuf>(std::move(Cb)); Ct.dims = cd; Ct.labels = cl;
    uint64_t pb = ctx->cnt.permute_bytes;
    psi[v1] = Ut; canonicalize(v1); ver[v1]++;
    psi[v2] = Ct; canonicalize(v2); ver[v2]++;
    ctx->cnt.permute_bytes = pb;   // write-back of the factors in canonical (first link, site, other links) order
    out.newdim = k;
    out.truncerr = fi.truncerr;
    out.decomp = fi.decomp;
    out.jacobi_sweeps = fi.sweeps;
  }
  if (set_ortho) ortho = {last};
  if (normalize) {
    double nrm = vec_nrm2<T>(ctx, psi[last].numel(), psi[last].data());
    if (nrm > 0) vec_scale<T>(ctx, psi[last].numel(), from_complex<T>(1.0 / nrm, 0.0), psi[last].data());
    ver[last]++;
  }
  theta = DTensor<T>();
  if (info) *info = out;
}

template <typename T>
void Net<T>::local_info(int32_t* rank, int32_t* legs, int64_t* dims) {
  NSB_REQUIRE(theta.valid(), NSB_EINVAL, "local_info: no local tensor");
  *rank = theta.rank();
  if (legs) encode_legs(theta.labels, legs);
  if (dims) for (int i = 0; i < theta.rank(); ++i) dims[i] = theta.dims[i];
}
template <typename T>
void Net<T>::local_download(void* host) {
  NSB_REQUIRE(theta.valid(), NSB_EINVAL, "local_download: no local tensor");
  ensure_theta_full();
  NSB_CUDA(cudaMemcpyAsync(host, theta.data(), sizeof(T) * theta.numel(), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
}
template <typename T>
void Net<T>::local_upload(const void* host) {
  NSB_REQUIRE(theta.valid(), NSB_EINVAL, "local_upload: no local tensor");
  theta_is_slab = false;
  NSB_CUDA(cudaMemcpyAsync(theta.data(), host, sizeof(T) * theta.numel(), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
}
template <typename T>
void Net<T>::matvec_host(const void* in, void* outp) {
  NSB_REQUIRE(theta.valid() && !plan.empty(), NSB_EINVAL, "matvec: call nsb_extract first");
  DTensor<T> x(ctx, theta.dims, theta.labels);
  NSB_CUDA(cudaMemcpyAsync(x.data(), in, sizeof(T) * x.numel(), cudaMemcpyHostToDevice, ctx->stream));
  DTensor<T> y = apply_heff(x);
  NSB_CUDA(cudaMemcpyAsync(outp, y.data(), sizeof(T) * y.numel(), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
}
template <typename T>
void Net<T>::shard_range(int64_t* lo, int64_t* hi, int64_t* last_dim) {
  NSB_REQUIRE(theta.valid(), NSB_EINVAL, "shard_range: call nsb_extract first");
  const int64_t d = theta.dims.back();
  if (last_dim) *last_dim = d;
  if (krylov_sharded()) { if (lo) *lo = shard_lo; if (hi) *hi = shard_hi; }
  else { if (lo) *lo = 0; if (hi) *hi = d; }
}
template <typename T>
void Net<T>::matvec_host_slab(const void* in, void* outp) {
  NSB_REQUIRE(theta.valid() && !plan.empty(), NSB_EINVAL, "matvec: call nsb_extract first");
  if (!krylov_sharded()) { matvec_host(in, outp); return; }
  // slab in, slab out: what the sharded Krylov solvers do, with this rank's 1 / G of the host <-> device traffic
  DTensor<T> full(ctx, theta.dims, theta.labels);                 // (only its slab is touched: a view needs a parent)
  DTensor<T> xs = full.last_mode_slab(shard_lo, shard_hi);
  NSB_CUDA(cudaMemcpyAsync(xs.data(), in, sizeof(T) * xs.numel(), cudaMemcpyHostToDevice, ctx->stream));
  DTensor<T> os = apply_heff_slab(xs);
  NSB_CUDA(cudaMemcpyAsync(outp, os.data(), sizeof(T) * os.numel(), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
}
template <typename T>
void Net<T>::matvec_device(int reps, void* host_out) {
  NSB_REQUIRE(theta.valid() && !plan.empty(), NSB_EINVAL, "matvec: call nsb_extract first");
  if (krylov_sharded()) {   // what the sharded Krylov solvers do: slab in, slab out
    DTensor<T> xs = theta_is_slab ? theta_slab : theta.last_mode_slab(shard_lo, shard_hi), os;
    for (int i = 0; i < reps; ++i) os = apply_heff_slab(xs);
    if (host_out && os.valid()) {
      last_out = DTensor<T>(ctx, theta.dims, theta.labels);
      comm_allgather(os.data(), last_out.data(), os.numel());
    } else {
      last_out = os;
      host_out = nullptr;
    }
  } else if (krylov_blocks() && bt_apply_ok) {   // QN network: as the Krylov solvers do -- block vector in, block vector out
    DTensor<T> v = kvec_start(), w;
    for (int i = 0; i < reps; ++i) w = kapply(v);
    if (host_out && w.valid()) {
      BTensor<T> yb;
      yb.st = theta_st; yb.buf = w.buf; yb.labels = theta.labels;
      last_out = to_dense<T>(ctx, yb);
    } else {
      last_out = w;
      host_out = nullptr;
    }
  } else {
    for (int i = 0; i < reps; ++i) last_out = apply_heff(theta);
  }
  if (host_out && last_out.valid()) {
    NSB_CUDA(cudaMemcpyAsync(host_out, last_out.data(), sizeof(T) * last_out.numel(), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
  }
}
template <typename T>
double Net<T>::norm() {
  NSB_REQUIRE(ortho.size() == 1, NSB_EINVAL, "norm: needs a single-vertex orthogonality centre");
  return vec_nrm2<T>(ctx, psi[ortho[0]].numel(), psi[ortho[0]].data());
}

// range_finder(linear_map, random_vector) with linear_map = x -> optimal_map(P, x) at the current position
// (src/sketched_linear_algebra/range_finder.jl:54-64 over the closure of src/eigsolve.jl:22)
template <typename T>
int64_t Net<T>::range_finder_heff(const void* probes_host, uint64_t seed, int64_t max_rank, int oversample, int north_pass, double thr,
                                  double cutoff, void* Qhost) {
  NSB_REQUIRE(theta.valid() && !plan.empty(), NSB_EINVAL, "range_finder: call nsb_extract first");
  if (max_rank <= 0) return 0;
  const int64_t n = theta.numel();
  const int64_t sketch = std::min(std::min(max_rank, n) + (int64_t)oversample, n);
  DevBuf dQ(ctx, sizeof(T) * (size_t)n * sketch), dP(ctx, probes_host ? sizeof(T) * (size_t)n * sketch : 0);
  if (probes_host) NSB_CUDA(cudaMemcpyAsync(dP.ptr, probes_host, sizeof(T) * (size_t)n * sketch, cudaMemcpyHostToDevice, ctx->stream));
  RangeMap<T> map = [&](const T* Om, T* Y, int64_t p) {
    for (int64_t j = 0; j < p; ++j) {
      DTensor<T> x(ctx, theta.dims, theta.labels);
      vec_copy<T>(ctx, n, Om + j * n, x.data());
      DTensor<T> y = apply_heff(x);
      vec_copy<T>(ctx, n, y.data(), Y + j * n);
    }
  };
  const int64_t have = range_finder_blocked<T>(ctx, n, n, map, probes_host ? (const T*)dP.ptr : nullptr, seed, max_rank, oversample, north_pass,
                                               thr, cutoff, (T*)dQ.ptr);
  if (have > 0) NSB_CUDA(cudaMemcpyAsync(Qhost, dQ.ptr, sizeof(T) * (size_t)n * have, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  return have;
}

template <typename T>
void Net<T>::expand_set_probe(int64_t rows, int64_t cols, const void* host) {
  expand_probe = DevBuf(ctx, sizeof(T) * (size_t)rows * cols);
  NSB_CUDA(cudaMemcpyAsync(expand_probe.ptr, host, sizeof(T) * (size_t)rows * cols, cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
  expand_probe_rows = rows; expand_probe_cols = cols;
}

template struct Net<double>;
template struct Net<cdouble>;

}  // namespace nsb
