#include "tensor.h"

#include <algorithm>

namespace nsb {

template <typename T>
DTensor<T> clone(Ctx* ctx, const DTensor<T>& A) {
  DTensor<T> out(ctx, A.dims, A.labels);
  vec_copy<T>(ctx, A.numel(), A.data(), out.data());
  return out;
}

template <typename T>
DTensor<T> permuted(Ctx* ctx, const DTensor<T>& A, const std::vector<Label>& order, bool conj) {
  NSB_REQUIRE((int)order.size() == A.rank(), NSB_EINTERNAL, "permuted: rank mismatch");
  std::vector<int> perm(A.rank());
  std::vector<int64_t> od(A.rank());
  bool ident = true;
  for (int i = 0; i < A.rank(); ++i) {
    perm[i] = A.find(order[i]);
    NSB_REQUIRE(perm[i] >= 0, NSB_EINTERNAL, "permuted: label missing");
    od[i] = A.dims[perm[i]];
    if (perm[i] != i) ident = false;
  }
  if (ident && !conj) return A;
  DTensor<T> out(ctx, od, order);
  // a permutation that only moves extent-1 modes is a plain copy
  bool trivial = true;
  {
    int last = -1;
    for (int i = 0; i < A.rank(); ++i) {
      if (od[i] == 1) continue;
      if (perm[i] < last) { trivial = false; break; }
      last = perm[i];
    }
  }
  if (trivial && !conj) {
    vec_copy<T>(ctx, A.numel(), A.data(), out.data());
    return out;
  }
  permute<T>(ctx, A.data(), out.data(), A.rank(), A.dims.data(), perm.data(), conj);
  return out;
}

namespace {

struct BlockInfo {
  std::vector<int> posX, posY;   // positions of shared labels (in X order)
  bool x_contig = false, y_contig = false, y_prefix = false, y_suffix = false, same_order = false;
  int x_start = 0;
};

template <typename T>
BlockInfo analyze(const DTensor<T>& X, const DTensor<T>& Y) {
  BlockInfo b;
  for (int i = 0; i < X.rank(); ++i) {
    int j = Y.find(X.labels[i]);
    if (j >= 0) { b.posX.push_back(i); b.posY.push_back(j); }
  }
  int k = (int)b.posX.size();
  if (k == 0) {  // outer product: empty block, treat as suffix of X / prefix of Y
    b.x_contig = b.y_contig = b.y_prefix = b.same_order = true;
    b.x_start = X.rank();
    return b;
  }
  // Extent-1 modes never break contiguity, but keeping the test strict is simpler and always safe.
  b.x_contig = true;
  for (int i = 1; i < k; ++i) if (b.posX[i] != b.posX[i - 1] + 1) b.x_contig = false;
  b.x_start = b.posX[0];
  b.same_order = true;
  for (int i = 1; i < k; ++i) if (b.posY[i] != b.posY[i - 1] + 1) b.same_order = false;
  b.y_contig = b.same_order;
  b.y_prefix = b.y_contig && b.posY[0] == 0;
  b.y_suffix = b.y_contig && b.posY[k - 1] == Y.rank() - 1;
  return b;
}

template <typename T>
bool is_direct(const DTensor<T>& X, const DTensor<T>& Y, const BlockInfo& b) {
  if (!(b.x_contig && b.y_contig && b.same_order && (b.y_prefix || b.y_suffix))) return false;
  // strided-batched form with a tiny leading extent is not worth it: fall back to a permute of X
  int k = (int)b.posX.size();
  int64_t P = 1, Q = 1;
  for (int i = 0; i < b.x_start; ++i) P *= X.dims[i];
  for (int i = b.x_start + k; i < X.rank(); ++i) Q *= X.dims[i];
  if (P > 1 && Q > 1 && P < 16 && Q > 4) return false;
  return true;
}

template <typename T>
std::vector<Label> out_labels_for(const DTensor<T>& X, const DTensor<T>& Y, const BlockInfo& b) {
  std::vector<Label> out;
  int k = (int)b.posX.size();
  for (int i = 0; i < b.x_start; ++i) out.push_back(X.labels[i]);
  for (int j = 0; j < Y.rank(); ++j) if (X.find(Y.labels[j]) < 0) out.push_back(Y.labels[j]);
  for (int i = b.x_start + k; i < X.rank(); ++i) out.push_back(X.labels[i]);
  return out;
}

// X has its shared block contiguous at x_start; Y has it as prefix or suffix in the same order.
template <typename T>
DTensor<T> run_direct(Ctx* ctx, const DTensor<T>& X, const DTensor<T>& Y, const BlockInfo& b, bool conjX, bool conjY) {
  int k = (int)b.posX.size();
  int64_t P = 1, Kc = 1, Q = 1, N = 1;
  for (int i = 0; i < b.x_start; ++i) P *= X.dims[i];
  for (int i = b.x_start; i < b.x_start + k; ++i) Kc *= X.dims[i];
  for (int i = b.x_start + k; i < X.rank(); ++i) Q *= X.dims[i];
  std::vector<Label> ol = out_labels_for(X, Y, b);
  std::vector<int64_t> od;
  for (int i = 0; i < b.x_start; ++i) od.push_back(X.dims[i]);
  for (int j = 0; j < Y.rank(); ++j) if (X.find(Y.labels[j]) < 0) { od.push_back(Y.dims[j]); N *= Y.dims[j]; }
  for (int i = b.x_start + k; i < X.rank(); ++i) od.push_back(X.dims[i]);
  DTensor<T> out(ctx, od, ol);
  const bool yprefix = (k == 0) ? true : b.y_prefix;
  const T one = from_complex<T>(1.0, 0.0), zero = zero_<T>();
  if (out.numel() == 0) return out;
  if (Q == 1 || P > 1) {
    // Out_q[P, N] = X_q[P, Kc] * Ymat, batched over q
    int opa = conjX ? OP_CONJ : OP_N;
    int opb = yprefix ? (conjY ? OP_CONJ : OP_N) : (conjY ? OP_C : OP_T);
    int64_t ldb = yprefix ? Kc : N;
    gemm<T>(ctx, opa, opb, P, N, Kc, one, X.data(), P, P * Kc, Y.data(), ldb, 0, zero, out.data(), P, P * N, Q);
  } else {
    // P == 1: Out[N, Q] = Ymat^T * X[Kc, Q]
    int opa = yprefix ? (conjY ? OP_C : OP_T) : (conjY ? OP_CONJ : OP_N);
    int64_t lda = yprefix ? Kc : N;
    int opb = conjX ? OP_CONJ : OP_N;
    gemm<T>(ctx, opa, opb, N, Q, Kc, one, Y.data(), lda, 0, X.data(), Kc, 0, zero, out.data(), N, 0, 1);
  }
  return out;
}

}  // namespace

template <typename T>
bool contract_direct_labels(const DTensor<T>& X, const DTensor<T>& Y, std::vector<Label>* out) {
  BlockInfo b = analyze(X, Y);
  if (!is_direct(X, Y, b)) return false;
  if (out) *out = out_labels_for(X, Y, b);
  return true;
}

template <typename T>
DTensor<T> contract(Ctx* ctx, const DTensor<T>& A, const DTensor<T>& B, bool conjA, bool conjB, int prefer_x) {
  if (prefer_x != 2) {
    BlockInfo b = analyze(A, B);
    if (is_direct(A, B, b)) return run_direct(ctx, A, B, b, conjA, conjB);
  }
  if (prefer_x != 1) {
    BlockInfo b = analyze(B, A);
    if (is_direct(B, A, b)) return run_direct(ctx, B, A, b, conjB, conjA);
  }
  if (prefer_x == 2) {
    BlockInfo b = analyze(A, B);
    if (is_direct(A, B, b)) return run_direct(ctx, A, B, b, conjA, conjB);
  } else if (prefer_x == 1) {
    BlockInfo b = analyze(B, A);
    if (is_direct(B, A, b)) return run_direct(ctx, B, A, b, conjB, conjA);
  }
  // Fallback: X = the larger operand keeps its free-label order; shared labels are moved to its end
  // (only if they are not already one contiguous block), Y is permuted to [shared (X order)..., free...].
  const bool a_is_x = (prefer_x == 1) || (prefer_x == 0 && A.numel() >= B.numel());
  DTensor<T> X = a_is_x ? A : B, Y = a_is_x ? B : A;
  bool cX = a_is_x ? conjA : conjB, cY = a_is_x ? conjB : conjA;
  BlockInfo b = analyze(X, Y);
  int k = (int)b.posX.size();
  int64_t P = 1, Q = 1;
  for (int i = 0; i < b.x_start; ++i) P *= X.dims[i];
  for (int i = b.x_start + k; i < X.rank(); ++i) Q *= X.dims[i];
  bool x_ok = b.x_contig && !(P > 1 && Q > 1 && P < 16 && Q > 4);
  if (!x_ok) {
    std::vector<Label> order;
    for (int i = 0; i < X.rank(); ++i) if (Y.find(X.labels[i]) < 0) order.push_back(X.labels[i]);
    for (int i = 0; i < k; ++i) order.push_back(X.labels[b.posX[i]]);
    X = permuted(ctx, X, order);
  }
  {
    std::vector<Label> order;
    for (int i = 0; i < X.rank(); ++i) if (Y.find(X.labels[i]) >= 0) order.push_back(X.labels[i]);
    for (int j = 0; j < Y.rank(); ++j) if (X.find(Y.labels[j]) < 0) order.push_back(Y.labels[j]);
    Y = permuted(ctx, Y, order);
  }
  b = analyze(X, Y);
  NSB_REQUIRE(b.x_contig && b.y_prefix && b.same_order, NSB_EINTERNAL, "contract: fallback layout failed");
  return run_direct(ctx, X, Y, b, cX, cY);
}

template <typename T>
DTensor<T> apply_small(Ctx* ctx, SmallOp<T>& op, const DTensor<T>& X, const DTensor<T>& W,
                       const std::vector<Label>& out_labels) {
  if (!op.built || op.in_labels != X.labels || op.in_dims != X.dims || op.out_labels != out_labels) {
    op = SmallOp<T>();
    op.in_labels = X.labels; op.in_dims = X.dims; op.out_labels = out_labels;
    std::vector<int64_t> xstride(X.rank());
    { int64_t s = 1; for (int i = 0; i < X.rank(); ++i) { xstride[i] = s; s *= X.dims[i]; } }
    // output dims
    op.out_dims.resize(out_labels.size());
    for (size_t i = 0; i < out_labels.size(); ++i) {
      int ix = X.find(out_labels[i]);
      if (ix >= 0) { NSB_REQUIRE(W.find(out_labels[i]) < 0, NSB_EINTERNAL, "apply_small: kept label also in W"); op.out_dims[i] = X.dims[ix]; }
      else { int iw = W.find(out_labels[i]); NSB_REQUIRE(iw >= 0, NSB_EINTERNAL, "apply_small: unknown output label"); op.out_dims[i] = W.dims[iw]; }
    }
    std::vector<int64_t> ostride(out_labels.size());
    { int64_t s = 1; for (size_t i = 0; i < out_labels.size(); ++i) { ostride[i] = s; s *= op.out_dims[i]; } }
    // contracted labels in X order; new labels in output order
    std::vector<Label> kl, nl;
    std::vector<int64_t> kd, ks, nd, ns;
    for (int i = 0; i < X.rank(); ++i)
      if (W.find(X.labels[i]) >= 0) { kl.push_back(X.labels[i]); kd.push_back(X.dims[i]); ks.push_back(xstride[i]); }
    for (size_t i = 0; i < out_labels.size(); ++i)
      if (X.find(out_labels[i]) < 0) { nl.push_back(out_labels[i]); nd.push_back(op.out_dims[i]); ns.push_back(ostride[i]); }
    NSB_REQUIRE((int)(kl.size() + nl.size()) == W.rank(), NSB_EINTERNAL, "apply_small: operator labels do not match");
    op.K = 1; for (auto d : kd) op.K *= (int)d;
    op.N = 1; for (auto d : nd) op.N *= (int)d;
    // big modes in output order
    op.nbig = 0; op.big_dims.clear(); op.xs.clear(); op.os.clear();
    for (size_t i = 0; i < out_labels.size(); ++i) {
      int ix = X.find(out_labels[i]);
      if (ix < 0) continue;
      if (X.dims[ix] == 1) continue;
      op.big_dims.push_back(X.dims[ix]); op.xs.push_back(xstride[ix]); op.os.push_back(ostride[i]); op.nbig++;
    }
    std::vector<int64_t> koff(op.K), noff(op.N);
    for (int k = 0; k < op.K; ++k) { int64_t r = k, o = 0; for (size_t i = 0; i < kd.size(); ++i) { o += (r % kd[i]) * ks[i]; r /= kd[i]; } koff[k] = o; }
    for (int n = 0; n < op.N; ++n) { int64_t r = n, o = 0; for (size_t i = 0; i < nd.size(); ++i) { o += (r % nd[i]) * ns[i]; r /= nd[i]; } noff[n] = o; }
    op.koff = DevBuf(ctx, sizeof(int64_t) * std::max(op.K, 1));
    op.noff = DevBuf(ctx, sizeof(int64_t) * std::max(op.N, 1));
    NSB_CUDA(cudaMemcpyAsync(op.koff.ptr, koff.data(), sizeof(int64_t) * op.K, cudaMemcpyHostToDevice, ctx->stream));
    NSB_CUDA(cudaMemcpyAsync(op.noff.ptr, noff.data(), sizeof(int64_t) * op.N, cudaMemcpyHostToDevice, ctx->stream));
    NSB_CUDA(cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
    // operator matrix [contracted (X order)..., new (output order)...]
    std::vector<Label> worder = kl;
    worder.insert(worder.end(), nl.begin(), nl.end());
    uint64_t pb = ctx->cnt.permute_bytes;
    DTensor<T> Wp = permuted(ctx, W, worder);
    ctx->cnt.permute_bytes = pb;   // operator tensors are O(w^2 d^2): not a layout permute of state data
    op.Wmat = DevBuf(ctx, sizeof(T) * (size_t)op.K * op.N);
    vec_copy<T>(ctx, (int64_t)op.K * op.N, Wp.data(), (T*)op.Wmat.ptr);
    op.built = true;
  }
  DTensor<T> out(ctx, op.out_dims, op.out_labels);
  small_apply<T>(ctx, X.data(), out.data(), (const T*)op.Wmat.ptr, op.nbig, op.big_dims.data(), op.xs.data(), op.os.data(),
                 op.K, (const int64_t*)op.koff.ptr, op.N, (const int64_t*)op.noff.ptr);
  return out;
}

#define INST(T)                                                                                           \
  template DTensor<T> clone<T>(Ctx*, const DTensor<T>&);                                                  \
  template DTensor<T> permuted<T>(Ctx*, const DTensor<T>&, const std::vector<Label>&, bool);              \
  template DTensor<T> contract<T>(Ctx*, const DTensor<T>&, const DTensor<T>&, bool, bool, int);           \
  template bool contract_direct_labels<T>(const DTensor<T>&, const DTensor<T>&, std::vector<Label>*);     \
  template DTensor<T> apply_small<T>(Ctx*, SmallOp<T>&, const DTensor<T>&, const DTensor<T>&, const std::vector<Label>&);
INST(double)
INST(cdouble)

}  // namespace nsb
