// Labelled dense device tensors and the pairwise contraction engine ("mode products").
//
// contract(A, B) sums over the labels the operands share (ITensor `*` semantics).  The engine never
// permutes a large intermediate when it can be avoided: if the shared labels form one contiguous block
// of operand X (anywhere) and a prefix or suffix of operand Y, in the same order, the contraction is a
// single (possibly strided-batched) GEMM and the result is X with the block replaced by Y's free labels.
// Layout permutes are the fallback and are accounted in Counters::permute_bytes.
#pragma once
#include "common.h"
#include "gemm.h"
#include "ops.h"

namespace nsb {

typedef int32_t Label;
enum LabelKind { LK_SITE = 0, LK_LINK = 1, LK_OP = 2, LK_AUX = 3 };
inline Label make_label(int kind, int id, int plev = 0) { return (Label)((id << 4) | (plev << 2) | kind); }
inline int label_kind(Label l) { return l & 3; }
inline int label_plev(Label l) { return (l >> 2) & 3; }
inline int label_id(Label l) { return l >> 4; }
inline Label label_setplev(Label l, int p) { return (Label)((l & ~(3 << 2)) | (p << 2)); }

template <typename T>
struct DTensor {
  std::shared_ptr<DevBuf> buf;
  int64_t offset = 0;            // element offset into buf (views of a slab along the last mode)
  std::vector<int64_t> dims;
  std::vector<Label> labels;

  DTensor() {}
  DTensor(Ctx* ctx, const std::vector<int64_t>& d, const std::vector<Label>& l) : dims(d), labels(l) {
    NSB_REQUIRE(d.size() == l.size(), NSB_EINTERNAL, "DTensor: rank mismatch");
    buf = std::make_shared<DevBuf>(ctx, sizeof(T) * (size_t)std::max<int64_t>(numel(), 1));
  }
  T* data() const { return buf ? reinterpret_cast<T*>(buf->ptr) + offset : nullptr; }
  int rank() const { return (int)dims.size(); }
  int64_t numel() const { int64_t n = 1; for (auto d : dims) n *= d; return n; }
  int find(Label l) const { for (int i = 0; i < rank(); ++i) if (labels[i] == l) return i; return -1; }
  int64_t dim_of(Label l) const { int i = find(l); NSB_REQUIRE(i >= 0, NSB_EINTERNAL, "label not found"); return dims[i]; }
  bool valid() const { return (bool)buf; }
  // shallow views
  DTensor<T> relabeled(const std::vector<Label>& nl) const { DTensor<T> t = *this; t.labels = nl; return t; }
  DTensor<T> primed(int inc = 1) const {
    DTensor<T> t = *this;
    for (auto& l : t.labels) l = label_setplev(l, label_plev(l) + inc);
    return t;
  }
  // view of indices [lo, hi) of the last (slowest) mode; shares the buffer
  DTensor<T> last_mode_slab(int64_t lo, int64_t hi) const {
    DTensor<T> t = *this;
    int64_t inner = 1;
    for (int i = 0; i + 1 < rank(); ++i) inner *= dims[i];
    t.offset = offset + lo * inner;
    t.dims.back() = hi - lo;
    return t;
  }
  DTensor<T> noprime() const {
    DTensor<T> t = *this;
    for (auto& l : t.labels) l = label_setplev(l, 0);
    return t;
  }
};

template <typename T> DTensor<T> clone(Ctx* ctx, const DTensor<T>& A);
template <typename T> DTensor<T> permuted(Ctx* ctx, const DTensor<T>& A, const std::vector<Label>& order, bool conj = false);

// prefer_x: 0 = try (X=A,Y=B) first then (X=B,Y=A); 1 = force X=A; 2 = force X=B.
template <typename T>
DTensor<T> contract(Ctx* ctx, const DTensor<T>& A, const DTensor<T>& B, bool conjA = false, bool conjB = false,
                    int prefer_x = 0);

// Predict the label order contract() would produce for a given operand role assignment without running it;
// returns false if that assignment is not permutation-free.
template <typename T>
bool contract_direct_labels(const DTensor<T>& X, const DTensor<T>& Y, std::vector<Label>* out);

// A cached small-operator application: out = sum_k W[k..., n...] X[..., k..., ...]
template <typename T>
struct SmallOp {
  bool built = false;
  std::vector<Label> in_labels, out_labels;
  std::vector<int64_t> in_dims, out_dims;
  int nbig = 0, K = 0, N = 0;
  std::vector<int64_t> big_dims, xs, os;
  DevBuf Wmat, koff, noff;
};

// Build (once) and run a small-operator application.  `W` carries the operator (all labels either shared
// with X = contracted, or new); `out_labels` gives the output order (must be X's kept labels plus W's new).
template <typename T>
DTensor<T> apply_small(Ctx* ctx, SmallOp<T>& op, const DTensor<T>& X, const DTensor<T>& W,
                       const std::vector<Label>& out_labels);

}  // namespace nsb
