// Block-sparse tensors for abelian quantum-number conservation (the storage ITensors gives QN tensors,
// `siteinds(...; conserve_qns=true)`, examples/dmrg.jl:10) and the sector-batched contraction engine on top of them.
//
// A mode (leg) is cut into sectors: for links the states are grouped by their charge key (stable, in order of first
// appearance), for the small modes -- site indices and operator links -- every state is its own sector of dimension one.
// A tensor stores only its non-vanishing blocks (one dense column-major block per tuple of sectors) back to back in ONE
// flat buffer, so Krylov vector algebra (dot, axpy, linear combinations) runs on the flat buffer unchanged.
//
//   bcontract     sum over shared labels: every pair of blocks whose sectors agree on the shared modes is one GEMM; all
//                 pairs that feed the same output block become the segments of one problem of the grouped GEMM
//                 (gemm_grouped, gemm.cu): one launch per contraction, each output tile written once.
//   bapply_small  application of a small dense operator (MPO site tensors) to modes whose sectors have dimension one: every
//                 non-zero operator element maps a source block onto a destination block with a scalar factor (grouped
//                 block axpy, one launch).
//   from_dense / to_dense / conform   gather / scatter between dense storage and blocks; the block list is either given
//                 (symmetry-allowed blocks of a local tensor) or detected from the exact zeros of the dense tensor.
// Plans (block pairings, offset tables) are cached on the structures' identities: the three H_eff applications of a region
// step and every environment of the same shape re-use them.
#pragma once
#include <map>
#include <memory>

#include "tensor.h"

namespace nsb {

struct BMode {
  int64_t dim = 0;                         // full extent of the mode
  std::vector<int64_t> key;                // per sector: charge key (link) or state index (small mode)
  std::vector<int64_t> sdim, soff;         // per sector: dimension, offset in the sector-sorted order
  std::vector<int32_t> state_sector, state_pos;   // dense state -> (sector, position inside it)
  bool small = false;                      // every sector has dimension one
  uint64_t id = 0;                         // identity of this sectorisation (modes contract only with equal ids or equal tables)
  int nsec() const { return (int)key.size(); }
};
std::shared_ptr<BMode> make_mode_small(int64_t dim);
std::shared_ptr<BMode> make_mode_from_keys(const std::vector<int64_t>& state_keys);   // one key per dense state

struct BStruct {
  std::vector<std::shared_ptr<BMode>> modes;
  struct Blk { std::vector<int32_t> s; int64_t off = 0, numel = 0; };
  std::vector<Blk> blocks;
  std::map<std::vector<int32_t>, int> index;
  int64_t total = 0;
  uint64_t id = 0;
  // device tables for gather / scatter (built on demand)
  DevBuf d_tables;
  int64_t ncand = 0;
  bool tables_ready = false;
  int rank() const { return (int)modes.size(); }
  std::vector<int64_t> block_dims(const Blk& b) const {
    std::vector<int64_t> d(modes.size());
    for (size_t m = 0; m < modes.size(); ++m) d[m] = modes[m]->sdim[b.s[m]];
    return d;
  }
  void add_block(const std::vector<int32_t>& s);
  void finalize();                          // offsets, total, identity
};

template <typename T>
struct BTensor {
  std::shared_ptr<BStruct> st;
  std::shared_ptr<DevBuf> buf;
  std::vector<Label> labels;
  bool valid() const { return (bool)st; }
  T* data() const { return buf ? reinterpret_cast<T*>(buf->ptr) : nullptr; }
  int rank() const { return (int)labels.size(); }
  int find(Label l) const { for (int i = 0; i < rank(); ++i) if (labels[i] == l) return i; return -1; }
  std::vector<int64_t> dims() const { std::vector<int64_t> d; for (auto& m : st->modes) d.push_back(m->dim); return d; }
  BTensor<T> primed(int inc = 1) const { BTensor<T> t = *this; for (auto& l : t.labels) l = label_setplev(l, label_plev(l) + inc); return t; }
  BTensor<T> noprime() const { BTensor<T> t = *this; for (auto& l : t.labels) l = label_setplev(l, 0); return t; }
  BTensor<T> relabeled(const std::vector<Label>& nl) const { BTensor<T> t = *this; t.labels = nl; return t; }
};

struct BCache;   // per-network plan cache (opaque)
std::shared_ptr<BCache> make_bcache();

// dense -> blocks.  st == nullptr: the block list is detected from the non-zero pattern (a block exists iff it holds a
// non-zero element); otherwise only the blocks of st are gathered (everything else is dropped).
template <typename T>
BTensor<T> from_dense(Ctx* ctx, const DTensor<T>& t, const std::vector<std::shared_ptr<BMode>>& modes, std::shared_ptr<BStruct> st = nullptr);
template <typename T> DTensor<T> to_dense(Ctx* ctx, const BTensor<T>& b);
// copy of x in the block layout `st` (same modes): blocks missing in x become zeros, blocks of x missing in st must vanish
template <typename T> BTensor<T> conform(Ctx* ctx, BCache& cache, const BTensor<T>& x, std::shared_ptr<BStruct> st, const std::vector<Label>& labels);

// prefer_x as in contract(): 1 = A is the operand that keeps its mode order.  Returns an invalid tensor when the contraction is
// not permutation-free (caller falls back to the dense engine).
template <typename T>
BTensor<T> bcontract(Ctx* ctx, BCache& cache, const BTensor<T>& A, const BTensor<T>& B, bool conjA, bool conjB, int prefer_x);

// out[kept modes of X..., new modes of W...] = sum_k W[k..., n...] X[..., k..., ...]; W is a small dense operator on the host
// (column-major, modes wlabels / wdims); every contracted mode of X must be small.
template <typename T>
BTensor<T> bapply_small(Ctx* ctx, BCache& cache, const BTensor<T>& X, const std::vector<T>& Whost, const std::vector<Label>& wlabels,
                        const std::vector<int64_t>& wdims, const std::vector<Label>& out_labels, uint64_t op_id);

}  // namespace nsb
