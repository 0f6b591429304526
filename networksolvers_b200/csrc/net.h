// Device-resident tree tensor network state + operator + projected-operator environments, and the three
// region hooks (extract / update / insert) that the NetworkSolvers.jl sweep driver calls.
#pragma once
#include <complex>
#include <functional>
#include <map>
#include <set>

#include "bsparse.h"
#include "linalg.h"
#include "tensor.h"

namespace nsb {


struct NetBase {
  Ctx* ctx = nullptr;
  int dtype = NSB_F64;
  virtual ~NetBase() {}
  virtual void site_upload(int v, int rank, const int32_t* legs, const int64_t* dims, const void* host) = 0;
  virtual void site_info(int v, int32_t* rank, int32_t* legs, int64_t* dims) = 0;
  virtual void site_download(int v, void* host) = 0;
  virtual void site_fill_random(int v, int rank, const int32_t* legs, const int64_t* dims, uint64_t seed, double scale) = 0;
  virtual void mpo_upload(int v, int rank, const int32_t* legs, const int64_t* dims, const void* host) = 0;
  virtual void set_ortho_region(const int32_t* verts, int n) = 0;
  virtual void get_ortho_region(int32_t* verts, int32_t* n) = 0;
  virtual int64_t linkdim(int u, int v) = 0;
  virtual int64_t maxlinkdim() = 0;
  virtual void env_drop_all() = 0;
  virtual int env_count() = 0;
  virtual void extract(const int32_t* region, int nreg, const nsb_trunc* trunc, const nsb_expand* expand, nsb_extract_info* info) = 0;
  virtual void update_eigsolve(const nsb_krylov* params, double* eigval, nsb_solve_info* info) = 0;
  virtual void update_exp(double tre, double tim, int solver, const nsb_krylov* params, int nsites, int next_vertex, nsb_solve_info* info) = 0;
  virtual void insert(const nsb_trunc* trunc, int normalize, int set_ortho, nsb_insert_info* info) = 0;
  virtual void fit_target_upload(int v, int rank, const int32_t* legs, const int64_t* dims, const void* host) = 0;
  virtual double update_fit() = 0;
  virtual void local_info(int32_t* rank, int32_t* legs, int64_t* dims) = 0;
  virtual void local_download(void* host) = 0;
  virtual void local_sync() = 0;   // multi-GPU: complete a sharded local tensor on every rank (collective)
  virtual void local_upload(const void* host) = 0;
  virtual void matvec_host(const void* in, void* out) = 0;
  virtual void matvec_host_slab(const void* in, void* out) = 0;
  virtual void shard_range(int64_t* lo, int64_t* hi, int64_t* last_dim) = 0;
  virtual void env_bytes(int64_t* resident, int64_t* replicated) = 0;
  virtual void matvec_device(int reps, void* host_out) = 0;
  virtual double matvec_flops() = 0;
  virtual double matvec_flops_executed() = 0;
  virtual double norm() = 0;
  virtual int64_t range_finder_heff(const void* probes_host, uint64_t seed, int64_t max_rank, int oversample, int north_pass, double thr,
                                    double cutoff, void* Qhost) = 0;
  virtual void expand_set_probe(int64_t rows, int64_t cols, const void* host) = 0;
  virtual int set_shard(int enable) = 0;   // returns 1 if the current position is sharded across ranks
  virtual void shard_emulate(int G, void* host_out, int32_t* mode_out) = 0;
  // abelian quantum numbers (dense storage, block-wise factorisations)
  virtual void qn_enable(int nq, const int32_t* total) = 0;
  virtual void qn_set_site(int v, const int32_t* charges) = 0;
  virtual void qn_set_link(int u, int v, const int32_t* charges) = 0;
  virtual void qn_get_link(int u, int v, int32_t* charges_out) = 0;
  virtual void qn_project(int v) = 0;
};

template <typename T>
struct Net : public NetBase {
  int nverts = 0;
  std::vector<std::pair<int, int>> edges;
  std::vector<std::vector<int>> adj;            // neighbours in edge-insertion order
  std::map<std::pair<int, int>, int> eid;       // both orientations -> edge id
  std::vector<int64_t> site_dims;
  std::vector<DTensor<T>> psi, W;
  std::vector<uint64_t> ver;                    // bumped whenever psi[v] changes
  std::vector<int> ortho;                       // orthogonality region
  // projected operator
  std::vector<int> pos;                         // current region (vertices); empty before the first extract
  bool pos_on_edge = false;
  // ident: operator-link channel w with t[:, w, :] = identity to 1e-10 (the "nothing to the left / right yet" channel of an
  // MPO-like operator between orthonormal bases), -1 if none
  // bt: block-sparse form (QN networks with ctx option qn_block_sparse): then t carries dims / labels only until a dense
  // consumer asks for it (env_dense)
  // shard: multi-GPU, ctx option shard_envs -- this rank's slab of t's last mode.  Environments that are not incident to the
  // current region live only as such slabs (1 / G of the HBM on every GPU); an all-gather re-forms t when the sweep comes back.
  struct Env { DTensor<T> t; std::vector<std::pair<int, uint64_t>> deps; int ident = -1; BTensor<T> bt; DTensor<T> shard; };
  std::map<std::pair<int, int>, Env> envs;      // key (u, v): everything on u's side, pointing into v
  // local problem
  DTensor<T> theta;
  std::vector<int> region;
  // type 0: environment (u -> v); 1: site operator at v; 2: the two site operators of a 2-site region merged into one
  // small operator (u = first site, v = second site, Wm = W[u] * W[v] over their shared operator link)
  struct Step { int type; int u, v; SmallOp<T> op; DTensor<T> Wm; std::vector<T> Wm_host; };
  std::vector<Step> plan;
  DTensor<T> last_out;                          // result of the last nsb_matvec_device
  // identity-channel skipping (g_skip_identity): first environment of the plan without its identity channel
  int first_ident = -1;
  DTensor<T> first_compact;
  void prepare_identity_skip();
  bool skip_first_identity(DTensor<T>& X, int64_t bra_lo = 0, int64_t bra_hi = 0, const DTensor<T>* Xs = nullptr);
  DTensor<T> run_plan_steps(DTensor<T> X, size_t i0, size_t i1);
  DTensor<T> last_env_contract(const DTensor<T>& X, double* skipped);
  double skipped_flops(const DTensor<T>& x) const;   // real flops per application the skipping saves (dry run)
  double skipped_last_apply = -1.0;                  // what the last apply_heff actually skipped (< 0: none ran yet)
  // multi-GPU: theta sharded along its last bond across the ranks of ctx->nccl_comm
  bool shard_enabled = false, shard_active = false;
  int shard_mode = 0;                           // 0 none, 1 RS (reduce-scatter), 2 AG (all-gather), 3 AR (all-reduce of full vectors)
  int64_t shard_lo = 0, shard_hi = 0;
  DTensor<T> shard_env;                         // rows [shard_lo, shard_hi) of the last environment (RS / AR positions)
  bool theta_is_slab = false;                   // the local tensor currently lives as this rank's slab (after a sharded update)
  DTensor<T> theta_slab;
  void shard_prepare();
  bool env_sharding_on() const;
  bool env_is_hot(int u, int v) const;          // incident to the current region (or to the position being entered)
  void env_promote(Env& e);                     // slab -> full tensor (all-gather); no-op when t is resident
  void env_demote(Env& e);                      // keep only the slab; no-op when the environment cannot be split evenly
  void env_rebalance();                         // promote the hot environments, demote all others
  std::vector<int> hot_region;                  // region whose incident environments must stay resident
  Env& env_at(int u, int v) { Env& e = envs.at({u, v}); env_promote(e); return e; }
  void ensure_theta_full();
  bool krylov_sharded() const { return shard_active && (shard_mode == 1 || shard_mode == 2); }
  DTensor<T> apply_heff_slab(const DTensor<T>& xs);
  DTensor<T> heff_partial_from_slab(const DTensor<T>& xs, double* skipped);
  DTensor<T> heff_slab_from_full(const DTensor<T>& xf, const DTensor<T>& xs, double* skipped);
  void shard_emulate(int G, void* host_out, int32_t* mode_out) override;
  bool slab_of(const DTensor<T>& t, Label l, int64_t lo, int64_t hi, DTensor<T>* out);
  void nccl_check(int r, const char* what);
  void comm_allreduce(T* buf, int64_t n);
  void comm_allgather(const T* send, T* recv, int64_t n_per_rank);
  void comm_reduce_scatter(const T* send, T* recv, int64_t n_per_rank);
  // Krylov vector algebra on full vectors or on slabs (partial sums + all-reduce of the scalars)
  DTensor<T> kvec_start();
  bool krylov_blocks() const { return qn_bs() && (bool)theta_st; }
  DTensor<T> kapply(const DTensor<T>& v);
  void kdot(const DTensor<T>& a, const DTensor<T>& b, double* re_out, double* im_out);
  double knrm2(const DTensor<T>& a);
  void kstore_theta(const DTensor<T>& x);
  const char* parallelism_note() const;
  // QN bookkeeping: every basis state of a link carries the charge of the subtree on the side of qn_side[e]
  bool qn_on = false;
  int nq = 0;
  std::vector<int64_t> qn_total;
  std::vector<std::vector<int64_t>> qn_site, qn_link;   // [v][state*nq + c], [edge][state*nq + c]
  std::vector<int> qn_side;                             // per edge: vertex whose side the charges describe
  std::vector<int64_t> side_charge(int v, int n) const;                       // charges of the subtree on n's side of {v, n}
  std::vector<int64_t> leg_charges(int owner, Label l) const;
  std::vector<int64_t> multi_keys(int owner, const std::vector<Label>& labels, const std::vector<int64_t>& dims, bool complement) const;
  void qn_store_link(int v, int n, const std::vector<int64_t>& keys_on_v_side);
  // ---- block-sparse engine for QN networks (bsparse.h): environments, local tensor and Krylov vectors as symmetry blocks ----
  bool qn_bs() const { return qn_on && ctx->opt.qn_block_sparse != 0 && !fit_mode; }
  std::shared_ptr<BCache> bcache;
  std::vector<std::shared_ptr<BMode>> link_mode, site_mode, op_mode;    // per edge / vertex / edge, built on demand
  std::shared_ptr<BMode> mode_for(Label l, int64_t dim);
  std::vector<std::shared_ptr<BMode>> modes_for(const std::vector<Label>& labels, const std::vector<int64_t>& dims);
  BTensor<T> bt_of(const DTensor<T>& t, std::shared_ptr<BStruct> st = nullptr);
  std::shared_ptr<BStruct> allowed_struct(const DTensor<T>& th);         // symmetry-allowed blocks of a local tensor
  std::shared_ptr<BStruct> allowed_struct_in(const DTensor<T>& th, const std::vector<int>& inside);
  void qn_project(int v) override;
  std::shared_ptr<BStruct> theta_st;                                     // block layout of the local tensor / Krylov vectors
  bool bt_apply_ok = true;                                               // false: this position falls back to the dense engine
  double bt_last_apply_flops = -1.0;
  std::vector<std::vector<T>> Whost;                                     // host copies of the site operators (bapply_small)
  const std::vector<T>& w_host(int v);
  bool make_env_bt(int u, int v, const std::vector<int>& others, Env* out);
  bool apply_heff_bt(const BTensor<T>& x, BTensor<T>* y);
  const DTensor<T>& env_dense(int u, int v);
  DTensor<T> fake_dense(const BTensor<T>& b) const { DTensor<T> f; f.labels = b.labels; f.dims = b.dims(); return f; }
  FactorInfo factorize_qn(const T* M, int64_t rows, int64_t cols, const std::vector<int64_t>& rk, const std::vector<int64_t>& ck,
                          double cutoff, int64_t mindim, int64_t maxdim, bool sqrt_spectrum, DevBuf& U, DevBuf& C,
                          std::vector<int64_t>& new_keys);
  void qr_qn(const T* M, int64_t rows, int64_t cols, const std::vector<int64_t>& rk, const std::vector<int64_t>& ck, DevBuf& Q,
             DevBuf& R, int64_t* kout, std::vector<int64_t>& new_keys);

  Net(Ctx* c, int nv, const int32_t* e, int ne, const int64_t* sd);

  // fitting mode (src/fitting.jl): the ket layer of the environments is the fixed target network x, whose links
  // carry prime level 2 (0 = ket links of psi, 1 = bra links of psi)
  bool fit_mode = false;
  std::vector<DTensor<T>> xket;
  Label lxlink(int u, int v) const { return make_label(LK_LINK, eid.at({u, v}), 2); }
  DTensor<T> fit_local();

  // labels
  Label lsite(int v, int p = 0) const { return make_label(LK_SITE, v, p); }
  Label llink(int u, int v, int p = 0) const { return make_label(LK_LINK, eid.at({u, v}), p); }
  Label lop(int u, int v) const { return make_label(LK_OP, eid.at({u, v}), 0); }
  std::vector<Label> canonical_labels(int v) const;
  std::vector<Label> decode_legs(int rank, const int32_t* legs, bool is_operator) const;
  void encode_legs(const std::vector<Label>& labels, int32_t* legs) const;
  void set_site(int v, const DTensor<T>& t) { psi[v] = t; ver[v]++; }
  void canonicalize(int v);

  // graph helpers
  std::vector<int> path(int a, int b) const;
  void subtree(int u, int v, std::vector<int>& out) const;   // vertices on u's side of edge (u, v)

  // hot path pieces
  int orthogonalize(const std::vector<int>& target);
  void qr_step(int a, int b);
  DTensor<T> build_theta(const std::vector<int>& reg);
  int position(const std::vector<int>& reg);
  int make_env(int u, int v);
  std::vector<Label> w_out_labels(const DTensor<T>& X, const DTensor<T>& Wv, int v, const std::vector<int>& reg) const;
  std::vector<Label> merged_out_labels(const DTensor<T>& X, int a, int b) const;
  void build_plan();
  DTensor<T> apply_heff(const DTensor<T>& x);
  bool expand_densitymatrix(const nsb_trunc& trunc, const nsb_expand& ex);
  bool expand_ortho(const nsb_trunc& trunc, const nsb_expand& ex);
  uint64_t expand_seed = 0x5eed0001ull;       // Philox stream of the random "ortho" expansion (advanced per call)
  DTensor<T> exp_solve(const std::function<DTensor<T>(const DTensor<T>&)>& H, std::complex<double> t, const DTensor<T>& x0,
                       int solver, const nsb_krylov* kp, int* nmv, int* lastK, int* conv, double* err, bool edge_local);

  // NetBase
  void site_upload(int v, int rank, const int32_t* legs, const int64_t* dims, const void* host) override;
  void site_info(int v, int32_t* rank, int32_t* legs, int64_t* dims) override;
  void site_download(int v, void* host) override;
  void site_fill_random(int v, int rank, const int32_t* legs, const int64_t* dims, uint64_t seed, double scale) override;
  void mpo_upload(int v, int rank, const int32_t* legs, const int64_t* dims, const void* host) override;
  void set_ortho_region(const int32_t* verts, int n) override;
  void get_ortho_region(int32_t* verts, int32_t* n) override;
  int64_t linkdim(int u, int v) override;
  int64_t maxlinkdim() override;
  void env_drop_all() override { envs.clear(); pos.clear(); plan.clear(); }
  int env_count() override { return (int)envs.size(); }
  void extract(const int32_t* region, int nreg, const nsb_trunc* trunc, const nsb_expand* expand, nsb_extract_info* info) override;
  void update_eigsolve(const nsb_krylov* params, double* eigval, nsb_solve_info* info) override;
  void update_exp(double tre, double tim, int solver, const nsb_krylov* params, int nsites, int next_vertex, nsb_solve_info* info) override;
  void insert(const nsb_trunc* trunc, int normalize, int set_ortho, nsb_insert_info* info) override;
  void fit_target_upload(int v, int rank, const int32_t* legs, const int64_t* dims, const void* host) override;
  double update_fit() override;
  void local_info(int32_t* rank, int32_t* legs, int64_t* dims) override;
  void local_download(void* host) override;
  void local_sync() override { ensure_theta_full(); }
  void local_upload(const void* host) override;
  void matvec_host(const void* in, void* out) override;
  void matvec_host_slab(const void* in, void* out) override;
  void shard_range(int64_t* lo, int64_t* hi, int64_t* last_dim) override;
  void env_bytes(int64_t* resident, int64_t* replicated) override;
  void matvec_device(int reps, void* host_out) override;
  double matvec_flops() override;
  double matvec_flops_executed() override;
  double norm() override;
  int64_t range_finder_heff(const void* probes_host, uint64_t seed, int64_t max_rank, int oversample, int north_pass, double thr, double cutoff,
                            void* Qhost) override;
  void expand_set_probe(int64_t rows, int64_t cols, const void* host) override;
  DevBuf expand_probe;                       // one-shot caller-supplied random tensor of the "ortho" expansion
  int64_t expand_probe_rows = 0, expand_probe_cols = 0;
  int set_shard(int enable) override;
  void qn_enable(int nq, const int32_t* total) override;
  void qn_set_site(int v, const int32_t* charges) override;
  void qn_set_link(int u, int v, const int32_t* charges) override;
  void qn_get_link(int u, int v, int32_t* charges_out) override;
};

// small dense host helpers (Ritz problems of the Krylov solvers)
void host_sym_eig(int n, std::vector<double>& A /* n*n col-major, destroyed */, std::vector<double>& evals,
                  std::vector<double>& evecs /* n*n col-major */);
void host_expm_complex(int n, std::vector<std::complex<double>>& A /* n*n col-major, in: A, out: exp(A) */);

}  // namespace nsb
