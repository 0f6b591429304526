/* nsb200.h -- C ABI of libnsb200.so, the B200-native sweep engine behind the NetworkSolvers.jl
 * hook API (extracter / updater / inserter).
 *
 * Every entry point returns an int status (NSB_OK == 0, negative on failure) and never throws or
 * aborts across the boundary; nsb_last_error(ctx) gives the message.  All tensors are dense,
 * column-major (first index fastest -- Julia / ITensor dense storage order), element type f64 or
 * complex f64 (interleaved re,im).  The library owns every device allocation behind opaque
 * handles; host buffers are caller-owned and are copied synchronously.
 *
 * Index ("leg") encoding, used by every leg list below: a leg is a pair of int32 (a, b)
 *     (v, NSB_SITE)      the site index of vertex v          (ket / operator "in" index)
 *     (v, NSB_SITE_OUT)  the primed site index of vertex v   (operator "out" index, MPO only)
 *     (v, n), n >= 0     the link index on the tree edge {v, n}  (state link, or operator link
 *                        when the tensor is an operator tensor)
 *
 * Reference interfaces each entry point replaces (paths relative to the reference repo):
 *   nsb_extract          src/extracter.jl:3-17   (itn.orthogonalize, prod(psi[v]), subspace_expand,
 *                                                 itn.position)
 *   nsb_update_eigsolve  src/eigsolve.jl:14-28 -> src/local_solvers/eigsolve.jl:3-29
 *                                                 (KrylovKit.eigsolve over optimal_map)
 *   nsb_update_exp       src/applyexp.jl:18-48 -> src/local_solvers/{runge_kutta,exponentiate}.jl
 *   nsb_insert           src/inserter.jl:3-33    (it.factorize + truncation, set_ortho_region)
 *   nsb_matvec_*         src/operator_map.jl:3-10 / :15-42 (optimal_map / operator_map)
 *   nsb_network_create / nsb_site_upload / nsb_mpo_upload
 *                        src/eigsolve.jl:69-74, src/applyexp.jl:84-89 (EigsolveProblem /
 *                        ApplyExpProblem construction: permute_indices + itn.ProjTTN)
 *   nsb_maxlinkdim       itn.maxlinkdim in the sweep printers src/eigsolve.jl:39, src/applyexp.jl:56
 *   nsb_range_finder     src/sketched_linear_algebra/range_finder.jl:6-64
 *   nsb_fit_target_upload / nsb_update_fit
 *                        src/fitting.jl:25-49 (FittingProblem extracter / updater: region environment of the
 *                        overlap network <psi| A |x>, overlap n / sqrt(n))
 */
#ifndef NSB200_H
#define NSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define NSB_OK 0
#define NSB_EINVAL (-1)
#define NSB_ENOMEM (-2)
#define NSB_ECUDA (-3)
#define NSB_ENCCL (-4)
#define NSB_ENOTCONV (-5)
#define NSB_EUNSUPPORTED (-6) /* e.g. region length not in {1,2}; mirrors src/inserter.jl:26 */
#define NSB_EINTERNAL (-7)

/* dtypes */
#define NSB_F64 0
#define NSB_C128 1

/* leg codes */
#define NSB_SITE (-1)
#define NSB_SITE_OUT (-2)

/* local solvers for nsb_update_exp (src/applyexp.jl:24 `solver=`) */
#define NSB_SOLVER_RK 0     /* runge_kutta_solver, order in nsb_krylov.rk_order (2 or 4) */
#define NSB_SOLVER_KRYLOV 1 /* exponentiate_solver */

/* subspace expansion back-ends (src/subspace/subspace.jl:8-26) */
#define NSB_EXPAND_NONE 0
#define NSB_EXPAND_DENSITYMATRIX 1
#define NSB_EXPAND_ORTHO 2 /* src/subspace/ortho_subspace.jl:19-77 (random expansion orthogonal to the current basis) */

/* phase timers */
#define NSB_T_GAUGE 0
#define NSB_T_THETA 1
#define NSB_T_EXPAND 2
#define NSB_T_ENV 3
#define NSB_T_MATVEC 4
#define NSB_T_KRYLOV 5
#define NSB_T_FACTORIZE 6
#define NSB_T_OTHER 7
#define NSB_NUM_TIMERS 8

typedef struct nsb_ctx nsb_ctx;
typedef struct nsb_net nsb_net;

typedef struct {
  double cutoff;  /* src/truncation_parameters.jl: default 0.0 */
  int64_t mindim; /* default 1 */
  int64_t maxdim; /* default INT64_MAX */
} nsb_trunc;

typedef struct {
  int32_t algorithm;       /* NSB_EXPAND_* */
  int32_t north_pass;      /* default 1 (src/subspace/densitymatrix.jl:10) */
  double expansion_factor; /* default 1.5 (src/subspace/subspace.jl:5) */
  int64_t max_expand;      /* default INT64_MAX */
} nsb_expand;

typedef struct {
  int32_t krylovdim; /* eigsolve default 3, exponentiate default 30 */
  int32_t maxiter;   /* eigsolve default 1, exponentiate default 100 */
  double tol;        /* eigsolve default 1e-14, exponentiate default 1e-12 */
  int32_t which;     /* 0 = :SR (smallest real), 1 = :LR */
  int32_t eager;     /* eigsolve default 0, exponentiate default 1 */
  int32_t rk_order;  /* runge_kutta_solver order, 2 or 4 (default 4) */
  int32_t reserved;
} nsb_krylov;

typedef struct {
  int32_t expanded;    /* 1 if the subspace expansion enlarged a bond, else 0 (soft-fail == 0) */
  int32_t env_builds;  /* environments (re)built by this call */
  int32_t qr_steps;    /* gauge-move QR steps performed */
  int32_t local_rank;  /* number of legs of the local tensor */
  int64_t local_numel; /* elements of the local tensor */
} nsb_extract_info;

typedef struct {
  int32_t nmatvec;   /* H_eff applications */
  int32_t krylovdim; /* Krylov dimension actually reached (last restart) */
  int32_t converged;
  int32_t reserved;
  double residual; /* eigsolve: |beta * y_K| ; exponentiate: accumulated error estimate */
} nsb_solve_info;

typedef struct {
  int64_t newdim;  /* dimension of the new bond (2-site), or current bond for 1-site */
  double truncerr; /* discarded weight / total weight (NDTensors truncate! rule) */
  int32_t decomp;  /* 0 none (1-site), 1 svd (one-sided Jacobi), 2 eigen (density matrix: the reference's route for cutoff > 1e-12,
                      and the device route at n >= eigh_min_n with cutoff == 0), 3 svd label computed as Gram + eigh with
                      Rayleigh-quotient refinement of the spectrum (0 < cutoff <= 1e-12, n >= eigh_min_n) */
  int32_t jacobi_sweeps;
} nsb_insert_info;

typedef struct {
  uint64_t kernel_launches;
  uint64_t gemm_calls;
  double gemm_flops;
  uint64_t permute_bytes;
  uint64_t matvecs;
  uint64_t env_builds;
  uint64_t qr_calls;
  uint64_t svd_calls;
  uint64_t jacobi_sweeps;
} nsb_counters;

/* ---- context ------------------------------------------------------------------------------ */
int nsb_ctx_create(int device, nsb_ctx** out);
int nsb_ctx_destroy(nsb_ctx* ctx);
const char* nsb_last_error(nsb_ctx* ctx); /* ctx may be NULL: last error of the calling thread */
const char* nsb_version(void);
int nsb_ctx_set_option(nsb_ctx* ctx, const char* key, int64_t value); /* "gemm_impl": 0 auto,1 naive,2 dmma,3 dmma+tma; "eigh_min_n": matrix size from which the truncating factorisation takes the Gram + eigh route (0 = never) */
int nsb_ctx_counters(nsb_ctx* ctx, nsb_counters* out);
int nsb_ctx_counters_reset(nsb_ctx* ctx);
int nsb_ctx_synchronize(nsb_ctx* ctx);
/* device-side stopwatch on the context's stream (CUDA events): tic records, toc records + waits + returns ms */
int nsb_event_tic(nsb_ctx* ctx);
int nsb_event_toc(nsb_ctx* ctx, double* ms_out);
int nsb_timers_enable(nsb_ctx* ctx, int on);
int nsb_timers_get(nsb_ctx* ctx, double* ms_out /* NSB_NUM_TIMERS */);
int nsb_timers_reset(nsb_ctx* ctx);
/* Per-launch timing of the GEMM kernels (CUDA events on the context's stream around every GEMM launch while enabled).
 * nsb_gemm_profile_read synchronises and returns, for launch i < min(count, cap): elapsed ms, real flops issued and
 * (M, N, K, batch).  bench.py derives its roofline from the GEMM launches of the timed region itself with this. */
int nsb_gemm_profile_enable(nsb_ctx* ctx, int32_t on); /* also clears the records */
int nsb_gemm_profile_read(nsb_ctx* ctx, int64_t cap, double* ms_out, double* flops_out, int64_t* mnkb_out /* 4*cap */,
                          int64_t* count_out);
int nsb_profiler(nsb_ctx* ctx, int32_t on); /* cudaProfilerStart / cudaProfilerStop around a region (ncu --profile-from-start off) */
int nsb_mem_info(nsb_ctx* ctx, int64_t* free_bytes, int64_t* total_bytes, int64_t* pool_used_bytes);

/* multi-GPU: one process per GPU; rank 0 creates the id, the host side (torch.distributed, MPI ...)
 * broadcasts the 128 bytes, every rank calls nsb_comm_init. */
int nsb_comm_unique_id(char id_out[128]);
int nsb_comm_init(nsb_ctx* ctx, const char id[128], int rank, int nranks);
int nsb_comm_destroy(nsb_ctx* ctx);
/* Shard every H_eff application of this network across the ranks of the context's communicator: theta is split
 * along its last bond, each rank contracts its slab (1/nranks of the flops) and one NCCL all-reduce sums the
 * partial theta'.  `*active` reports whether the current position is shardable (the last bond of theta must
 * carry the last environment); otherwise the matvec stays replicated.  All ranks must make the same calls. */
int nsb_net_set_shard(nsb_net* net, int32_t enable, int32_t* active);
/* Test hook: theta' = H_eff theta computed on ONE device with the arithmetic of an `nranks`-way partition (every rank's partial
 * result / result slab formed exactly as that rank would, the collective replaced by a local sum / concatenation); mode_out:
 * 1 reduce-scatter position, 2 all-gather position, 3 all-reduce (uneven bond).  Lets a single-GPU box check the N > 1 path. */
int nsb_shard_emulate(nsb_net* net, int32_t nranks, void* host_out, int32_t* mode_out);
/* Fused GEMM + reduce-scatter (ctx option "shard_fused" = 1): the last GEMM of the sharded matvec writes every output
 * tile straight into the owning rank's staging window over NVLink peer memory (P2P stores from the epilogue), the owner
 * sums the nranks partial slabs locally and one all-gather completes theta'.  Every rank creates a window of at least
 * (bytes of the local tensor) and opens the windows of all peers (cudaIpc handles exchanged by the host). */
int nsb_peer_window_create(nsb_ctx* ctx, int64_t bytes, char handle_out[64]);
int nsb_peer_window_open(nsb_ctx* ctx, int32_t peer_rank, const char handle[64]);

/* ---- network (state + operator on a tree) -------------------------------------------------- */
int nsb_network_create(nsb_ctx* ctx, int32_t nverts, const int32_t* edges /* 2*nedges */, int32_t nedges,
                       const int64_t* site_dims /* nverts */, int32_t dtype, nsb_net** out);
int nsb_network_destroy(nsb_net* net);
int nsb_site_upload(nsb_net* net, int32_t v, int32_t rank, const int32_t* legs /* 2*rank */,
                    const int64_t* dims, const void* host);
int nsb_site_info(nsb_net* net, int32_t v, int32_t* rank, int32_t* legs /* cap 2*16 */, int64_t* dims /* cap 16 */);
int nsb_site_download(nsb_net* net, int32_t v, void* host);
int nsb_site_fill_random(nsb_net* net, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims,
                         uint64_t seed, double scale);
int nsb_mpo_upload(nsb_net* net, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims,
                   const void* host);
int nsb_set_ortho_region(nsb_net* net, const int32_t* verts, int32_t n);
int nsb_get_ortho_region(nsb_net* net, int32_t* verts /* cap nverts */, int32_t* n);
int nsb_linkdim(nsb_net* net, int32_t u, int32_t v, int64_t* dim);
int nsb_maxlinkdim(nsb_net* net, int64_t* dim);
int nsb_env_drop_all(nsb_net* net); /* forget cached environments (ProjTTN(H) freshly constructed) */
int nsb_env_count(nsb_net* net, int32_t* n);
/* bytes of environment tensors resident on this GPU, and what they would occupy replicated.  They differ on a multi-GPU
 * communicator with ctx option "shard_envs" (default on): environments not incident to the current region are kept as 1 / G
 * slabs per GPU and re-formed by an all-gather when the sweep returns to them. */
int nsb_env_bytes(nsb_net* net, int64_t* resident, int64_t* replicated);

/* ---- abelian quantum numbers (QN-conserving ITensors: `siteinds(...; conserve_qns=true)`, examples/dmrg.jl:10) ----
 * Tensors stay dense (symmetry-forbidden entries are exact zeros); with QNs enabled the three factorisations on the
 * path (nsb_insert, the expansion's eigen, the gauge QR) are done sector by sector with the merged-spectrum truncation
 * of NDTensors, so no step ever mixes sectors.  Every basis state of the link on edge {u, v} carries the charge
 * (nq int32 components) of the subtree on u's side; the other side is total - charge.  A site tensor must be non-zero
 * only where  charge(site state) + sum over neighbours n of charge(subtree beyond n) == total. */
int nsb_qn_enable(nsb_net* net, int32_t nq /* 1..4 */, const int32_t* total_charge /* nq */);
int nsb_qn_set_site(nsb_net* net, int32_t v, const int32_t* charges /* site_dim x nq, state-major */);
int nsb_qn_set_link(nsb_net* net, int32_t u, int32_t v, const int32_t* charges /* linkdim x nq: subtree on u's side */);
int nsb_qn_get_link(nsb_net* net, int32_t u, int32_t v, int32_t* charges_out /* linkdim x nq: subtree on u's side */);
/* zero every entry of the site tensor of v that charge conservation forbids (after nsb_site_fill_random + nsb_qn_set_link:
 * a random symmetric tensor; synthetic QN states of the benchmarks) */
int nsb_qn_project(nsb_net* net, int32_t v);

/* ---- the three hooks ---------------------------------------------------------------------- */
int nsb_extract(nsb_net* net, const int32_t* region, int32_t nreg, const nsb_trunc* trunc /* extracter's */,
                const nsb_expand* expand /* NULL == none */, nsb_extract_info* info /* nullable */);
int nsb_update_eigsolve(nsb_net* net, const nsb_krylov* params, double* eigval, nsb_solve_info* info);
int nsb_update_exp(nsb_net* net, double t_re, double t_im, int32_t solver, const nsb_krylov* params,
                   int32_t nsites, int32_t next_vertex /* 1-site TDVP: next hop, -1 if none */,
                   nsb_solve_info* info);
int nsb_insert(nsb_net* net, const nsb_trunc* trunc /* inserter's */, int32_t normalize, int32_t set_ortho,
               nsb_insert_info* info);

/* ---- single-process multi-device entry points -------------------------------------------------
 * For a host with ONE thread of control (the Julia sweep driver of src/sweep_solve.jl:12-40): an nsb_multi owns one context
 * and one replica network per device and the NCCL communicator joining them; every nsb_multi_* hook runs the corresponding
 * single-device call on all devices concurrently (one host thread per device inside the library) and returns when all are
 * done.  The replicas execute the sharded region step of nsb_net_set_shard in lock step; scalars returned are device 0's
 * (identical on all).  nsb_multi_ctx / nsb_multi_net expose the per-device handles for the non-collective calls
 * (options, counters, timers, nsb_site_download, nsb_linkdim, nsb_maxlinkdim, nsb_norm ... on device 0).
 * nsb_local_sync is the collective that completes a sharded local tensor before nsb_local_download. */
typedef struct nsb_multi nsb_multi;
int nsb_multi_create(const int32_t* devices, int32_t ndev /* 1..8 */, nsb_multi** out);
int nsb_multi_destroy(nsb_multi* m);
const char* nsb_multi_last_error(nsb_multi* m);
int nsb_multi_ndev(nsb_multi* m, int32_t* ndev);
int nsb_multi_ctx(nsb_multi* m, int32_t r, nsb_ctx** out);
int nsb_multi_net(nsb_multi* m, int32_t r, nsb_net** out);
int nsb_multi_network_create(nsb_multi* m, int32_t nverts, const int32_t* edges, int32_t nedges, const int64_t* site_dims, int32_t dtype);
int nsb_multi_site_upload(nsb_multi* m, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, const void* host);
int nsb_multi_mpo_upload(nsb_multi* m, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, const void* host);
int nsb_multi_site_fill_random(nsb_multi* m, int32_t v, int32_t rank, const int32_t* legs, const int64_t* dims, uint64_t seed, double scale);
int nsb_multi_set_ortho_region(nsb_multi* m, const int32_t* verts, int32_t n);
int nsb_multi_set_shard(nsb_multi* m, int32_t enable, int32_t* active);
int nsb_multi_extract(nsb_multi* m, const int32_t* region, int32_t nreg, const nsb_trunc* trunc, const nsb_expand* expand, nsb_extract_info* info);
int nsb_multi_update_eigsolve(nsb_multi* m, const nsb_krylov* params, double* eigval, nsb_solve_info* info);
int nsb_multi_update_exp(nsb_multi* m, double t_re, double t_im, int32_t solver, const nsb_krylov* params, int32_t nsites, int32_t next_vertex,
                         nsb_solve_info* info);
int nsb_multi_insert(nsb_multi* m, const nsb_trunc* trunc, int32_t normalize, int32_t set_ortho, nsb_insert_info* info);
int nsb_multi_matvec_device(nsb_multi* m, int32_t reps);
int nsb_multi_local_download(nsb_multi* m, void* host);
int nsb_multi_synchronize(nsb_multi* m);
int nsb_local_sync(nsb_net* net);

/* ---- fitting (src/fitting.jl) ----------------------------------------------------------------
 * Uploading a target tensor for every vertex switches the network to fitting mode: the ket layer of the
 * environments is the fixed target |x> (own link dimensions), the operator layer the uploaded operator (an identity
 * network with links of dimension 1 for itn.truncate), the bra layer the state being fitted.  nsb_extract then leaves
 * the region environment -- the optimal new region tensor -- as the local tensor (src/fitting.jl:25-40),
 * nsb_update_fit returns the overlap n / sqrt(n) (:42-49), nsb_insert writes it back (normalize = 1,
 * set_ortho = 0 as in src/fitting.jl:78). */
int nsb_fit_target_upload(nsb_net* net, int32_t v, int32_t rank, const int32_t* legs /* 2*rank */,
                          const int64_t* dims, const void* host);
int nsb_update_fit(nsb_net* net, double* overlap);

/* ---- pieces exposed for tests and benchmarks ---------------------------------------------- */
int nsb_local_info(nsb_net* net, int32_t* rank, int32_t* legs /* cap 2*16 */, int64_t* dims /* cap 16 */);
int nsb_local_download(nsb_net* net, void* host);
int nsb_local_upload(nsb_net* net, const void* host);
/* theta' = H_eff theta through the reference-facing call with HOST buffers (H2D + matvec + D2H). */
int nsb_matvec_host(nsb_net* net, const void* host_in, void* host_out);
/* The same with the vector DISTRIBUTED over the ranks of the communicator, as the sharded Krylov solvers hold it
 * (theta sharded along its last bond): every rank passes only its slab [lo, hi) of the last mode (contiguous in
 * the column-major local tensor) and receives the matching slab of theta'.  nsb_shard_range reports the slab; when the
 * position is not slab-sharded (single GPU, uneven bond) the range is the whole mode and the call equals nsb_matvec_host. */
int nsb_shard_range(nsb_net* net, int64_t* lo, int64_t* hi, int64_t* last_dim);
int nsb_matvec_host_slab(nsb_net* net, const void* host_in_slab, void* host_out_slab);
/* theta' = H_eff theta, `reps` times, device resident (output kept internally; download optional). */
int nsb_matvec_device(nsb_net* net, int32_t reps, void* host_out /* nullable */);
/* analytic flop count (real flops) of one H_eff application at the current position */
int nsb_matvec_flops(nsb_net* net, double* flops);
/* flops actually issued per application: the dense count above minus the identity channel of the first / last environment
 * when it is skipped (ctx option "skip_identity", default on; the channel with E[:, w, :] = 1 of an MPO-like operator
 * between orthonormal bases contributes theta itself, so no GEMM work is spent on it).  Throughput is reported from this. */
int nsb_matvec_flops_executed(nsb_net* net, double* flops);
/* norm of the state = norm of the orthogonality-centre tensor (requires a single-vertex centre) */
int nsb_norm(nsb_net* net, double* out);

/* ---- dense helpers (test hooks for the kernels under the hooks) ---------------------------- */
/* C(m x n) = op(A) op(B); opX: 0 = N, 1 = T, 2 = C (conj-transpose), 3 = conj (no transpose). */
int nsb_gemm_host(nsb_ctx* ctx, int32_t dtype, int32_t opa, int32_t opb, int64_t m, int64_t n, int64_t k,
                  const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int32_t impl);
/* time `reps` device-resident GEMMs of that shape (random data); returns avg ms per GEMM */
int nsb_gemm_bench(nsb_ctx* ctx, int32_t dtype, int32_t opa, int32_t opb, int64_t m, int64_t n, int64_t k,
                   int32_t impl, int32_t reps, double* ms_out);
/* FP64 tensor-pipe (DMMA) issue ceiling of this device, measured live (register-resident mma.sync loop);
 * the roofline denominator for the GEMM-shaped kernels (MEASURED_PEAKS.json has no FP64 entry). */
int nsb_dmma_peak(nsb_ctx* ctx, double* tflops_out);
/* truncated factorisation of a host matrix (rows x cols): U (rows x newdim), C = U^H M (newdim x cols),
 * spectrum (sigma^2, descending, length min(rows, cols)) -- src/inserter.jl:23 semantics. */
int nsb_factorize_host(nsb_ctx* ctx, int32_t dtype, int64_t rows, int64_t cols, const void* M,
                       const nsb_trunc* trunc, void* U /* rows*min */, void* C /* min*cols */,
                       double* spectrum, nsb_insert_info* info);
/* Hermitian eigen-decomposition of a host matrix A (n x n, full storage): w ascending, U (n x n, nullable) with
 * A U = U diag(w).  The device eigensolver behind the large-matrix `eigen` route of nsb_insert / nsb_factorize_host
 * (ITensors.factorize which_decomp = "eigen") and of the expansion's eigen(rho)
 * (src/subspace/densitymatrix.jl:53). */
int nsb_eigh_host(nsb_ctx* ctx, int32_t dtype, int64_t n, const void* A, double* w, void* U);
/* thin QR of a host matrix (rows x cols): Q (rows x k), R (k x cols), k = min(rows, cols). */
int nsb_qr_host(nsb_ctx* ctx, int32_t dtype, int64_t rows, int64_t cols, const void* M, void* Q, void* R);
/* time `reps` device-resident thin QRs of a random rows x cols matrix (blocked compact-WY Householder); avg ms per QR */
int nsb_qr_bench(nsb_ctx* ctx, int32_t dtype, int64_t rows, int64_t cols, int32_t reps, double* ms_out);
/* Randomised range finder (src/sketched_linear_algebra/range_finder.jl:6-64), blocked: probes go through the linear map in
 * panels, projection against the accepted vectors by GEMMs, the reference's per-vector acceptance rule (north_pass
 * Gram-Schmidt passes, stop at the first residual below orthogonal_threshold; experimental `cutoff`: keep that vector, then
 * stop) evaluated on the device with one host look per panel.  probes: caller-supplied domain vectors, column k = the k-th
 * random_vector() (domain_size x min(max_rank + oversample, range, domain)), or NULL for Philox N(0,1) with `seed`.
 * Q receives the orthonormal basis (range_size x rank).
 *   nsb_range_finder_host: linear map = a host matrix A (m x n), uploaded once.
 *   nsb_range_finder_heff: linear map = the projected operator at the current position (the closure
 *                          psi -> optimal_map(P, psi) of src/eigsolve.jl:22); domain = range = the local tensor. */
int nsb_range_finder_host(nsb_ctx* ctx, int32_t dtype, int64_t m, int64_t n, const void* A, const void* probes /* nullable */,
                          int64_t max_rank, int32_t oversample, int32_t north_pass, double orthogonal_threshold, double cutoff,
                          uint64_t seed, void* Q /* m*(max_rank+oversample) */, int64_t* rank_out);
int nsb_range_finder_heff(nsb_net* net, const void* probes /* nullable */, uint64_t seed, int64_t max_rank, int32_t oversample,
                          int32_t north_pass, double orthogonal_threshold, double cutoff, void* Q, int64_t* rank_out);
/* One-shot random tensor for the next "ortho" subspace expansion (rows = basis size of the previous vertex, cols =
 * expand_space(basis size), src/subspace/ortho_subspace.jl:4,56): replaces the device Philox draw, so that a host can
 * reproduce random_itensor(basis_inds, ax) of its own generator. */
int nsb_expand_set_probe(nsb_net* net, int64_t rows, int64_t cols, const void* host);

/* EXPERIMENTAL (round-2 groundwork, not used by the hooks above; eigenvalues checked on a B200): band -> tridiagonal by bulge
 * chasing, stage 2 of a two-stage tridiagonalisation (csrc/sbr.cu, csrc/sbr_chase.h, tools/proto_sbr.py).  ab: lower band
 * storage ld x n column-major, ab[(i - j) + j ld] = A[i, j] for 0 <= i - j <= b, rows b + 1 .. 2 b zero (room for the bulge),
 * ld >= 2 b + 1; on return row 0 holds the diagonal and row 1 the sub-diagonal of the tridiagonal matrix.  V2 (n x n) receives
 * the reflectors of sweep j in column j (rows j + 1 ..), tau2 (ldtau x n) their scalars. */
int nsb_sbr_chase_host(nsb_ctx* ctx, int64_t n, int32_t b, double* ab, int64_t ld, double* V2, double* tau2, int64_t ldtau);

#ifdef __cplusplus
}
#endif
#endif /* NSB200_H */
