// Bandwidth-bound tensor kernels: permute, small-operator apply (MPO application, K2), Krylov vector
// algebra (K4/K5/K14), directsum / zero-pad (K11), Philox fill.
#pragma once
#include "common.h"

namespace nsb {

constexpr int MAX_RANK = 12;

// out[perm(i)] = in[i]:  out dim d = in dim perm[d]   (out index order given by perm over input modes)
template <typename T>
void permute(Ctx* ctx, const T* in, T* out, int rank, const int64_t* in_dims, const int* perm, bool conj = false);

// Small-operator apply:  out[big..., n...] = sum_k W[k, n] * X[big..., k...]
//   nbig  big (kept) modes with extents big_dims, element strides in X (xs_big) and in out (os_big)
//   K contracted combos: k_off[k] = element offset in X;  N new combos: n_off[n] = element offset in out
//   W is a K x N matrix (column-major, k fastest) in device memory.
template <typename T>
void small_apply(Ctx* ctx, const T* X, T* out, const T* W, int nbig, const int64_t* big_dims,
                 const int64_t* xs_big, const int64_t* os_big, int K, const int64_t* k_off, int N,
                 const int64_t* n_off);

// ---- vector algebra (device-resident scalars avoided: results come back to the host, one sync) ----
template <typename T> void vec_dot(Ctx* ctx, int64_t n, const T* x, const T* y, double* re_out, double* im_out);  // <x|y>
template <typename T> double vec_nrm2(Ctx* ctx, int64_t n, const T* x);
// asynchronous form: <x|y> lands in device slot `slot` (2 doubles at dot_slot_ptr); a collective may sum the slots over the
// ranks before dot_slots_fetch brings the first nslots back (one stream synchronisation)
template <typename T> void vec_dot_slot(Ctx* ctx, int64_t n, const T* x, const T* y, int slot);
double* dot_slot_ptr(Ctx* ctx, int slot);
void dot_slots_fetch(Ctx* ctx, int nslots, double* out_host /* 2 * nslots */);
template <typename T> void vec_axpy(Ctx* ctx, int64_t n, T a, const T* x, T* y);                  // y += a x
template <typename T> void vec_scale(Ctx* ctx, int64_t n, T a, T* x);                             // x *= a
template <typename T> void vec_copy(Ctx* ctx, int64_t n, const T* x, T* y);
template <typename T> void vec_zero(Ctx* ctx, int64_t n, T* x);
// y = sum_i c[i] * xs[i]   (nvec <= 64 pointers)
template <typename T> void vec_lincomb(Ctx* ctx, int64_t n, int nvec, const T* const* xs, const T* c, T* y);
// fused multi-dot: out[i] = <xs[i] | y>, i < nvec (one pass over y)
template <typename T> void vec_multi_dot(Ctx* ctx, int64_t n, int nvec, const T* const* xs, const T* y, T* out_host);

// out (cols x rows) = conj-transpose of in (rows x cols, ld)
template <typename T> void transpose_conj(Ctx* ctx, const T* in, int64_t rows, int64_t cols, int64_t ld, T* out, int64_t ldo, bool conj);
// copy a (rows x cols) block: out[r + c*ldo] = in[r + c*ldi]
template <typename T> void copy_block(Ctx* ctx, const T* in, int64_t ldi, T* out, int64_t ldo, int64_t rows, int64_t cols);
// gather columns: out[:, j] = scale[j] * in[:, idx[j]]  (scale nullable)
template <typename T> void gather_cols(Ctx* ctx, const T* in, int64_t ld, int64_t rows, const int32_t* idx_dev, int64_t ncols, const double* scale_dev, T* out, int64_t ldo);
// gather rows: out[i, :] = in[idx[i], :]
template <typename T> void gather_rows(Ctx* ctx, const T* in, int64_t ld, const int32_t* idx_dev, int64_t nrows, int64_t ncols, T* out, int64_t ldo);
// concatenate along one mode: out[pre, a+b, post] from A[pre, a, post], B[pre, b, post] (B nullable => zero pad)
// max_{a,a'} |E[a, w, a'] - delta| per channel w of an [n, W, n] tensor, returned on the host (synchronises)
template <typename T> void identity_deviation(Ctx* ctx, const T* E, int64_t n, int64_t W, double* out_host);
// out[pre, m + 1, post] = A[pre, m, post] with B[pre, post] inserted as slice `pos` of the middle mode
template <typename T> void insert_mode(Ctx* ctx, const T* A, const T* B, T* out, int64_t pre, int64_t m, int64_t pos, int64_t post);
template <typename T> void concat_mode(Ctx* ctx, const T* A, const T* B, T* out, int64_t pre, int64_t a, int64_t b, int64_t post);
// Philox-4x32 N(0,1) fill (real and imaginary parts independent), scaled
template <typename T> void fill_normal(Ctx* ctx, T* x, int64_t n, uint64_t seed, double scale);
template <typename T> void set_identity(Ctx* ctx, T* x, int64_t rows, int64_t cols, int64_t ld);
// out[i] = sum_{s < nslabs} in[s * slab + i]   (owner-side reduction of the fused reduce-scatter)
template <typename T> void sum_slabs(Ctx* ctx, const T* in, int nslabs, int64_t slab, T* out);
// column squared norms of a (rows x cols) matrix
template <typename T> void col_norms2(Ctx* ctx, const T* A, int64_t rows, int64_t cols, int64_t ld, double* out_dev);

}  // namespace nsb
