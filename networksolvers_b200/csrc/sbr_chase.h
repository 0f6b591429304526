// Band -> tridiagonal by bulge chasing (stage 2 of a two-stage tridiagonalisation, tools/proto_sbr.py) -- the arithmetic of
// one chasing task on packed band storage, written once for the host (one thread, tests/dc_cpu_harness.cpp) and for a
// device thread team (a CTA: loops strided by the team's thread index, team barriers between the phases).
// No CUDA types here.  Round-2 groundwork: used by the experimental kernel of csrc/sbr.cu (test hook only, not by the
// eigensolver yet) and exercised on the CPU by tests/test_cpu_dc.py::test_bulge_chasing_on_band_storage; real FP64.
//
// Storage: lower band with room for the bulge, AB[(i - j) + j * ld] = A[i, j] for 0 <= i - j <= 2 b, ld >= 2 b + 1
// (n x (2 b + 1) doubles: 8.4 MB at n = 8192, b = 64 -- resident in L2).
//
// Task (j, s), window rows r0 = j + 1 + s b .. r1 - 1 (r1 = min(r0 + b, n)), column c = j (s = 0) or r0 - b (s > 0):
//   G = I - tau v v^T with G A[r0:r1, c] = beta e_0;
//   E = A[r0:r1, c+1:r0]  <- G E        (the rest of the bulge the previous step left; empty for s = 0)
//   D = A[r0:r1, r0:r1]   <- G D G      (symmetric rank-2 update, lower part)
//   F = A[r1:r2, r0:r1]   <- F G        (r2 = min(r1 + b, n): creates the next bulge)
// Sweep j + 1 may run step s once sweep j has finished step s + 2 (wavefront t = 3 j + s).
// The reflectors of sweep j tile rows j + 1 .. n - 1, so they are stored as column j of a lower-triangular n x n matrix V2
// (v[0] = 1 stored explicitly) with tau2[s + j * nsteps_max].
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define NSB_HD __host__ __device__
#else
#define NSB_HD
#endif

namespace nsb {
namespace sbr {

struct Band {
  double* ab; int64_t ld, n; int b;
  // element (i, j), 0 <= i - j <= 2 b.  On the device the band is shared between CTAs that hand tasks to each other through
  // flags, and L1 is not coherent across SMs: loads go to L2 (ld.global.cg), stores are write-through anyway.
  NSB_HD double get(int64_t i, int64_t j) const {
#ifdef __CUDA_ARCH__
    return __ldcg(ab + (i - j) + j * ld);
#else
    return ab[(i - j) + j * ld];
#endif
  }
  NSB_HD void set(int64_t i, int64_t j, double x) const { ab[(i - j) + j * ld] = x; }
};

// One host thread standing in for a team.
struct SerialTeam {
  int tid = 0, size = 1;
  NSB_HD void sync() const {}
  NSB_HD double sum(double x, double*) const { return x; }   // team-wide sum, valid in every thread
};

NSB_HD inline int64_t nsteps(int64_t n, int b, int64_t j) { return n - j - 1 <= 0 ? 0 : (n - j - 1 + b - 1) / b; }

// Runs task (j, s).  v (>= b doubles, team-visible), work (>= 2 b doubles, team-visible) and red (team reduction scratch) are
// scratch; on return v[0 .. len) holds the reflector and *tau_out its scalar.  Returns len (0 or 1: nothing to annihilate).
template <class Team>
NSB_HD int chase_task(const Team& tm, const Band& B, int64_t j, int s, double* v, double* tau_out, double* work, double* red) {
  const int64_t n = B.n;
  const int b = B.b;
  const int64_t r0 = j + 1 + (int64_t)s * b, r1 = (r0 + b < n) ? r0 + b : n, c = (s == 0) ? j : r0 - b;
  const int len = (int)(r1 - r0);
  if (len < 2) { if (tm.tid == 0) *tau_out = 0.0; return len < 0 ? 0 : len; }
  // ---- reflector from x = A[r0:r1, c]
  double part = 0.0;
  for (int i = 1 + tm.tid; i < len; i += tm.size) { const double x = B.get(r0 + i, c); part += x * x; }
  const double sigma = tm.sum(part, red);
  const double alpha = B.get(r0, c);
  double tau = 0.0, beta = alpha, scale = 0.0;
  if (sigma != 0.0) {
    beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
  for (int i = tm.tid; i < len; i += tm.size) v[i] = (i == 0) ? 1.0 : scale * B.get(r0 + i, c);
  tm.sync();
  if (tm.tid == 0) { *tau_out = tau; B.set(r0, c, beta); }
  for (int i = 1 + tm.tid; i < len; i += tm.size) B.set(r0 + i, c, 0.0);
  if (tau == 0.0) return len;
  // ---- E <- G E: columns c + 1 .. r0 - 1
  for (int64_t cc = c + 1 + tm.tid; cc < r0; cc += tm.size) {
    double w = 0.0;
    for (int i = 0; i < len; ++i) w += v[i] * B.get(r0 + i, cc);
    w *= tau;
    for (int i = 0; i < len; ++i) B.set(r0 + i, cc, B.get(r0 + i, cc) - w * v[i]);
  }
  // ---- D <- G D G (lower storage): p = tau D v, w = p - (tau/2) (v^T p) v, D -= v w^T + w v^T
  double* p = work;
  double dotp = 0.0;
  for (int i = tm.tid; i < len; i += tm.size) {
    double acc = 0.0;
    for (int k = 0; k < len; ++k) acc += ((i >= k) ? B.get(r0 + i, r0 + k) : B.get(r0 + k, r0 + i)) * v[k];
    p[i] = tau * acc;
    dotp += v[i] * p[i];
  }
  tm.sync();
  const double vtp = tm.sum(dotp, red);
  const double hv = 0.5 * tau * vtp;
  double* w = work + b;
  for (int i = tm.tid; i < len; i += tm.size) w[i] = p[i] - hv * v[i];
  tm.sync();
  for (int k = tm.tid; k < len; k += tm.size)          // column k of the lower triangle
    for (int i = k; i < len; ++i) B.set(r0 + i, r0 + k, B.get(r0 + i, r0 + k) - (v[i] * w[k] + w[i] * v[k]));
  // ---- F <- F G: rows r1 .. r2 - 1
  const int64_t r2 = (r1 + b < n) ? r1 + b : n;
  for (int64_t i = r1 + tm.tid; i < r2; i += tm.size) {
    double u = 0.0;
    for (int k = 0; k < len; ++k) u += B.get(i, r0 + k) * v[k];
    u *= tau;
    for (int k = 0; k < len; ++k) B.set(i, r0 + k, B.get(i, r0 + k) - u * v[k]);
  }
  tm.sync();
  return len;
}

// The same task with its three blocks staged in a team-local buffer `stage` (>= 3 b b doubles: shared memory on the device):
// the band is read and written once, coalesced along the columns of the packed storage, and all arithmetic runs on the
// staged copies.  (The element-wise version above costs ~170 us per task on a B200 through L2 latency; this is the form the
// device kernel is meant to use.)  Same results as chase_task up to the order of the floating-point sums.
template <class Team>
NSB_HD int chase_task_staged(const Team& tm, const Band& B, int64_t j, int s, double* v, double* tau_out, double* work, double* red,
                             double* stage) {
  const int64_t n = B.n;
  const int b = B.b;
  const int64_t r0 = j + 1 + (int64_t)s * b, r1 = (r0 + b < n) ? r0 + b : n, c = (s == 0) ? j : r0 - b;
  const int len = (int)(r1 - r0);
  if (len < 2) { if (tm.tid == 0) *tau_out = 0.0; return len < 0 ? 0 : len; }
  const int64_t r2 = (r1 + b < n) ? r1 + b : n;
  const int ne = (int)(r0 - c), nf = (int)(r2 - r1);      // columns of E (the first one is x), rows of F
  double* E = stage;                                      // len x ne, ld len
  double* D = stage + (size_t)b * b;                      // len x len, ld len (full symmetric copy)
  double* F = stage + 2 * (size_t)b * b;                  // nf x len, ld nf
  for (int e = tm.tid; e < len * ne; e += tm.size) { const int i = e % len, cc = e / len; E[e] = B.get(r0 + i, c + cc); }
  for (int e = tm.tid; e < len * len; e += tm.size) {
    const int i = e % len, k = e / len;
    if (i >= k) { const double x = B.get(r0 + i, r0 + k); D[i + k * len] = x; D[k + i * len] = x; }
  }
  for (int e = tm.tid; e < nf * len; e += tm.size) { const int i = e % nf, k = e / nf; F[e] = B.get(r1 + i, r0 + k); }
  tm.sync();
  // ---- reflector from x = E[:, 0]
  double part = 0.0;
  for (int i = 1 + tm.tid; i < len; i += tm.size) part += E[i] * E[i];
  const double sigma = tm.sum(part, red);
  const double alpha = E[0];
  double tau = 0.0, beta = alpha, scale = 0.0;
  if (sigma != 0.0) {
    beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
  for (int i = tm.tid; i < len; i += tm.size) v[i] = (i == 0) ? 1.0 : scale * E[i];
  tm.sync();
  if (tm.tid == 0) *tau_out = tau;
  for (int i = tm.tid; i < len; i += tm.size) E[i] = (i == 0) ? beta : 0.0;
  if (tau != 0.0) {
    // ---- E <- G E (columns 1 ..)
    for (int cc = 1 + tm.tid; cc < ne; cc += tm.size) {
      double w = 0.0;
      for (int i = 0; i < len; ++i) w += v[i] * E[i + cc * len];
      w *= tau;
      for (int i = 0; i < len; ++i) E[i + cc * len] -= w * v[i];
    }
    // ---- D <- G D G
    double* p = work;
    double dotp = 0.0;
    for (int i = tm.tid; i < len; i += tm.size) {
      double acc = 0.0;
      for (int k = 0; k < len; ++k) acc += D[i + k * len] * v[k];
      p[i] = tau * acc;
      dotp += v[i] * p[i];
    }
    tm.sync();
    const double vtp = tm.sum(dotp, red);
    const double hv = 0.5 * tau * vtp;
    double* w = work + b;
    for (int i = tm.tid; i < len; i += tm.size) w[i] = p[i] - hv * v[i];
    tm.sync();
    for (int e = tm.tid; e < len * len; e += tm.size) { const int i = e % len, k = e / len; D[e] -= v[i] * w[k] + w[i] * v[k]; }
    // ---- F <- F G
    for (int i = tm.tid; i < nf; i += tm.size) {
      double u = 0.0;
      for (int k = 0; k < len; ++k) u += F[i + k * nf] * v[k];
      u *= tau;
      for (int k = 0; k < len; ++k) F[i + k * nf] -= u * v[k];
    }
  }
  tm.sync();
  // ---- write back (lower part of D only)
  for (int e = tm.tid; e < len * ne; e += tm.size) { const int i = e % len, cc = e / len; B.set(r0 + i, c + cc, E[e]); }
  for (int e = tm.tid; e < len * len; e += tm.size) { const int i = e % len, k = e / len; if (i >= k) B.set(r0 + i, r0 + k, D[e]); }
  for (int e = tm.tid; e < nf * len; e += tm.size) { const int i = e % nf, k = e / nf; B.set(r1 + i, r0 + k, F[e]); }
  tm.sync();
  return len;
}

}  // namespace sbr
}  // namespace nsb
