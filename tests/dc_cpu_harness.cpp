// CPU harness for the divide & conquer logic shared with the device path (networksolvers_b200/csrc/dc_secular.h):
// leaf partition, deflation planning and the secular root finder are the very functions libnsb200.so runs (the
// root finder inside a kernel); the merge loop below mirrors dc_solve() of csrc/eigh.cu with plain loops in place
// of the kernels and GEMMs.  Test infrastructure only (built and driven by tests/test_cpu_dc.py).
#include "../networksolvers_b200/csrc/dc_secular.h"

#include <cstring>

using namespace nsb::dc;

extern "C" {

int64_t dc_leaf_bounds(int64_t n, int64_t leaf, int64_t* out, int64_t cap) {
  std::vector<int64_t> b = leaf_bounds(n, leaf);
  if ((int64_t)b.size() > cap) return -1;
  for (size_t i = 0; i < b.size(); ++i) out[i] = b[i];
  return (int64_t)b.size();
}

double dc_secular_root(int K, int i, const double* d, const double* z2, double rho, double* delta, int* iters) {
  return secular_root(K, i, d, z2, rho, delta, 1, iters);
}

// D: leaf eigenvalues (n), Z: block-diagonal leaf eigenvectors (n x n, column-major), e: off-diagonal of T,
// bounds: nb+1 leaf boundaries (nb a power of two).  On return D / Z hold the eigen-decomposition of T
// (unsorted).  stats[0] = sum of K over merges, stats[1] = max secular iterations, stats[2] = rotations.
int dc_merge_all(int64_t n, const double* e, const int64_t* bounds_in, int64_t nb, double* D, double* Z, int64_t* stats) {
  std::vector<int64_t> b(bounds_in, bounds_in + nb + 1);
  std::vector<double> Z2((size_t)n * n, 0.0), z(n);
  double* Zc = Z;
  double* Zn = Z2.data();
  stats[0] = stats[1] = stats[2] = 0;
  MergePlan mp;
  while (b.size() > 2) {
    std::vector<int64_t> nbounds;
    nbounds.push_back(b[0]);
    for (size_t k = 0; k + 2 <= b.size() - 1; k += 2) {
      const int64_t lo = b[k], mid = b[k + 1], hi = b[k + 2], N = hi - lo;
      const double beta = e[mid - 1];
      const double rho = 2.0 * std::fabs(beta), sgn = beta < 0.0 ? -1.0 : 1.0, isq = 1.0 / std::sqrt(2.0);
      for (int64_t c = lo; c < mid; ++c) z[c] = Zc[(mid - 1) + c * n] * isq;
      for (int64_t c = mid; c < hi; ++c) z[c] = sgn * Zc[mid + c * n] * isq;
      plan_merge(D + lo, z.data() + lo, N, rho, mp);
      for (size_t r = 0; r < mp.rot_p.size(); ++r) {
        double* x = Zc + lo + (lo + mp.rot_p[r]) * n;
        double* y = Zc + lo + (lo + mp.rot_n[r]) * n;
        const double c = mp.rot_c[r], s = mp.rot_s[r];
        for (int64_t q = 0; q < N; ++q) { const double a = x[q], bq = y[q]; x[q] = c * a + s * bq; y[q] = c * bq - s * a; }
      }
      stats[2] += (int64_t)mp.rot_p.size();
      const int K = (int)mp.nd.size();
      stats[0] += K;
      std::vector<double> dl(K), z2(K), zz(K), Dt((size_t)K * K), lam(K), zh(K);
      for (int t = 0; t < K; ++t) { dl[t] = mp.D[mp.nd[t]]; zz[t] = mp.z[mp.nd[t]]; z2[t] = zz[t] * zz[t]; }
      for (int i = 0; i < K; ++i) {
        int it = 0;
        lam[i] = secular_root(K, i, dl.data(), z2.data(), rho, Dt.data() + i, K, &it);   // Dt[i + j K] = d_j - lam_i
        stats[1] = std::max<int64_t>(stats[1], it);
      }
      for (int j = 0; j < K; ++j) {
        double pr = Dt[j + (size_t)j * K];
        for (int i = 0; i < K; ++i) if (i != j) pr *= Dt[i + (size_t)j * K] / (dl[j] - dl[i]);
        zh[j] = std::copysign(std::sqrt(std::fabs(pr)), zz[j]);
      }
      for (int i = 0; i < K; ++i) {
        double s = 0.0;
        for (int j = 0; j < K; ++j) { const double v = zh[j] / Dt[i + (size_t)j * K]; s += v * v; }
        const double inv = 1.0 / std::sqrt(s);
        for (int j = 0; j < K; ++j) Dt[i + (size_t)j * K] = zh[j] / Dt[i + (size_t)j * K] * inv;
      }
      // Zn[lo:hi, lo + i] = sum_j Zc[lo:hi, lo + nd[j]] * Ut[i, j]
      for (int i = 0; i < K; ++i) {
        double* out = Zn + lo + (lo + i) * n;
        for (int64_t q = 0; q < N; ++q) out[q] = 0.0;
        for (int j = 0; j < K; ++j) {
          const double u = Dt[i + (size_t)j * K];
          const double* src = Zc + lo + (lo + mp.nd[j]) * n;
          for (int64_t q = 0; q < N; ++q) out[q] += src[q] * u;
        }
      }
      std::vector<double> Dn(N);
      for (int i = 0; i < K; ++i) Dn[i] = lam[i];
      for (size_t t = 0; t < mp.df.size(); ++t) {
        std::memcpy(Zn + lo + (lo + K + (int64_t)t) * n, Zc + lo + (lo + mp.df[t]) * n, sizeof(double) * N);
        Dn[K + t] = mp.D[mp.df[t]];
      }
      std::memcpy(D + lo, Dn.data(), sizeof(double) * N);
      nbounds.push_back(hi);
    }
    b = nbounds;
    std::swap(Zc, Zn);
    // blocks of the retired buffer outside the merged ranges are never read again
  }
  if (Zc != Z) std::memcpy(Z, Zc, sizeof(double) * (size_t)n * n);
  return 0;
}

}  // extern "C"

// ---- band -> tridiagonal by bulge chasing (networksolvers_b200/csrc/sbr_chase.h, round-2 groundwork) ----------------
#include "../networksolvers_b200/csrc/sbr_chase.h"

extern "C" {

// ab: lower band storage (ld x n, ld >= 2 b + 1, rows b + 1 .. 2 b zero on entry); V2: n x n, zero on entry, receives the
// reflectors of sweep j in column j (rows j + 1 ..); tau2: ldtau x n.  wavefront = 1 runs the tasks in the order
// t = 3 j + s (all tasks of one t are independent), 0 sweep after sweep; staged = 1 uses the shared-memory form of the task.
// Returns the number of tasks run.
int64_t sbr_chase_all(int64_t n, int b, double* ab, int64_t ld, double* V2, double* tau2, int64_t ldtau, int wavefront, int staged) {
  nsb::sbr::Band B{ab, ld, n, b};
  nsb::sbr::SerialTeam tm;
  std::vector<double> v(b), work(2 * (size_t)b), red(64), stage(3 * (size_t)b * b);
  int64_t ntask = 0;
  auto run = [&](int64_t j, int s) {
    double tau = 0.0;
    int len = staged ? nsb::sbr::chase_task_staged(tm, B, j, s, v.data(), &tau, work.data(), red.data(), stage.data())
                     : nsb::sbr::chase_task(tm, B, j, s, v.data(), &tau, work.data(), red.data());
    const int64_t r0 = j + 1 + (int64_t)s * b;
    for (int i = 0; i < len && len >= 2; ++i) V2[(r0 + i) + j * n] = v[i];
    tau2[s + j * ldtau] = tau;
    ++ntask;
  };
  if (!wavefront) {
    for (int64_t j = 0; j + 2 < n; ++j)
      for (int s = 0; s < (int)nsb::sbr::nsteps(n, b, j); ++s) run(j, s);
  } else {
    const int64_t tmax = 3 * (n - 2) + nsb::sbr::nsteps(n, b, 0);
    for (int64_t t = 0; t <= tmax; ++t)
      for (int64_t j = 0; j + 2 < n && 3 * j <= t; ++j) {
        const int64_t s = t - 3 * j;
        if (s >= 0 && s < nsb::sbr::nsteps(n, b, j)) run(j, (int)s);
      }
  }
  return ntask;
}

}  // extern "C"
