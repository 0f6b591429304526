// Divide & conquer for the symmetric tridiagonal eigenproblem -- the pieces that are pure arithmetic on O(n)
// data: leaf partition, deflation planning (host) and the secular-equation root finder (host + device).
// No CUDA types here: tests/dc_cpu_harness.cpp compiles this header with g++ to check the logic on the CPU.
//
// Merge step (Cuppen): T = diag(T1', T2') + |beta| w w^T, w = [e_last; sign(beta) e_first];  with T1' = Q1 D1 Q1^T,
// T2' = Q2 D2 Q2^T this is Q (D + rho z z^T) Q^T, z = Q^T w / sqrt(2), rho = 2 |beta|.  Deflation follows the
// LAPACK dlaed2 rules; the roots are found with the origin shifted to the nearer pole so that every difference
// d_j - lambda_i is known to high relative accuracy, which is what the Gu-Eisenstat recomputation of z needs
// for numerically orthogonal eigenvectors.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#ifdef __CUDACC__
#define NSB_HD __host__ __device__
#else
#define NSB_HD
#endif

namespace nsb {
namespace dc {

constexpr double DC_EPS = 2.220446049250313e-16;

// Leaf boundaries: 2^p leaves of (almost) equal size <= leaf.
inline std::vector<int64_t> leaf_bounds(int64_t n, int64_t leaf) {
  int64_t nl = 1;
  while (n > nl * leaf) nl *= 2;
  std::vector<int64_t> b(nl + 1);
  for (int64_t i = 0; i <= nl; ++i) b[i] = (n * i) / nl;
  return b;
}

struct MergePlan {
  std::vector<int32_t> nd, df;            // non-deflated (ascending d) / deflated local indices
  std::vector<int32_t> rot_p, rot_n;      // Givens rotations on eigenvector columns, in application order
  std::vector<double> rot_c, rot_s;
  std::vector<double> D, z;               // updated eigenvalues and coupling vector
};

// D: eigenvalues of the two sub-problems (any order), z: coupling vector (norm ~ 1), rho > 0.
inline void plan_merge(const double* Din, const double* zin, int64_t N, double rho, MergePlan& mp) {
  mp.nd.clear(); mp.df.clear(); mp.rot_p.clear(); mp.rot_n.clear(); mp.rot_c.clear(); mp.rot_s.clear();
  mp.D.assign(Din, Din + N);
  mp.z.assign(zin, zin + N);
  std::vector<double>& D = mp.D;
  std::vector<double>& z = mp.z;
  double dmax = 0.0, zmax = 0.0;
  for (int64_t i = 0; i < N; ++i) { dmax = std::max(dmax, std::fabs(D[i])); zmax = std::max(zmax, std::fabs(z[i])); }
  const double tol = 8.0 * DC_EPS * std::max(dmax, zmax);
  std::vector<int32_t> order(N);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return D[a] < D[b]; });
  if (rho * zmax <= tol) { mp.df = order; return; }
  int32_t pj = -1;
  for (int64_t t = 0; t < N; ++t) {
    const int32_t j = order[t];
    if (rho * std::fabs(z[j]) <= tol) { mp.df.push_back(j); continue; }
    if (pj < 0) { pj = j; continue; }
    double s = z[pj], c = z[j];
    const double tau = std::hypot(c, s);
    const double dd = D[j] - D[pj];
    c /= tau;
    s = -s / tau;
    if (std::fabs(dd * c * s) <= tol) {
      z[j] = tau;
      z[pj] = 0.0;
      mp.rot_p.push_back(pj); mp.rot_n.push_back(j); mp.rot_c.push_back(c); mp.rot_s.push_back(s);
      const double tt = D[pj] * c * c + D[j] * s * s;
      D[j] = D[pj] * s * s + D[j] * c * c;
      D[pj] = tt;
      mp.df.push_back(pj);
      pj = j;
    } else {
      mp.nd.push_back(pj);
      pj = j;
    }
  }
  if (pj >= 0) mp.nd.push_back(pj);
}

// Root i of  1 + rho sum_j z2[j] / (d[j] - lambda) = 0,  d ascending and distinct, z2 > 0:
// lambda_i in (d_i, d_{i+1}) for i < K-1, lambda_{K-1} in (d_{K-1}, d_{K-1} + rho sum z2).
// Writes delta[j * stride] = d_j - lambda_i (accurate differences) and returns lambda_i.
NSB_HD inline double secular_root(int K, int i, const double* d, const double* z2, double rho, double* delta,
                                  int64_t stride, int* iters_out = nullptr) {
  if (iters_out) *iters_out = 0;
  if (K == 1) {
    const double tau = rho * z2[0];
    delta[0] = -tau;
    return d[0] + tau;
  }
  const bool last = (i == K - 1);
  int org;
  double lo, hi, tau;
  bool done = false;
  if (last) {
    org = K - 1;
    double sz = 0.0;
    for (int j = 0; j < K; ++j) sz += z2[j];
    lo = 0.0;
    hi = rho * sz;
    tau = 0.5 * hi;
  } else {
    const double gap = d[i + 1] - d[i], half = 0.5 * gap, di = d[i];
    double f = 0.0;
    for (int j = 0; j < K; ++j) f += z2[j] / ((d[j] - di) - half);
    f = 1.0 + rho * f;
    if (f > 0.0) { org = i; lo = 0.0; hi = half; }
    else { org = i + 1; lo = -half; hi = 0.0; }
    tau = 0.5 * (lo + hi);
    if (f == 0.0) { tau = -half; done = true; }
  }
  const double dorg = d[org];
  const int ip = last ? K - 2 : i;   // psi: poles 0..ip (left of the root), phi: poles ip+1..K-1
  int it = 0;
  for (; !done && it < 100; ++it) {
    double psi = 0.0, phi = 0.0, dpsi = 0.0, dphi = 0.0;
    for (int j = 0; j <= ip; ++j) {
      const double dl = (d[j] - dorg) - tau, t = z2[j] / dl;
      psi += t;
      dpsi += t / dl;
    }
    for (int j = ip + 1; j < K; ++j) {
      const double dl = (d[j] - dorg) - tau, t = z2[j] / dl;
      phi += t;
      dphi += t / dl;
    }
    psi *= rho; phi *= rho; dpsi *= rho; dphi *= rho;
    const double g = 1.0 + psi + phi;
    const double erretm = 1.0 + fabs(psi) + fabs(phi);
    if (fabs(g) <= 4.0 * DC_EPS * erretm) break;
    if (g > 0.0) hi = tau; else lo = tau;
    // two-pole rational model through the bracketing poles: psi ~ s + S / (dA - x), phi ~ r + R / (dB - x)
    const double dA = (d[ip] - dorg) - tau, dB = (d[ip + 1] - dorg) - tau;
    const double S = dpsi * dA * dA, s_ = psi - dpsi * dA;
    const double R = dphi * dB * dB, r_ = phi - dphi * dB;
    const double c0 = 1.0 + s_ + r_;
    const double a = c0;
    const double b = -(c0 * (dA + dB) + S + R);
    const double cc = c0 * dA * dB + S * dB + R * dA;
    double tn = 0.0;
    bool ok = false;
    if (a == 0.0) {
      if (b != 0.0) { tn = tau - cc / b; ok = (tn > lo && tn < hi); }
    } else {
      const double disc = b * b - 4.0 * a * cc;
      if (disc >= 0.0) {
        const double sq = sqrt(disc);
        const double q = -0.5 * (b + (b >= 0.0 ? sq : -sq));
        if (q != 0.0) { tn = tau + cc / q; ok = (tn > lo && tn < hi); }
        if (!ok) { tn = tau + q / a; ok = (tn > lo && tn < hi); }
      }
    }
    if (!ok) tn = 0.5 * (lo + hi);
    const double width = hi - lo, big = fmax(fabs(lo), fabs(hi));
    const bool stuck = (tn == tau) || (width <= 2.0 * DC_EPS * big);
    tau = tn;
    if (stuck) break;
  }
  if (iters_out) *iters_out = it;
  for (int j = 0; j < K; ++j) delta[(int64_t)j * stride] = (d[j] - dorg) - tau;
  return dorg + tau;
}

}  // namespace dc
}  // namespace nsb
