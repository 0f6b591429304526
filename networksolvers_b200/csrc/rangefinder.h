// Blocked randomised range finder (src/sketched_linear_algebra/range_finder.jl:6-64); see rangefinder.cu.
#pragma once
#include <functional>

#include "common.h"

namespace nsb {

// Y (m x p, ld m) = linear_map(Omega (ndom x p, ld ndom)), both on the device
template <typename T> using RangeMap = std::function<void(const T* Omega, T* Y, int64_t p)>;

// Orthonormal basis Q (m x rank, ld m, device; capacity min(max_rank + oversample, m, ndom) columns) of the range of the
// map.  probes: caller-supplied domain vectors (ndom x sketch, device) or null (Philox N(0,1) with `seed`).
// Returns the rank; norms_out (optional) receives the residual norm of every accepted vector.
template <typename T>
int64_t range_finder_blocked(Ctx* ctx, int64_t m, int64_t ndom, const RangeMap<T>& apply_map, const T* probes, uint64_t seed,
                             int64_t max_rank, int oversample, int north_pass, double thr, double cutoff, T* Q,
                             std::vector<double>* norms_out = nullptr);

}  // namespace nsb
