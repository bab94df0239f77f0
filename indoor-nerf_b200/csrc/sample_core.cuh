// Inversion of a per-ray cdf (run_nerf_helpers.py:381-397) as host/device functions: shared by
// sample.cu and the test-only host emulation.
#pragma once
#include "pn_common.cuh"

namespace pn {

// first index i in [0, n) with cdf[i] > u, or n  (torch.searchsorted(..., right=True))
PN_HD int upper_bound(const float *cdf, int n, float u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;       // NaN u: comparisons false -> lo grows -> n, as torch
  }
  return lo;
}

PN_HD float invert_cdf(const float *cdf, const float *bins, int nb, float u, int *ind_out) {
  const int ind = upper_bound(cdf, nb, u);
  const int below = ind - 1 > 0 ? ind - 1 : 0;                       // run_nerf_helpers.py:382
  const int above = ind < nb - 1 ? ind : nb - 1;                     // :383
  const float cb = cdf[below], ca = cdf[above];
  float denom = pn_sub(ca, cb);                                      // :392
  if (denom < 1e-5f) denom = 1.0f;                                   // :393
  const float t = pn_div(pn_sub(u, cb), denom);                      // :394
  const float bb = bins[below], ba = bins[above];
  *ind_out = ind;
  return pn_add(bb, pn_mul(t, pn_sub(ba, bb)));                      // :395
}

}  // namespace pn
