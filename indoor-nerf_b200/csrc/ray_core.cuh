// Ray arithmetic as host/device functions (shared by rays.cu, the fused kernels and the test-only host
// emulation).  Every op individually rounded, in the reference's order.
#pragma once
#include "pn_common.cuh"

namespace pn {

struct Cam {
  float fx, fy, cx, cy;
  float R[3][3];
  float t[3];
};

// get_rays (run_nerf_helpers.py:311-320): dirs = [(i-cx)/fx, -(j-cy)/fy, -1]; d[c] = sum_k dirs[k]*c2w[c][k]
PN_HD void ray_dir(const Cam &cam, int i, int j, float d[3]) {
  const float d0 = pn_div(pn_sub((float)i, cam.cx), cam.fx);
  const float d1 = pn_div(-pn_sub((float)j, cam.cy), cam.fy);
  const float d2 = -1.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    d[c] = pn_add(pn_add(pn_mul(d0, cam.R[c][0]), pn_mul(d1, cam.R[c][1])), pn_mul(d2, cam.R[c][2]));
}

// pts = o + d*z  (run_nerf.py:490,513): mul, then add
PN_HD float point_at(float o, float d, float z) { return pn_add(o, pn_mul(d, z)); }

PN_HD float coarse_z_plain(float nr, float fr, float t, bool lindisp) {
  if (!lindisp) return pn_add(pn_mul(nr, pn_sub(1.0f, t)), pn_mul(fr, t));                                  // run_nerf.py:468
  return pn_div(1.0f, pn_add(pn_mul(pn_div(1.0f, nr), pn_sub(1.0f, t)), pn_mul(pn_div(1.0f, fr), t)));      // :470
}

// Stratified coarse depth of sample s (run_nerf.py:466-488); t_rand == nullptr means no perturbation.
PN_HD float coarse_z_at(float nr, float fr, const float *t_vals, int s, int S, bool lindisp, const float *t_rand) {
  const float zc = coarse_z_plain(nr, fr, t_vals[s], lindisp);
  if (!t_rand) return zc;
  const float upper = (s + 1 < S) ? pn_mul(0.5f, pn_add(coarse_z_plain(nr, fr, t_vals[s + 1], lindisp), zc)) : zc;   // :476-477
  const float lower = (s > 0) ? pn_mul(0.5f, pn_add(zc, coarse_z_plain(nr, fr, t_vals[s - 1], lindisp))) : zc;       // :478
  return pn_add(lower, pn_mul(pn_sub(upper, lower), *t_rand));                                                        // :488
}

// ndc_rays (run_nerf_helpers.py:333-350).  cw = -1/(W/(2 focal)), ch = -1/(H/(2 focal)) and two_near = 2*near
// are python floats in the reference (computed in double, cast to fp32 when they meet a tensor).
PN_HD void ndc_ray(float cw, float ch, float near, float two_near, const float o[3], const float d[3], float oo[3],
                   float od[3]) {
  const float t = pn_div(-pn_add(near, o[2]), d[2]);
  const float ox = pn_add(o[0], pn_mul(t, d[0])), oy = pn_add(o[1], pn_mul(t, d[1])), oz = pn_add(o[2], pn_mul(t, d[2]));
  oo[0] = pn_div(pn_mul(cw, ox), oz);
  oo[1] = pn_div(pn_mul(ch, oy), oz);
  oo[2] = pn_add(1.0f, pn_div(two_near, oz));
  od[0] = pn_mul(cw, pn_sub(pn_div(d[0], d[2]), pn_div(ox, oz)));
  od[1] = pn_mul(ch, pn_sub(pn_div(d[1], d[2]), pn_div(oy, oz)));
  od[2] = pn_div(-two_near, oz);
}

}  // namespace pn
