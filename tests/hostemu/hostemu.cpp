// TEST INFRASTRUCTURE — host build of the kernels' per-element arithmetic.
//
// The bit-exact parts of the CUDA path (hash cell / index / trilinear interpolation / fake-quant, SH,
// cdf inversion, ray generation, coarse depths, o + d*z) are written once as host/device functions in
// indoor-nerf_b200/csrc/*_core.cuh.  This file compiles those same functions for the host (g++,
// -ffp-contract=off) behind a tiny C interface so that the CPU test-suite can check them against the
// golden vectors without a GPU.  It is loaded by tests/test_hostemu.py only; the product library
// (libpocketnerf.so) contains none of it and has no CPU path.
#include <cstring>
#include "../../indoor-nerf_b200/csrc/hash_core.cuh"
#include "../../indoor-nerf_b200/csrc/ray_core.cuh"
#include "../../indoor-nerf_b200/csrc/sample_core.cuh"
#include "../../indoor-nerf_b200/csrc/io_core.cuh"

using namespace pn;

extern "C" {

void emu_hash_encode(const pn_hash_grid *grid, const float *const *tables, const float *qparams, const float *x,
                     int64_t P, float *feat, uint8_t *keep, int32_t *idx) {
  const HashGridDev G = make_grid_dev(*grid);
  const int L = G.n_levels;
  for (int64_t p = 0; p < P; ++p) {
    const float xv[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
    for (int l = 0; l < L; ++l) {
      Cell c;
      point_cell(G, l, xv, c);
      float e0[8], e1[8];
      for (int k = 0; k < 8; ++k) {
        const uint32_t h = corner_index(G, c, k);
        if (idx) idx[(p * L + l) * 8 + k] = (int32_t)h;
        e0[k] = tables[l][2 * h];
        e1[k] = tables[l][2 * h + 1];
        if (qparams && qparams[l * PN_QROW + 5] != 0.f) {
          const float *q = qparams + l * PN_QROW;
          e0[k] = fake_quant(e0[k], q[0], q[1], q[2], q[3], q[4], q[6] != 0.f);
          e1[k] = fake_quant(e1[k], q[0], q[1], q[2], q[3], q[4], q[6] != 0.f);
        }
      }
      feat[p * 2 * L + 2 * l] = trilerp(e0, c.w);
      feat[p * 2 * L + 2 * l + 1] = trilerp(e1, c.w);
    }
    if (keep) keep[p] = point_keep(G, xv) ? 1 : 0;
  }
}

void emu_hash_bwd(const pn_hash_grid *grid, float *const *dtables, const float *x, const float *dfeat, int64_t P) {
  const HashGridDev G = make_grid_dev(*grid);
  const int L = G.n_levels;
  for (int64_t p = 0; p < P; ++p) {
    const float xv[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
    for (int l = 0; l < L; ++l) {
      Cell c;
      point_cell(G, l, xv, c);
      for (int k = 0; k < 8; ++k) {
        const uint32_t h = corner_index(G, c, k);
        dtables[l][2 * h] += corner_weight_times(dfeat[p * 2 * L + 2 * l], c.w, k);
        dtables[l][2 * h + 1] += corner_weight_times(dfeat[p * 2 * L + 2 * l + 1], c.w, k);
      }
    }
  }
}

// Voxel indices of the bf16 paths' point_cell<false> (reciprocal multiply + exact-division fallback) next to the
// reference form point_cell<true>; returns the number of (point, level, axis) triples where they differ (must be 0)
// and, in *n_fallback, how often the fallback division was taken.
int64_t emu_cell_index_mismatches(const pn_hash_grid *grid, const float *x, int64_t P, int64_t *n_fallback) {
  const HashGridDev G = make_grid_dev(*grid);
  int64_t bad = 0, fb = 0;
  for (int64_t p = 0; p < P; ++p) {
    const float xv[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
    for (int l = 0; l < G.n_levels; ++l) {
      Cell a, b;
      point_cell<true>(G, l, xv, a);
      point_cell<false>(G, l, xv, b);
      bad += (a.hx0 != b.hx0) + (a.hy0 != b.hy0) + (a.hz0 != b.hz0);
      for (int ax = 0; ax < 3; ++ax) {
        const float xc = fminf(fmaxf(xv[ax], G.bmin[ax]), G.bmax[ax]);
        const float q = (xc - G.bmin[ax]) * G.rg[l][ax];
        const float fr = q - floorf(q), tol = q * 4.76837158203125e-7f;
        fb += (fr < tol || fr > 1.0f - tol);
      }
    }
  }
  if (n_fallback) *n_fallback = fb;
  return bad;
}

// fake_quant_rcp (reciprocal multiply + exact-division fallback) next to fake_quant: number of values whose results
// differ in any bit (must be 0); *n_fallback = how often the fallback division was taken.
int64_t emu_fake_quant_mismatches(const float *x, int64_t n, float scale, float denom, float zp, float qmin, float qmax,
                                  int train_form, int64_t *n_fallback) {
  const float rdenom = pn_div(1.0f, denom);
  int64_t bad = 0, fb = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float a = fake_quant(x[i], scale, denom, zp, qmin, qmax, train_form != 0);
    const float b = fake_quant_rcp(x[i], scale, denom, rdenom, zp, qmin, qmax, train_form != 0);
    uint32_t ua, ub;
    memcpy(&ua, &a, 4);
    memcpy(&ub, &b, 4);
    bad += (ua != ub) && !(a != a && b != b);
    const float p = pn_mul(x[i], rdenom), s = pn_add(p, zp);
    const float tol = pn_mul(pn_add(fabsf(p), fabsf(s)), 2.384185791015625e-7f);
    fb += fabsf(pn_sub(s, rintf(s))) >= pn_sub(0.5f, tol);
  }
  if (n_fallback) *n_fallback = fb;
  return bad;
}

void emu_sh4(const float *dirs, int64_t n, float *out) {
  for (int64_t i = 0; i < n; ++i) sh4(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], out + 16 * i);
}

void emu_sample_from_cdf(const float *cdf, const float *bins, const float *u, int64_t u_stride, int64_t N, int nb,
                         int M, float *samples, int32_t *inds) {
  for (int64_t r = 0; r < N; ++r)
    for (int j = 0; j < M; ++j) {
      int ind;
      samples[r * M + j] = invert_cdf(cdf + r * nb, bins + r * nb, nb, u[r * u_stride + j], &ind);
      if (inds) inds[r * M + j] = ind;
    }
}

void emu_gen_rays(int H, int W, const float *K, const float *c2w, float *rays_o, float *rays_d) {
  Cam cam;
  cam.fx = K[0]; cam.cx = K[2]; cam.fy = K[4]; cam.cy = K[5];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) cam.R[r][c] = c2w[4 * r + c];
    cam.t[r] = c2w[4 * r + 3];
  }
  for (int j = 0; j < H; ++j)
    for (int i = 0; i < W; ++i) {
      float d[3];
      ray_dir(cam, i, j, d);
      for (int c = 0; c < 3; ++c) {
        rays_d[3 * ((int64_t)j * W + i) + c] = d[c];
        rays_o[3 * ((int64_t)j * W + i) + c] = cam.t[c];
      }
    }
}

void emu_make_points(const float *o, const float *d, const float *z, int64_t N, int S, float *pts) {
  for (int64_t r = 0; r < N; ++r)
    for (int s = 0; s < S; ++s)
      for (int c = 0; c < 3; ++c) pts[3 * (r * S + s) + c] = point_at(o[3 * r + c], d[3 * r + c], z[r * S + s]);
}

void emu_coarse_z(const float *near, const float *far, const float *t_vals, const float *t_rand, int64_t N, int S,
                  int lindisp, float *z) {
  for (int64_t r = 0; r < N; ++r)
    for (int s = 0; s < S; ++s)
      z[r * S + s] = coarse_z_at(near[r], far[r], t_vals, s, S, lindisp != 0, t_rand ? &t_rand[r * S + s] : nullptr);
}

void emu_ndc_rays(int H, int W, double focal, double near, const float *o, const float *d, int64_t n, float *oo,
                  float *od) {
  const float cw = (float)(-1.0 / (W / (2.0 * focal))), ch = (float)(-1.0 / (H / (2.0 * focal)));
  for (int64_t p = 0; p < n; ++p) ndc_ray(cw, ch, (float)near, (float)(2.0 * near), o + 3 * p, d + 3 * p, oo + 3 * p, od + 3 * p);
}

// ---- data formats either side of the path (io_core.cuh) -------------------------------------------------------
void emu_ray_bank(const int64_t *ids, int64_t B, int H, int W, const double *K, const float *poses, int64_t pose_stride,
                  const int32_t *image_index, const float *images, int f64_dirs, float *rays, float *target) {
  CamF64 cam;
  cam.fx = K[0]; cam.cx = K[2]; cam.fy = K[4]; cam.cy = K[5];
  const int64_t hw = (int64_t)H * W;
  for (int64_t b = 0; b < B; ++b) {
    const int64_t slot = ids[b] / hw, pix = ids[b] - slot * hw;
    const int j = (int)(pix / W), i = (int)(pix - (int64_t)j * W);
    const int64_t img = image_index ? image_index[slot] : slot;
    const float *c2w = poses + img * pose_stride;
    float d[3];
    if (f64_dirs) ray_dir_f64(cam, c2w, 4, i, j, d);
    else ray_dir_f32(cam, c2w, 4, i, j, d);
    for (int c = 0; c < 3; ++c) {
      rays[3 * b + c] = c2w[4 * c + 3];
      rays[3 * (B + b) + c] = d[c];
      if (target) target[3 * b + c] = images[(img * hw + pix) * 3 + c];
    }
  }
}

void emu_to8b(const float *x, int64_t n, uint8_t *out) {
  for (int64_t k = 0; k < n; ++k) out[k] = to8b_one(x[k]);
}

void emu_quant_pack(const float *x, int64_t n, const float *qrow, int bits, uint32_t *words) {
  for (int64_t g = 0; g < n / 32; ++g) {
    uint32_t code[32];
    for (int e = 0; e < 32; ++e) code[e] = quant_code(x[g * 32 + e], qrow[1], qrow[2], qrow[3], qrow[4]);
    pack32(code, bits, words + g * bits);
  }
}

void emu_quant_unpack(const uint32_t *words, int64_t n, const float *qrow, int bits, float *x) {
  for (int64_t k = 0; k < n; ++k)
    x[k] = quant_value(unpack_one(words + (k >> 5) * bits, (int)(k & 31), bits), qrow[0], qrow[2], qrow[3]);
}

double emu_ssim_sum(const float *a, const float *b, int H, int W, int C, double data_range) {
  double tot = 0.0;
  for (int c = 0; c < C; ++c)
    for (int y = 0; y + 7 <= H; ++y)
      for (int x = 0; x + 7 <= W; ++x) {
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        for (int dy = 0; dy < 7; ++dy)
          for (int dx = 0; dx < 7; ++dx) {
            const double u = a[((int64_t)(y + dy) * W + x + dx) * C + c], v = b[((int64_t)(y + dy) * W + x + dx) * C + c];
            sx += u; sy += v; sxx += u * u; syy += v * v; sxy += u * v;
          }
        tot += ssim_from_sums(sx, sy, sxx, syy, sxy, 49, data_range);
      }
  return tot;
}

void emu_quant_codes(const float *x, int64_t n, const float *qrow, int code_bytes, void *out) {
  for (int64_t k = 0; k < n; ++k) {
    const uint32_t c = quant_code(x[k], qrow[1], qrow[2], qrow[3], qrow[4]);
    if (code_bytes == 1) ((uint8_t *)out)[k] = (uint8_t)c; else ((uint16_t *)out)[k] = (uint16_t)c;
  }
}

void emu_quant_unpack_codes(const uint32_t *words, int64_t n, int bits, int code_bytes, void *out) {
  for (int64_t k = 0; k < n; ++k) {
    const uint32_t c = unpack_one(words + (k >> 5) * bits, (int)(k & 31), bits);
    if (code_bytes == 1) ((uint8_t *)out)[k] = (uint8_t)c; else ((uint16_t *)out)[k] = (uint16_t)c;
  }
}

// hash_fwd_packed_kernel's per-point arithmetic: entries read as little-endian code pairs, decoded, interpolated
void emu_hash_encode_packed(const pn_hash_grid *grid, const pn_packed_tables *pk, const float *x, int64_t P, float *feat,
                            uint8_t *keep) {
  const HashGridDev G = make_grid_dev(*grid);
  const int L = G.n_levels;
  PackedDev T;
  for (int l = 0; l < L; ++l) {
    T.t[l] = pk->codes[l]; T.eb[l] = (uint8_t)pk->entry_bytes[l];
    T.scale[l] = pk->scale[l]; T.sub[l] = packed_sub_const(pk->zero_point[l], pk->qmin[l]);
  }
  for (int64_t p = 0; p < P; ++p) {
    const float xv[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
    for (int l = 0; l < L; ++l) {
      Cell c;
      point_cell(G, l, xv, c);
      float e0[8], e1[8];
      for (int k = 0; k < 8; ++k) {
        const uint32_t h = corner_index(G, c, k);
        if (T.eb[l] == 2) packed_decode(T, l, ((const uint16_t *)T.t[l])[h], false, e0[k], e1[k]);
        else if (T.eb[l] == 4) packed_decode(T, l, ((const uint32_t *)T.t[l])[h], true, e0[k], e1[k]);
        else {
          e0[k] = ((const float *)T.t[l])[2 * h];
          e1[k] = ((const float *)T.t[l])[2 * h + 1];
        }
      }
      feat[p * 2 * L + 2 * l] = trilerp(e0, c.w);
      feat[p * 2 * L + 2 * l + 1] = trilerp(e1, c.w);
    }
    if (keep) keep[p] = point_keep(G, xv) ? 1 : 0;
  }
}

}  // extern "C"
