// Per-element arithmetic of the data formats either side of the hot path, as host/device
// functions shared by dataio.cu and the test-only host emulation:
//   * training rays straight from (image, pixel) ids in the arithmetic of get_rays_np (float64, then the
//     astype(float32) of run_nerf.py:905) or of get_rays (float32),
//   * to8b, squared error and SSIM window statistics for test-set evaluation,
//   * integer codes of the A-CAQ fake-quantiser and their bit-packing at the learned width.
#pragma once
#include <string.h>

#include "hash_core.cuh"
#include "ray_core.cuh"

#if defined(__CUDA_ARCH__)
PN_HD double pn_dmul(double a, double b) { return __dmul_rn(a, b); }
PN_HD double pn_dadd(double a, double b) { return __dadd_rn(a, b); }
PN_HD double pn_dsub(double a, double b) { return __dsub_rn(a, b); }
PN_HD double pn_ddiv(double a, double b) { return __ddiv_rn(a, b); }
#else
PN_HD double pn_dmul(double a, double b) { volatile double r = a * b; return r; }
PN_HD double pn_dadd(double a, double b) { volatile double r = a + b; return r; }
PN_HD double pn_dsub(double a, double b) { volatile double r = a - b; return r; }
PN_HD double pn_ddiv(double a, double b) { volatile double r = a / b; return r; }
#endif

namespace pn {

// get_rays_np (run_nerf_helpers.py:323-330) as run_nerf.py:899 calls it: K is a float64 ndarray, so
// (i - K[0][2]) / K[0][0] promotes the float32 pixel grid to float64 (numpy >= 2, NEP 50); c2w is float32 and
// is promoted in the product; np.sum over the 3-vector adds left to right; run_nerf.py:905 rounds to float32.
struct CamF64 {
  double fx, fy, cx, cy;
};

PN_HD void ray_dir_f64(const CamF64 &cam, const float *R, int64_t r_stride, int i, int j, float d[3]) {
  const double d0 = pn_ddiv(pn_dsub((double)i, cam.cx), cam.fx);
  const double d1 = pn_ddiv(-pn_dsub((double)j, cam.cy), cam.fy);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double s = pn_dadd(pn_dadd(pn_dmul(d0, (double)R[c * r_stride + 0]), pn_dmul(d1, (double)R[c * r_stride + 1])),
                             pn_dmul(-1.0, (double)R[c * r_stride + 2]));
    d[c] = (float)s;
  }
}

// the float32 arithmetic of get_rays (run_nerf_helpers.py:311-320) on one pixel of a pose stored in device memory
PN_HD void ray_dir_f32(const CamF64 &cam, const float *R, int64_t r_stride, int i, int j, float d[3]) {
  Cam c32;
  c32.fx = (float)cam.fx; c32.fy = (float)cam.fy; c32.cx = (float)cam.cx; c32.cy = (float)cam.cy;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k) c32.R[r][k] = R[r * r_stride + k];
  ray_dir(c32, i, j, d);
}

// to8b (run_nerf_helpers.py:13): (255*np.clip(x,0,1)).astype(np.uint8) — float32 product, truncation.
PN_HD uint8_t to8b_one(float x) {
  const float c = fminf(fmaxf(x, 0.0f), 1.0f);
  return (uint8_t)(int)pn_mul(255.0f, c);
}

// ---- A-CAQ integer codes (quantization.py:177-187, eval form) ----------------------------------------------
// q = clamp(round(x/(scale+1e-8) + zp), qmin, qmax);  code = q - qmin  in [0, 2^bits - 1];  value = (q - zp)*scale.
PN_HD uint32_t quant_code(float x, float denom, float zp, float qmin, float qmax) {
  float q = rintf(pn_add(pn_div(x, denom), zp));
  q = fminf(fmaxf(q, qmin), qmax);
  return (uint32_t)(int64_t)pn_sub(q, qmin);
}

PN_HD float quant_value(uint32_t code, float scale, float zp, float qmin) {
  const float q = pn_add((float)code, qmin);          // exact: |q| < 2^24
  return pn_mul(pn_sub(q, zp), scale);
}

// ---- hash tables held as integer codes (inference) -----------------------------------------------------------
// Level l is an array of 2^T entries; an entry is the pair (feature 0, feature 1) as two u8 codes (2 bytes), two u16
// codes (4 bytes) or two fp32 values (8 bytes, already dequantised: levels whose learned width exceeds 16 bits).
//
// Decode without an int->float conversion: with M = 1.5 * 2^23 = 0x4B400000, the bit pattern M | c IS the float
// M + c for c < 2^22, so one byte permute builds it from the loaded word and one exact subtraction of the per-level
// constant (M - (qmin - zp)) yields q - zp = c + qmin - zp — the same integer-valued float that
// pn_sub(pn_add(float(c), qmin), zp) produces (all three are integers below 2^24, hence exact).  The value is then
// (q - zp) * scale, rounded once: bit-identical to quant_value / to the quantiser's eval forward.
struct PackedDev {
  const void *t[PN_MAX_LEVELS];
  float scale[PN_MAX_LEVELS];
  float sub[PN_MAX_LEVELS];      // M - (qmin - zp), exact
  uint8_t eb[PN_MAX_LEVELS];
};

#define PN_CODE_MAGIC 0x4B400000u   // 12582912.0f

PN_HD float packed_sub_const(float zp, float qmin) { return pn_sub(12582912.0f, pn_sub(qmin, zp)); }

PN_HD uint32_t pn_byte_perm(uint32_t x, uint32_t y, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, y, s);
#else
  const uint64_t v = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
#endif
}

PN_HD float pn_bits_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// q - zp of the two codes of one entry word (u8 pair in the low 16 bits, or u16 pair), as exact integer-valued floats
PN_HD void packed_levels_q(uint32_t word, bool u16, float sub, float &d0, float &d1) {
  const uint32_t p0 = pn_byte_perm(word, PN_CODE_MAGIC, u16 ? 0x7610u : 0x7650u);
  const uint32_t p1 = pn_byte_perm(word, PN_CODE_MAGIC, u16 ? 0x7632u : 0x7651u);
  d0 = pn_sub(pn_bits_float(p0), sub);
  d1 = pn_sub(pn_bits_float(p1), sub);
}

// the dequantised pair of one entry word
PN_HD void packed_decode(const PackedDev &T, int l, uint32_t word, bool u16, float &e0, float &e1) {
  float d0, d1;
  packed_levels_q(word, u16, T.sub[l], d0, d1);
  const float scale = T.scale[l];
  e0 = pn_mul(d0, scale);
  e1 = pn_mul(d1, scale);
}

#if defined(__CUDACC__)
// The 8 corner entries of one level.  The format branch is taken once per level (uniform across the grid) and, inside
// it, all 8 loads are issued before the first decode touches a loaded word, so the gathers overlap exactly as in the
// fp32 kernels (a per-corner branch made each decode wait for its own load: 8 serialised L2 round trips per level).
// EXACT = true : e = (q - zp) * scale per gathered value (fp32 path, bit-exact with the fake-quantised fp32 entry).
// EXACT = false: e = q - zp; the caller interpolates in the code domain and multiplies the result by scale once
//                (bf16 path: 16 multiplies per level become 2; differs from EXACT by fp32 rounding only).
template <bool EXACT>
__device__ __forceinline__ void packed_gather8(const PackedDev &T, int l, const HashGridDev &G, const Cell &c, float e0[8],
                                               float e1[8]) {
  const int eb = T.eb[l];
  if (eb == 8) {
    const float2 *__restrict__ tab = reinterpret_cast<const float2 *>(T.t[l]);
    float2 e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = __ldg(tab + corner_index(G, c, k));
#pragma unroll
    for (int k = 0; k < 8; ++k) { e0[k] = e[k].x; e1[k] = e[k].y; }
    return;
  }
  uint32_t w[8];
  if (eb == 4) {
    const uint32_t *__restrict__ tab = reinterpret_cast<const uint32_t *>(T.t[l]);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = __ldg(tab + corner_index(G, c, k));
  } else {
    const uint16_t *__restrict__ tab = reinterpret_cast<const uint16_t *>(T.t[l]);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = __ldg(tab + corner_index(G, c, k));
  }
  const bool u16 = eb == 4;
  const float sub = T.sub[l], scale = T.scale[l];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    packed_levels_q(w[k], u16, sub, e0[k], e1[k]);
    if (EXACT) { e0[k] = pn_mul(e0[k], scale); e1[k] = pn_mul(e1[k], scale); }
  }
}
#endif

// 32 codes of `bits` bits each <-> `bits` little-endian 32-bit words (element e occupies stream bits [e*bits, (e+1)*bits)).
PN_HD void pack32(const uint32_t code[32], int bits, uint32_t *words) {
  uint64_t acc = 0;
  int have = 0, w = 0;
  for (int e = 0; e < 32; ++e) {
    acc |= (uint64_t)code[e] << have;
    have += bits;
    if (have >= 32) {
      words[w++] = (uint32_t)acc;
      acc >>= 32;
      have -= 32;
    }
  }
}

PN_HD uint32_t unpack_one(const uint32_t *words, int e, int bits) {
  const int bit = e * bits, w = bit >> 5, sh = bit & 31;
  uint64_t v = words[w];
  if (sh + bits > 32) v |= (uint64_t)words[w + 1] << 32;
  return (uint32_t)((v >> sh) & ((bits == 32) ? 0xffffffffull : ((1ull << bits) - 1ull)));
}

// ---- SSIM (scikit-image 0.25.2 structural_similarity, defaults: 7x7 uniform window, sample covariance,
// K1 = 0.01, K2 = 0.03; evaluation_utils.py:33 passes data_range = 1, channel_axis = 2) ------------------------
// S at one interior pixel from the five window sums over NP = win*win samples.
PN_HD double ssim_from_sums(double sx, double sy, double sxx, double syy, double sxy, int NP, double data_range) {
  const double inv = 1.0 / NP, cov_norm = (double)NP / (NP - 1);
  const double ux = sx * inv, uy = sy * inv;
  const double vx = cov_norm * (sxx * inv - ux * ux), vy = cov_norm * (syy * inv - uy * uy);
  const double vxy = cov_norm * (sxy * inv - ux * uy);
  const double C1 = (0.01 * data_range) * (0.01 * data_range), C2 = (0.03 * data_range) * (0.03 * data_range);
  return ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
}

}  // namespace pn
