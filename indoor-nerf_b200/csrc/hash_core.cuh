// Per-point arithmetic of the multiresolution hash grid, written once as host/device functions so
// that the kernels (hash_encode.cu, later the fused field kernels) and the test-only host emulation
// (tests/hostemu) execute the same source.
//
// Reference arithmetic, in the reference's op order (every op individually rounded):
//   utils.py:95-117   keep = x == max(min(x, bmax), bmin);  xc = clamp(x, bmin, bmax)
//                     g = (bmax - bmin) / res;  i = floor((xc - bmin) / g)
//                     vmin = i*g + bmin;  vmax = vmin + g
//   utils.py:13-24    h = (cx*1) ^ (cy*2654435761) ^ (cz*805459861)  & (2^T - 1)   [int64 in torch;
//                     the low T<=30 bits are those of the uint32-wrapped product, which is what we do]
//   hash_encoding.py:64   w = (x - vmin) / (vmax - vmin)      with the UNclamped x (:103)
//   hash_encoding.py:68-78  lerp along x (corner pairs c, c+4), then y, then z
#pragma once
#include <stdlib.h>

#include "pn_common.cuh"

namespace pn {

struct HashGridDev {
  float bmin[3];
  float bmax[3];
  float g[PN_MAX_LEVELS][3];   // (bmax - bmin) / res[l], IEEE fp32, computed once on the host
  float rg[PN_MAX_LEVELS][3];  // 1 / g (bf16 paths only)
  int n_levels;
  uint32_t mask;
  int pair_gather;             // fetch x-adjacent corner rows as one 16-byte pair (table set larger than the L2)
};

// utils.py:107  grid_size = (box_max - box_min) / resolution
inline HashGridDev make_grid_dev(const pn_hash_grid &h) {
  HashGridDev G;
  for (int a = 0; a < 3; ++a) { G.bmin[a] = h.box_min[a]; G.bmax[a] = h.box_max[a]; }
  for (int l = 0; l < PN_MAX_LEVELS; ++l)
    for (int a = 0; a < 3; ++a)
    {
      G.g[l][a] = (l < h.n_levels) ? pn_div(pn_sub(h.box_max[a], h.box_min[a]), h.resolution[l]) : 1.0f;
      G.rg[l][a] = 1.0f / G.g[l][a];
    }
  G.n_levels = h.n_levels;
  G.mask = (1u << h.log2_hashmap_size) - 1u;
  // Measured (round 2, 12.6 M ray-ordered points): T = 2^22 (512 MiB of tables, DRAM-bound) 2.92 -> 2.57 ms with paired
  // loads; T = 2^19 (64 MiB, L2-resident) 1.76 -> 1.84 ms and the fused forward 2.45 -> 2.75 ms: with every sector an L2
  // hit the even / odd divergence costs more than the saved sectors.  PN_PAIR_GATHER=0|1 overrides.
  static int force = -2;
  if (force == -2) {
    const char *e = getenv("PN_PAIR_GATHER");
    force = e ? atoi(e) : -1;
  }
  G.pair_gather = force >= 0 ? force : (((size_t)h.n_levels << h.log2_hashmap_size) * 8 > (size_t)100 * 1024 * 1024);
  return G;
}

struct TablePtrs {
  const float2 *t[PN_MAX_LEVELS];
};

#define PN_PRIME_Y 2654435761u
#define PN_PRIME_Z 805459861u

struct Cell {
  uint32_t hx0, hx1;   // ix, ix+1           (prime 1)
  uint32_t hy0, hy1;   // iy*P_Y, (iy+1)*P_Y
  uint32_t hz0, hz1;   // iz*P_Z, (iz+1)*P_Z
  float w[3];          // trilinear weights (may leave [0,1] for points outside the box)
};

// keep flag of one point: inside (inclusive) the box on all axes; NaN -> not kept.
PN_HD bool point_keep(const HashGridDev &G, const float x[3]) {
  bool k = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float c = fmaxf(fminf(x[a], G.bmax[a]), G.bmin[a]);
    k = k && (x[a] == c);
  }
  return k;
}

// EXACT_W = true reproduces the reference's interpolation weights bit for bit (IEEE division, no contraction);
// EXACT_W = false is for the bf16 paths, where the features are rounded to 8 bits of mantissa anyway: the voxel
// index is still the exact one, the weight is (x - vmin) * (1/g) with ordinary contraction (error ~1e-6).
template <bool EXACT_W = true>
PN_HD void point_cell(const HashGridDev &G, int level, const float x[3], Cell &c) {
  uint32_t ii[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float g = G.g[level][a];
    // torch.clamp(x, min, max) == min(max(x, lo), hi)
    const float xc = fminf(fmaxf(x[a], G.bmin[a]), G.bmax[a]);
    const float t = pn_sub(xc, G.bmin[a]);
    float fi;
    if (EXACT_W) {
      fi = floorf(pn_div(t, g));
    } else {
      // The voxel index must be the reference's floor(RN(t / g)) bit for bit, but the IEEE division (~10 instructions,
      // 48 per point) need not be taken: q = RN(t * RN(1/g)) differs from RN(t / g) by at most 3 * 2^-24 relative, so
      // floor(q) is the exact index unless q lies within that distance of an integer; only then (tolerance 2^-21 * q:
      // a few lanes in a thousand at the finest level) is the division evaluated.  Checked against the exact form on
      // random and on knife-edge points by tests/test_hostemu.py (same source, host build).
      const float q = t * G.rg[level][a];
      fi = floorf(q);
      const float fr = q - fi, tol = q * 4.76837158203125e-7f;
      if (fr < tol || fr > 1.0f - tol) fi = floorf(pn_div(t, g));
    }
    const int i = (int)fi;                                   // .int() : fi is integral, >= 0
    if (EXACT_W) {
      const float vmin = pn_add(pn_mul((float)i, g), G.bmin[a]);
      const float vmax = pn_add(vmin, g);                    // 1.0*g == g exactly
      c.w[a] = pn_div(pn_sub(x[a], vmin), pn_sub(vmax, vmin));
    } else {
      c.w[a] = (x[a] - ((float)i * g + G.bmin[a])) * G.rg[level][a];
    }
    ii[a] = (uint32_t)i;
  }
  c.hx0 = ii[0];
  c.hx1 = ii[0] + 1u;
  c.hy0 = ii[1] * PN_PRIME_Y;
  c.hy1 = c.hy0 + PN_PRIME_Y;
  c.hz0 = ii[2] * PN_PRIME_Z;
  c.hz1 = c.hz0 + PN_PRIME_Z;
}

// corner id = 4*dx + 2*dy + dz   (utils.py:9)
PN_HD uint32_t corner_index(const HashGridDev &G, const Cell &c, int corner) {
  const uint32_t hx = (corner & 4) ? c.hx1 : c.hx0;
  const uint32_t hy = (corner & 2) ? c.hy1 : c.hy0;
  const uint32_t hz = (corner & 1) ? c.hz1 : c.hz0;
  return (hx ^ hy ^ hz) & G.mask;
}

// The 8 corner rows of a voxel, e0[k] / e1[k] = the two features of corner k = 4*dx + 2*dy + dz.  Corners k and k+4 differ
// only in x (prime 1): when the voxel's x index is even their rows are h and h ^ 1, i.e. ONE aligned 16-byte entry pair,
// fetched with one load instead of two.  A scattered gather costs the LSU one wavefront per distinct line it touches
// whatever the access width, and a gather that misses the L2 moves a 32-byte DRAM sector per row: pairing removes a
// quarter of both.  Only used by the stand-alone encoder (PAIR, a kernel template parameter the launcher sets from
// G.pair_gather) when the table set does not fit the L2: inside the fused forward the two-way divergence costs more than it saves at either table size
// (T = 2^22, fine pass: 3.35 -> 4.32 ms).  Values are bit-identical.
#if defined(__CUDACC__)
template <bool PAIR>
__device__ __forceinline__ void gather8(const HashGridDev &G, const float2 *__restrict__ tab, const Cell &c, float e0[8],
                                        float e1[8]) {
  if (PAIR && (c.hx0 & 1u) == 0u) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t i0 = corner_index(G, c, k);                        // dx = 0 row; the dx = 1 row is i0 ^ 1
      const float4 v = __ldg(reinterpret_cast<const float4 *>(tab) + (i0 >> 1));
      const bool odd = (i0 & 1u) != 0u;
      e0[k] = odd ? v.z : v.x;      e1[k] = odd ? v.w : v.y;
      e0[k + 4] = odd ? v.x : v.z;  e1[k + 4] = odd ? v.y : v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 e = __ldg(tab + corner_index(G, c, k));
      e0[k] = e.x; e1[k] = e.y;
    }
  }
}
#endif

// LearnedBitwidthQuantizer.forward on one value (quantization.py:177-187); q = row of PN_QROW floats.
PN_HD float fake_quant(float x, float scale, float denom, float zp, float qmin, float qmax, bool train_form) {
  float q = rintf(pn_add(pn_div(x, denom), zp));             // torch.round = half to even
  q = fminf(fmaxf(q, qmin), qmax);
  const float dq = pn_mul(pn_sub(q, zp), scale);
  return train_form ? pn_add(x, pn_sub(dq, x)) : dq;
}

// fake_quant without the IEEE division on the common path, same result bit for bit: RN(x * RN(1/denom)) is within
// 2^-22 relative of RN(x / denom), so after adding the zero point the rounded integer can only differ from the exact
// form's when the sum lies within tol of a half-integer — only then is the division evaluated (a few values in ten
// thousand at 8 bits).  ~15 instead of ~40 instructions per value; used where a kernel quantises activations (64 values per
// point in the tensor-core MLP).  Checked against fake_quant on random and knife-edge inputs by tests/test_hostemu.py.
PN_HD float fake_quant_rcp(float x, float scale, float denom, float rdenom, float zp, float qmin, float qmax,
                           bool train_form) {
  const float a = pn_mul(x, rdenom);
  const float s = pn_add(a, zp);
  float q = rintf(s);
  const float tol = pn_mul(pn_add(fabsf(a), fabsf(s)), 2.384185791015625e-7f);       // 2^-22
  if (fabsf(pn_sub(s, q)) >= pn_sub(0.5f, tol)) q = rintf(pn_add(pn_div(x, denom), zp));
  q = fminf(fmaxf(q, qmin), qmax);
  const float dq = pn_mul(pn_sub(q, zp), scale);
  return train_form ? pn_add(x, pn_sub(dq, x)) : dq;
}

// The same fake-quant for the bf16 paths: multiply by the reciprocal of the divisor instead of an IEEE division (the
// rounded integer can differ from the exact form only when x/denom + zp sits within an ulp of a tie).
PN_HD float fake_quant_fast(float x, float scale, float rdenom, float zp, float qmin, float qmax, bool train_form) {
  float q = rintf(x * rdenom + zp);
  q = fminf(fmaxf(q, qmin), qmax);
  const float dq = (q - zp) * scale;
  return train_form ? x + (dq - x) : dq;
}

// contraction-friendly trilinear interpolation for the bf16 paths
PN_HD float trilerp_fast(const float e[8], const float w[3]) {
  const float c00 = e[0] + w[0] * (e[4] - e[0]), c01 = e[1] + w[0] * (e[5] - e[1]);
  const float c10 = e[2] + w[0] * (e[6] - e[2]), c11 = e[3] + w[0] * (e[7] - e[3]);
  const float c0 = c00 + w[1] * (c10 - c00), c1 = c01 + w[1] * (c11 - c01);
  return c0 + w[2] * (c1 - c0);
}

// hash_encoding.py:68-78 for one feature channel; e[corner].
PN_HD float trilerp(const float e[8], const float w[3]) {
  const float wx = w[0], wy = w[1], wz = w[2];
  const float ox = pn_sub(1.0f, wx), oy = pn_sub(1.0f, wy), oz = pn_sub(1.0f, wz);
  const float c00 = pn_add(pn_mul(e[0], ox), pn_mul(e[4], wx));
  const float c01 = pn_add(pn_mul(e[1], ox), pn_mul(e[5], wx));
  const float c10 = pn_add(pn_mul(e[2], ox), pn_mul(e[6], wx));
  const float c11 = pn_add(pn_mul(e[3], ox), pn_mul(e[7], wx));
  const float c0 = pn_add(pn_mul(c00, oy), pn_mul(c10, wy));
  const float c1 = pn_add(pn_mul(c01, oy), pn_mul(c11, wy));
  return pn_add(pn_mul(c0, oz), pn_mul(c1, wz));
}

// d(trilerp)/d(e[corner]) in autograd's multiplication order: ((g*wz')*wy')*wx'.
PN_HD float corner_weight_times(float g, const float w[3], int corner) {
  const float fz = (corner & 1) ? w[2] : pn_sub(1.0f, w[2]);
  const float fy = (corner & 2) ? w[1] : pn_sub(1.0f, w[1]);
  const float fx = (corner & 4) ? w[0] : pn_sub(1.0f, w[0]);
  return pn_mul(pn_mul(pn_mul(g, fz), fy), fx);
}

// Degree-4 real SH (hash_encoding.py:153-191), python-float constants rounded to fp32 where torch
// multiplies a tensor by a python scalar.
PN_HD void sh4(float x, float y, float z, float o[16]) {
  const float C1 = (float)(0.4886025119029199);
  const float xx = pn_mul(x, x), yy = pn_mul(y, y), zz = pn_mul(z, z);
  const float xy = pn_mul(x, y), yz = pn_mul(y, z), xz = pn_mul(x, z);
  o[0] = (float)(0.28209479177387814);
  o[1] = pn_mul(-C1, y);
  o[2] = pn_mul(C1, z);
  o[3] = pn_mul(-C1, x);
  o[4] = pn_mul((float)(1.0925484305920792), xy);
  o[5] = pn_mul((float)(-1.0925484305920792), yz);
  o[6] = pn_mul((float)(0.31539156525252005), pn_sub(pn_sub(pn_mul(2.0f, zz), xx), yy));
  o[7] = pn_mul((float)(-1.0925484305920792), xz);
  o[8] = pn_mul((float)(0.5462742152960396), pn_sub(xx, yy));
  o[9] = pn_mul(pn_mul((float)(-0.5900435899266435), y), pn_sub(pn_mul(3.0f, xx), yy));
  o[10] = pn_mul(pn_mul((float)(2.890611442640554), xy), z);
  const float t4 = pn_sub(pn_sub(pn_mul(4.0f, zz), xx), yy);
  o[11] = pn_mul(pn_mul((float)(-0.4570457994644658), y), t4);
  o[12] = pn_mul(pn_mul((float)(0.3731763325901154), z),
                 pn_sub(pn_sub(pn_mul(2.0f, zz), pn_mul(3.0f, xx)), pn_mul(3.0f, yy)));
  o[13] = pn_mul(pn_mul((float)(-0.4570457994644658), x), t4);
  o[14] = pn_mul(pn_mul((float)(1.445305721320277), z), pn_sub(xx, yy));
  o[15] = pn_mul(pn_mul((float)(-0.5900435899266435), x), pn_sub(xx, pn_mul(3.0f, yy)));
}

}  // namespace pn
