// K3 — alpha compositing along rays: raw2outputs (run_nerf.py:347-411) forward and backward.
// One warp per ray; sample s lives in lane s%32, slot s/32, so every global access is a coalesced
// 128-byte row.  Transmittance is an exclusive product scan (shuffle scan per 32-sample round with a
// running carry); the backward needs the matching reverse sum scan.  HBM-bound and tiny next to K1/K2:
// ~24 B/sample forward.
#include <float.h>

#include "pn_common.cuh"

namespace pn {

constexpr int kRayWarps = 4;
#define PN_FULL 0xffffffffu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PN_FULL, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// Per-lane state of one ray after the forward recurrence.
template <int K>
struct RayState {
  float zs[K], dist[K], sg[K], a[K], t[K], Tex[K], w[K];
};

template <int K>
__device__ __forceinline__ void ray_forward(const float *__restrict__ raw, int C, const float *__restrict__ z,
                                            const float *__restrict__ rays_d, const float *__restrict__ noise,
                                            int64_t r, int S, int lane, RayState<K> &st) {
  const float dx = rays_d[3 * r + 0], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
  const float dn = sqrtf(dx * dx + dy * dy + dz * dz);
  float carry = 1.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int s = k * 32 + lane;
    const bool valid = s < S;
    const float zc = valid ? z[r * S + s] : 0.f;
    float d = (s + 1 < S) ? (z[r * S + s + 1] - zc) : 1e10f;       // run_nerf.py:364-365
    d = d * dn;                                                     // :367
    float sg = 0.f;
    if (valid) {
      sg = raw[(r * S + s) * C + 3];
      if (noise) sg = sg + noise[r * S + s];
    }
    const float a = valid ? (1.0f - expf(-fmaxf(sg, 0.f) * d)) : 0.f;   // :362,388
    const float t = valid ? ((1.0f - a) + 1e-10f) : 1.0f;               // :390
    float incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float v = __shfl_up_sync(PN_FULL, incl, o);
      if (lane >= o) incl *= v;
    }
    float excl = __shfl_up_sync(PN_FULL, incl, 1);
    if (lane == 0) excl = 1.0f;
    st.Tex[k] = carry * excl;
    carry = carry * __shfl_sync(PN_FULL, incl, 31);
    st.zs[k] = zc; st.dist[k] = d; st.sg[k] = sg; st.a[k] = a; st.t[k] = t;
    st.w[k] = a * st.Tex[k];
  }
}

#ifndef PN_COMP_FWD_MINB
#define PN_COMP_FWD_MINB 8     // 64 registers: 32 warps per SM hide the scan and exp latency (0.23 -> 0.17 ms at 65536 x 192)
#endif
// N7: the raw rows carry a normal (7 channels); false compiles the normal map out (C == 4)
template <int K, bool N7>
__global__ void __launch_bounds__(kRayWarps * 32, PN_COMP_FWD_MINB)
composite_fwd_kernel(const float *__restrict__ raw, int C_, const float *__restrict__ z,
                     const float *__restrict__ rays_d, const float *__restrict__ noise, int64_t N, int S,
                     int white, int vec4, float *__restrict__ rgb, float *__restrict__ disp, float *__restrict__ acc,
                     float *__restrict__ weights, float *__restrict__ depth, float *__restrict__ sparsity,
                     float *__restrict__ normal) {
  const int C = N7 ? 7 : 4;
  (void)C_;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * kRayWarps + warp; r < N; r += (int64_t)gridDim.x * kRayWarps) {
    RayState<K> st;
    ray_forward<K>(raw, C, z, rays_d, noise, r, S, lane, st);
    float sw = 0.f, swz = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int s = k * 32 + lane;
      if (s < S) {
        const float w = st.w[k];
        const float *rw = raw + (r * S + s) * C;
        if (weights) weights[r * S + s] = w;
        sw += w;
        swz += w * st.zs[k];
        float r0, r1, r2;
        if (vec4) {
          const float4 v = *reinterpret_cast<const float4 *>(rw);
          r0 = v.x; r1 = v.y; r2 = v.z;
        } else {
          r0 = rw[0]; r1 = rw[1]; r2 = rw[2];
        }
        c0 += w * sigmoidf(r0);
        c1 += w * sigmoidf(r1);
        c2 += w * sigmoidf(r2);
        if (C == 7) { n0 += w * rw[4]; n1 += w * rw[5]; n2 += w * rw[6]; }
      }
    }
    sw = warp_sum(sw); swz = warp_sum(swz);
    c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
    if (C == 7) { n0 = warp_sum(n0); n1 = warp_sum(n1); n2 = warp_sum(n2); }
    // entropy of Categorical(probs = [w, clamp(1 - sum w, 1e-6)])   run_nerf.py:400-403
    const float rest = fmaxf(1.0f - sw, 1e-6f);
    const float Q = sw + rest;
    float h = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int s = k * 32 + lane;
      if (s < S) {
        const float p = st.w[k] / Q;
        h -= p * logf(fminf(fmaxf(p, FLT_EPSILON), 1.0f - FLT_EPSILON));
      }
    }
    if (lane == 0) {
      const float p = rest / Q;
      h -= p * logf(fminf(fmaxf(p, FLT_EPSILON), 1.0f - FLT_EPSILON));
    }
    h = warp_sum(h);
    if (lane == 0) {
      const float dep = swz / sw;                                   // :393 (NaN when sw == 0, as the reference)
      if (white) { const float bg = 1.0f - sw; c0 += bg; c1 += bg; c2 += bg; }   // :398
      if (rgb) { rgb[3 * r + 0] = c0; rgb[3 * r + 1] = c1; rgb[3 * r + 2] = c2; }
      if (depth) depth[r] = dep;
      if (disp) disp[r] = (dep != dep) ? dep : 1.0f / fmaxf(1e-10f, dep);   // :394 (torch.max propagates NaN)
      if (acc) acc[r] = sw;
      if (sparsity) sparsity[r] = h;
      if (C == 7 && normal) {
        const float nn = fmaxf(sqrtf(n0 * n0 + n1 * n1 + n2 * n2), 1e-12f);   // F.normalize, :408
        normal[3 * r + 0] = n0 / nn; normal[3 * r + 1] = n1 / nn; normal[3 * r + 2] = n2 / nn;
      }
    }
  }
}

#ifndef PN_COMP_BWD_MINB
#define PN_COMP_BWD_MINB 5     // measured: 5 blocks x 4 warps (<= 102 registers) beats 3 (143) and 6 (85, spills)
#endif
template <int K, bool N7>
__global__ void __launch_bounds__(kRayWarps * 32, K <= 6 ? PN_COMP_BWD_MINB : 3)
composite_bwd_kernel(const float *__restrict__ raw, int C_, const float *__restrict__ z,
                     const float *__restrict__ rays_d, const float *__restrict__ noise, int64_t N, int S,
                     int white, int vec4, const float *__restrict__ d_rgb, const float *__restrict__ d_disp,
                     const float *__restrict__ d_acc, const float *__restrict__ d_weights,
                     const float *__restrict__ d_depth, const float *__restrict__ d_sparsity,
                     const float *__restrict__ d_normal, float *__restrict__ draw) {
  const int C = N7 ? 7 : 4;
  (void)C_;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * kRayWarps + warp; r < N; r += (int64_t)gridDim.x * kRayWarps) {
    RayState<K> st;
    ray_forward<K>(raw, C, z, rays_d, noise, r, S, lane, st);
    float sw = 0.f, swz = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int s = k * 32 + lane;
      if (s < S) {
        sw += st.w[k];
        swz += st.w[k] * st.zs[k];
        if (C == 7 && d_normal) {
          const float *rw = raw + (r * S + s) * C;
          n0 += st.w[k] * rw[4]; n1 += st.w[k] * rw[5]; n2 += st.w[k] * rw[6];
        }
      }
    }
    sw = warp_sum(sw); swz = warp_sum(swz);
    const float gr0 = d_rgb ? d_rgb[3 * r + 0] : 0.f, gr1 = d_rgb ? d_rgb[3 * r + 1] : 0.f,
                gr2 = d_rgb ? d_rgb[3 * r + 2] : 0.f;
    float g_acc = d_acc ? d_acc[r] : 0.f;
    if (white) g_acc -= (gr0 + gr1 + gr2);
    // depth = swz / sw ; disp = 1 / max(1e-10, depth)
    const float dep = swz / sw;
    bool use_depth = false;
    float g_dep = 0.f;
    if (d_depth) { g_dep += d_depth[r]; use_depth = true; }
    if (d_disp) { g_dep += (dep > 1e-10f) ? (-d_disp[r] / (dep * dep)) : 0.f; use_depth = true; }
    const float g_num = use_depth ? g_dep / sw : 0.f;                 // 0/0 -> NaN like autograd
    const float g_den = use_depth ? -g_dep * swz / (sw * sw) : 0.f;
    // entropy
    float g_sp = 0.f, Q = 1.f, Gbar = 0.f, rest_term = 0.f;
    float Gk[K];                                                      // d entropy / d p_s, kept for the output loop
    if (d_sparsity) {
      g_sp = d_sparsity[r];
      const float rest = fmaxf(1.0f - sw, 1e-6f);
      Q = sw + rest;
      float gq = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int s = k * 32 + lane;
        Gk[k] = 0.f;
        if (s < S) {
          const float p = st.w[k] / Q;
          const bool in = (p >= FLT_EPSILON) && (p <= 1.0f - FLT_EPSILON);
          Gk[k] = -(logf(fminf(fmaxf(p, FLT_EPSILON), 1.0f - FLT_EPSILON)) + (in ? 1.0f : 0.f));
          gq += Gk[k] * st.w[k];
        }
      }
      const float pr = rest / Q;
      const bool inr = (pr >= FLT_EPSILON) && (pr <= 1.0f - FLT_EPSILON);
      const float Gr = -(logf(fminf(fmaxf(pr, FLT_EPSILON), 1.0f - FLT_EPSILON)) + (inr ? 1.0f : 0.f));
      gq = warp_sum(gq) + Gr * rest;
      Gbar = gq / (Q * Q);
      const float pass = ((1.0f - sw) >= 1e-6f) ? 1.0f : 0.f;       // clamp(min) backward
      rest_term = pass * (Gr / Q - Gbar);
    }
    // normal map
    float dv0 = 0.f, dv1 = 0.f, dv2 = 0.f;
    const bool has_n = (C == 7) && d_normal;
    if (has_n) {
      n0 = warp_sum(n0); n1 = warp_sum(n1); n2 = warp_sum(n2);
      const float nn = sqrtf(n0 * n0 + n1 * n1 + n2 * n2);
      const float g0 = d_normal[3 * r + 0], g1 = d_normal[3 * r + 1], g2 = d_normal[3 * r + 2];
      if (nn > 1e-12f) {
        const float m0 = n0 / nn, m1 = n1 / nn, m2 = n2 / nn;
        const float dot = m0 * g0 + m1 * g1 + m2 * g2;
        dv0 = (g0 - m0 * dot) / nn; dv1 = (g1 - m1 * dot) / nn; dv2 = (g2 - m2 * dot) / nn;
      } else {
        dv0 = g0 / 1e-12f; dv1 = g1 / 1e-12f; dv2 = g2 / 1e-12f;
      }
    }
    // gradient w.r.t. each weight and the reverse scan of the transmittance chain, one 32-sample round at a time from the
    // far end (the suffix sum only needs the rounds behind it): each sample's colour sigmoids are evaluated once and its
    // gradient row leaves as one 16-byte store
    float carry = 0.f;
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
      const int s = k * 32 + lane;
      const bool valid = s < S;
      const float *rw = raw + (r * S + s) * C;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, g = 0.f;
      if (valid) {
        float c0, c1, c2;
        if (vec4) {
          const float4 v = *reinterpret_cast<const float4 *>(rw);
          c0 = v.x; c1 = v.y; c2 = v.z;
        } else {
          c0 = rw[0]; c1 = rw[1]; c2 = rw[2];
        }
        g = g_acc;
        if (d_weights) g += d_weights[r * S + s];
        if (d_rgb) {
          s0 = sigmoidf(c0); s1 = sigmoidf(c1); s2 = sigmoidf(c2);
          g += gr0 * s0 + gr1 * s1 + gr2 * s2;
        }
        if (use_depth) g += g_num * st.zs[k] + g_den;
        if (d_sparsity) g += g_sp * ((Gk[k] / Q - Gbar) - rest_term);
        if (has_n) g += dv0 * rw[4] + dv1 * rw[5] + dv2 * rw[6];
      }
      const float v = valid ? g * st.w[k] : 0.f;
      float incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_down_sync(PN_FULL, incl, o);
        if (lane + o < 32) incl += u;
      }
      float excl = __shfl_down_sync(PN_FULL, incl, 1);
      if (lane == 31) excl = 0.f;
      const float suffix = excl + carry;                              // sum_{j>s} gw_j w_j
      carry += __shfl_sync(PN_FULL, incl, 0);
      if (valid) {
        float *out = draw + (r * S + s) * C;
        const float da = g * st.Tex[k] - suffix / st.t[k];
        const float sg = st.sg[k];
        const float w = st.w[k];
        const float o3 = (sg > 0.f) ? da * (st.dist[k] * expf(-sg * st.dist[k])) : 0.f;
        const float o0 = d_rgb ? gr0 * w * (s0 * (1.0f - s0)) : 0.f;
        const float o1 = d_rgb ? gr1 * w * (s1 * (1.0f - s1)) : 0.f;
        const float o2 = d_rgb ? gr2 * w * (s2 * (1.0f - s2)) : 0.f;
        if (vec4) {
          *reinterpret_cast<float4 *>(out) = make_float4(o0, o1, o2, o3);
        } else {
          out[0] = o0; out[1] = o1; out[2] = o2; out[3] = o3;
          if (C == 7) {
            out[4] = has_n ? w * dv0 : 0.f;
            out[5] = has_n ? w * dv1 : 0.f;
            out[6] = has_n ? w * dv2 : 0.f;
          }
        }
      }
    }
  }
}

// One resident wave: the ray loop is grid-strided, so a grid of exactly the co-resident blocks (queried per kernel — the
// backward holds 120+ registers) has no partial last wave.
template <typename Kern>
static int ray_blocks(int64_t N, Kern kern) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRayWarps * 32, 0) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 4;
  }
  const int64_t need = ceil_div(N, kRayWarps);
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace pn

using namespace pn;

#define PN_DISPATCH_K(S, CALL)                 \
  do {                                         \
    if ((S) <= 64) { CALL(2); }                \
    else if ((S) <= 128) { CALL(4); }          \
    else if ((S) <= 192) { CALL(6); }          \
    else if ((S) <= 256) { CALL(8); }          \
    else { CALL(16); }                         \
  } while (0)

extern "C" int pn_composite_fwd(const float *raw, int channels, const float *z, const float *rays_d,
                                const float *noise, int64_t n_rays, int n_samples, int white_bkgd, float *rgb,
                                float *disp, float *acc, float *weights, float *depth, float *sparsity,
                                float *normal, pn_stream_t stream) {
  PN_REQUIRE(raw && z && rays_d, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(channels == 4 || channels == 7, PN_ESHAPE, "channels %d (4 or 7)", channels);
  PN_REQUIRE(n_samples >= 1 && n_samples <= 512, PN_ESHAPE, "n_samples %d outside 1..512", n_samples);
  if (n_rays <= 0) return 0;
  const int vec4 = channels == 4 && ((uintptr_t)raw & 15) == 0;       // 16-byte rows: one load per sample
#define CALL(K)                                                                                          \
  composite_fwd_kernel<K, N7><<<ray_blocks(n_rays, composite_fwd_kernel<K, N7>), kRayWarps * 32, 0, as_stream(stream)>>>(                            \
      raw, channels, z, rays_d, noise, n_rays, n_samples, white_bkgd, vec4, rgb, disp, acc, weights, depth, \
      sparsity, normal)
  if (channels == 7) {
    constexpr bool N7 = true;
    PN_DISPATCH_K(n_samples, CALL);
  } else {
    constexpr bool N7 = false;
    PN_DISPATCH_K(n_samples, CALL);
  }
#undef CALL
  count_launch();
  return check_launch("composite_fwd_kernel");
}

extern "C" int pn_composite_bwd(const float *raw, int channels, const float *z, const float *rays_d,
                                const float *noise, int64_t n_rays, int n_samples, int white_bkgd,
                                const float *d_rgb, const float *d_disp, const float *d_acc,
                                const float *d_weights, const float *d_depth, const float *d_sparsity,
                                const float *d_normal, float *draw, pn_stream_t stream) {
  PN_REQUIRE(raw && z && rays_d && draw, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(channels == 4 || channels == 7, PN_ESHAPE, "channels %d (4 or 7)", channels);
  PN_REQUIRE(n_samples >= 1 && n_samples <= 512, PN_ESHAPE, "n_samples %d outside 1..512", n_samples);
  if (n_rays <= 0) return 0;
  const int vec4 = channels == 4 && (((uintptr_t)raw | (uintptr_t)draw) & 15) == 0;
#define CALL(K)                                                                                        \
  composite_bwd_kernel<K, N7><<<ray_blocks(n_rays, composite_bwd_kernel<K, N7>), kRayWarps * 32, 0, as_stream(stream)>>>(                          \
      raw, channels, z, rays_d, noise, n_rays, n_samples, white_bkgd, vec4, d_rgb, d_disp, d_acc, d_weights, \
      d_depth, d_sparsity, d_normal, draw)
  if (channels == 7) {
    constexpr bool N7 = true;
    PN_DISPATCH_K(n_samples, CALL);
  } else {
    constexpr bool N7 = false;
    PN_DISPATCH_K(n_samples, CALL);
  }
#undef CALL
  count_launch();
  return check_launch("composite_bwd_kernel");
}
