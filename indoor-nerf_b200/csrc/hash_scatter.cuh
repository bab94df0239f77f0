// Backward scatter of one level of the hash grid for the 32 consecutive points held by a warp, with
// run aggregation: consecutive samples of a ray that fall into the same voxel form a contiguous run of
// lanes (samples are depth-ordered), so their 8 corner contributions are summed with a segmented
// shuffle reduction and written by the run's first lane — 8 atomics per run instead of 8 per point.
// At the coarse levels (a voxel spans 10-25 samples) this removes >90 % of the L2 atomics and all of the
// same-address serialisation inside a warp; warps whose lanes are (nearly) all in different voxels skip
// the reduction and issue their atomics directly.
// Measured and rejected (round 2): merging the x-adjacent corner rows h, h^1 of even-x voxels into one 16-byte
// red.global.add.v4.f32 — the SM's reduction port is paced by bytes, not requests (a v4 costs ~2.5x a v2), and the
// even/odd divergence adds a second pass: dense-gradient backward 6.9 -> 7.8 ms.
#pragma once
#include "hash_core.cuh"

namespace pn {

struct GradPtrs {
  float2 *t[PN_MAX_LEVELS];
};

// All 32 lanes must call this (lanes without a point pass g0 = g1 = 0 and any x).
// direct_heads: warps with more run heads than this skip the reduction (default 20); max_len: runs are reduced in
// sub-runs of at most this many lanes (power of two; default 32 = whole runs) — both tunable because the shuffle tree
// and the reductions compete for the same LSU data pipe (scripts/sweep_scatter.py).
template <bool EXACT_W = true>
__device__ __forceinline__ void scatter_level(const HashGridDev &G, float2 *__restrict__ tab, int level,
                                              const float xv[3], float g0, float g1, int lane, int direct_heads = 20,
                                              int max_len = 32) {
  const bool nz = (g0 != 0.f) || (g1 != 0.f);
  if (__ballot_sync(0xffffffffu, nz) == 0u) return;   // nothing to add anywhere in this warp
  Cell c;
  point_cell<EXACT_W>(G, level, xv, c);
  // run heads: first lane, or voxel differs from the previous lane's
  const uint32_t px = __shfl_up_sync(0xffffffffu, c.hx0, 1), py = __shfl_up_sync(0xffffffffu, c.hy0, 1),
                 pz = __shfl_up_sync(0xffffffffu, c.hz0, 1);
  const bool head = (lane == 0) || (px != c.hx0) || (py != c.hy0) || (pz != c.hz0) || ((lane & (max_len - 1)) == 0);
  const uint32_t heads = __ballot_sync(0xffffffffu, head);
  if (__popc(heads) > direct_heads) {                       // (almost) no sharing in this warp: direct atomics
    if (nz) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        atomicAdd(tab + corner_index(G, c, k), make_float2(corner_weight_times(g0, c.w, k), corner_weight_times(g1, c.w, k)));
    }
    return;
  }
  float a0[8], a1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a0[k] = corner_weight_times(g0, c.w, k);
    a1[k] = corner_weight_times(g1, c.w, k);
  }
  // end of this lane's run = next head above it
  const uint32_t above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
  const int end = above ? (__ffs(above) - 1) : 32;
  // only as many doubling steps as the longest run in this warp needs (coarse levels: 4-5, mid levels: 1-3)
  const int maxlen = (int)__reduce_max_sync(0xffffffffu, (unsigned)(head ? end - lane : 0));
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    if (off >= maxlen) break;
    const bool take = (lane + off) < end;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t0 = __shfl_down_sync(0xffffffffu, a0[k], off), t1 = __shfl_down_sync(0xffffffffu, a1[k], off);
      if (take) { a0[k] += t0; a1[k] += t1; }
    }
  }
  if (head) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (a0[k] != 0.f || a1[k] != 0.f) atomicAdd(tab + corner_index(G, c, k), make_float2(a0[k], a1[k]));
  }
}

// ------------------------------------------------------------------------------------------------------
// Serial run aggregation: ONE thread walks NS consecutive samples of a ray at one level, keeps the 8 x 2 corner sums of
// the voxel it is in and flushes them (8 reductions) when the next sample leaves that voxel.  No shuffles: in the
// warp-cooperative form above the shuffle tree is ~50 instructions per doubling step and its wavefronts share the LSU
// data pipe with the reductions themselves (ncu, round 2: 82 % busy, 48 % of it shuffles, 2.3 G of the fused backward's
// 3.1 G instructions in the scatter).  Runs that straddle two threads' segments are flushed twice — correct, and at the
// coarse levels, where it happens, reductions are few anyway.  Samples with a zero gradient take part as zeros.
// ------------------------------------------------------------------------------------------------------
template <bool EXACT_W>
__device__ __forceinline__ void scatter_segment(const HashGridDev &G, float2 *__restrict__ tab, int level,
                                                const float *__restrict__ x, const float2 *g, int ns) {
  // x: positions of the segment's first sample ([ns][3], read-only input), g: its gradients at this level ([ns] float2,
  // written earlier in this kernel: plain loads).  The sample loop stays rolled: one sample's state live at a time.
  float a0[8], a1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a0[k] = 0.f; a1[k] = 0.f; }
  uint32_t phx = 0u, phy = 0u, phz = 0u;
#pragma unroll 1
  for (int s = 0; s < ns; ++s) {
    const float xv[3] = {__ldg(x + 3 * s), __ldg(x + 3 * s + 1), __ldg(x + 3 * s + 2)};
    const float2 gs = g[s];
    Cell c;
    point_cell<EXACT_W>(G, level, xv, c);
    if (s > 0 && (c.hx0 != phx || c.hy0 != phy || c.hz0 != phz)) {
      Cell pc;
      pc.hx0 = phx; pc.hx1 = phx + 1u; pc.hy0 = phy; pc.hy1 = phy + PN_PRIME_Y; pc.hz0 = phz; pc.hz1 = phz + PN_PRIME_Z;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (a0[k] != 0.f || a1[k] != 0.f) atomicAdd(tab + corner_index(G, pc, k), make_float2(a0[k], a1[k]));
        a0[k] = 0.f; a1[k] = 0.f;
      }
    }
    phx = c.hx0; phy = c.hy0; phz = c.hz0;
    const float fx1 = c.w[0], fx0 = EXACT_W ? pn_sub(1.0f, c.w[0]) : 1.0f - c.w[0];
    const float fy1 = c.w[1], fy0 = EXACT_W ? pn_sub(1.0f, c.w[1]) : 1.0f - c.w[1];
    const float fz1 = c.w[2], fz0 = EXACT_W ? pn_sub(1.0f, c.w[2]) : 1.0f - c.w[2];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (EXACT_W) {
        a0[k] = pn_add(a0[k], corner_weight_times(gs.x, c.w, k));
        a1[k] = pn_add(a1[k], corner_weight_times(gs.y, c.w, k));
      } else {
        const float w = ((k & 4) ? fx1 : fx0) * ((k & 2) ? fy1 : fy0) * ((k & 1) ? fz1 : fz0);
        a0[k] += gs.x * w;
        a1[k] += gs.y * w;
      }
    }
  }
  if (ns > 0) {
    Cell pc;
    pc.hx0 = phx; pc.hx1 = phx + 1u; pc.hy0 = phy; pc.hy1 = phy + PN_PRIME_Y; pc.hz0 = phz; pc.hz1 = phz + PN_PRIME_Z;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (a0[k] != 0.f || a1[k] != 0.f) atomicAdd(tab + corner_index(G, pc, k), make_float2(a0[k], a1[k]));
  }
}

}  // namespace pn
