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
template <bool EXACT_W = true>
__device__ __forceinline__ void scatter_level(const HashGridDev &G, float2 *__restrict__ tab, int level,
                                              const float xv[3], float g0, float g1, int lane) {
  const bool nz = (g0 != 0.f) || (g1 != 0.f);
  if (__ballot_sync(0xffffffffu, nz) == 0u) return;   // nothing to add anywhere in this warp
  Cell c;
  point_cell<EXACT_W>(G, level, xv, c);
  // run heads: first lane, or voxel differs from the previous lane's
  const uint32_t px = __shfl_up_sync(0xffffffffu, c.hx0, 1), py = __shfl_up_sync(0xffffffffu, c.hy0, 1),
                 pz = __shfl_up_sync(0xffffffffu, c.hz0, 1);
  const bool head = (lane == 0) || (px != c.hx0) || (py != c.hy0) || (pz != c.hz0);
  const uint32_t heads = __ballot_sync(0xffffffffu, head);
  if (__popc(heads) > 20) {                       // (almost) no sharing in this warp: direct atomics
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

}  // namespace pn
