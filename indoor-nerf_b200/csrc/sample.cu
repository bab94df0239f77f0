// K4 — hierarchical sampling: sample_pdf (run_nerf_helpers.py:354-397) and the sort of
// cat(z_vals, z_samples) that follows it in render_rays (run_nerf.py:512).  One warp per ray with the
// ray's cdf / keys in shared memory.
//
// Bit-exactness: given a cdf, the searchsorted(right=True) index and the lerp that follows are integer
// / individually-rounded fp32 work and are reproduced exactly (pn_sample_from_cdf).  The cdf itself is
// a 62-term float sum + prefix sum whose association order is an implementation detail of
// torch.sum/torch.cumsum that differs between torch's own CPU and CUDA kernels; ours is fixed and
// documented here: total = warp tree over lane partials (lane j owns elements j, j+32, ...), cdf =
// shuffle inclusive scan per 32-element round plus running carry.
#include "sample_core.cuh"

namespace pn {

constexpr int kSampWarps = 4;
constexpr int kMaxBins = 512;
#define PN_FULL 0xffffffffu

template <bool FROM_CDF>
__global__ void __launch_bounds__(kSampWarps * 32)
sample_pdf_kernel(const float *__restrict__ bins, const float *__restrict__ weights, int64_t w_stride,
                  const float *__restrict__ cdf_in, const float *__restrict__ u, int64_t u_stride, int64_t N,
                  int nb, int M, float *__restrict__ samples, int32_t *__restrict__ inds,
                  float *__restrict__ cdf_out) {
  __shared__ float s_cdf[kSampWarps][kMaxBins];
  __shared__ float s_bins[kSampWarps][kMaxBins];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *cdf = s_cdf[warp], *bn = s_bins[warp];
  for (int64_t r = (int64_t)blockIdx.x * kSampWarps + warp; r < N; r += (int64_t)gridDim.x * kSampWarps) {
    for (int i = lane; i < nb; i += 32) bn[i] = bins[r * nb + i];
    if (FROM_CDF) {
      for (int i = lane; i < nb; i += 32) cdf[i] = cdf_in[r * nb + i];
    } else {
      const int nw = nb - 1;
      const float *w = weights + r * w_stride;
      // weights + 1e-5, total                                                         :356-357
      float part = 0.f;
      for (int i = lane; i < nw; i += 32) part = pn_add(part, pn_add(w[i], 1e-5f));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part = pn_add(part, __shfl_xor_sync(PN_FULL, part, o));
      const float total = part;
      // cdf = [0, cumsum(pdf)]                                                        :358-359
      float carry = 0.f;
      if (lane == 0) cdf[0] = 0.f;
      for (int base = 0; base < nw; base += 32) {
        const int i = base + lane;
        float v = (i < nw) ? pn_div(pn_add(w[i], 1e-5f), total) : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float t = __shfl_up_sync(PN_FULL, v, o);
          if (lane >= o) v = pn_add(v, t);
        }
        if (i < nw) cdf[i + 1] = pn_add(carry, v);
        carry = pn_add(carry, __shfl_sync(PN_FULL, v, 31));
      }
    }
    __syncwarp();
    if (cdf_out)
      for (int i = lane; i < nb; i += 32) cdf_out[r * nb + i] = cdf[i];
    for (int j = lane; j < M; j += 32) {
      const float uj = u[r * u_stride + j];
      int ind;
      const float sv = invert_cdf(cdf, bn, nb, uj, &ind);
      samples[r * M + j] = sv;
      if (inds) inds[r * M + j] = ind;
    }
    __syncwarp();
  }
}

// Ascending bitonic sort of key[0, np2) (np2 a power of two >= 32) by one warp.  Each pass visits the np2/2 compare-exchange
// pairs exactly once: pair t of stride j is (i, i | j) with i = t's bits spread around bit log2(j).
__device__ __forceinline__ void warp_bitonic(float *key, int np2, int lane) {
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (np2 >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const float x = key[i], y = key[i | j];
        const bool up = ((i & k) == 0);
        if ((x > y) == up) { key[i] = y; key[i | j] = x; }
      }
      __syncwarp();
    }
  }
}

// sort(cat(a, b)) per ray, values only (run_nerf.py:512).  In render_rays `a` is the coarse pass's stratified depths —
// already ascending — and `b` the sa..sb importance samples in the order of their uniform draws.  So: sort b alone
// (bitonic over sb instead of sa + sb keys), then place every key by rank — a[i] at i + #{b < a[i]}, b[j] at
// j + #{a <= b[j]}: a permutation for any two sorted sequences, ties included.  A ray whose `a` is not ascending, or that
// holds a NaN, takes the general path (bitonic over all keys; NaNs stay where that network leaves them, as before).
__global__ void __launch_bounds__(kSampWarps * 32)
sort_merge_kernel(const float *__restrict__ a, int sa, const float *__restrict__ b, int sb, int64_t N,
                  float *__restrict__ out) {
  __shared__ float s_key[kSampWarps][kMaxBins];      // general path: all keys; fast path: the merged row
  __shared__ float s_a[kSampWarps][kMaxBins];
  __shared__ float s_b[kSampWarps][kMaxBins];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *key = s_key[warp], *ka = s_a[warp], *kb = s_b[warp];
  const int n = sa + sb;
  int np2 = 32, np2b = 32;
  while (np2 < n) np2 <<= 1;
  while (np2b < sb) np2b <<= 1;
  for (int64_t r = (int64_t)blockIdx.x * kSampWarps + warp; r < N; r += (int64_t)gridDim.x * kSampWarps) {
    bool ok = true;
    for (int i = lane; i < sa; i += 32) ka[i] = a[r * sa + i];
    for (int i = lane; i < np2b; i += 32) {
      const float v = (i < sb) ? b[r * sb + i] : INFINITY;
      ok = ok && (v == v);
      kb[i] = v;
    }
    __syncwarp();
    for (int i = lane; i + 1 < sa; i += 32) ok = ok && (ka[i] <= ka[i + 1]);      // false for a NaN too
    if (sa > 0 && lane == 0) ok = ok && (ka[sa - 1] == ka[sa - 1]);
    ok = __all_sync(PN_FULL, ok);
    if (ok) {
      bool b_sorted = true;
      for (int i = lane; i + 1 < sb; i += 32) b_sorted = b_sorted && (kb[i] <= kb[i + 1]);
      if (!__all_sync(PN_FULL, b_sorted)) warp_bitonic(kb, np2b, lane);
      for (int i = lane; i < n; i += 32) {
        const bool from_a = i < sa;
        const float v = from_a ? ka[i] : kb[i - sa];
        const float *other = from_a ? kb : ka;
        int lo = 0, hi = from_a ? sb : sa;            // #{other < v} (a keys) or #{other <= v} (b keys)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const float o = other[mid];
          const bool below = from_a ? (o < v) : (o <= v);
          if (below) lo = mid + 1; else hi = mid;
        }
        key[(from_a ? i : i - sa) + lo] = v;
      }
    } else {
      for (int i = lane; i < np2; i += 32) key[i] = (i < sa) ? ka[i] : ((i < n) ? kb[i - sa] : INFINITY);
      __syncwarp();
      warp_bitonic(key, np2, lane);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) out[r * n + i] = key[i];
    __syncwarp();
  }
}

// one resident wave of the grid-strided ray loop (static shared memory and registers decide how many blocks that is)
template <typename Kern>
static int samp_blocks(int64_t N, Kern kern) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSampWarps * 32, 0) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 8;
  }
  const int64_t need = ceil_div(N, kSampWarps);
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace pn

using namespace pn;

extern "C" int pn_sample_pdf(const float *bins, const float *weights, int64_t w_stride, const float *u,
                             int64_t u_stride, int64_t n_rays, int nb, int n_samples, float *samples,
                             int32_t *inds, float *cdf, pn_stream_t stream) {
  PN_REQUIRE(bins && weights && u && samples, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(nb >= 2 && nb <= kMaxBins, PN_ESHAPE, "nb %d outside 2..%d", nb, kMaxBins);
  PN_REQUIRE(n_samples >= 1, PN_EINVAL, "n_samples %d", n_samples);
  if (n_rays <= 0) return 0;
  sample_pdf_kernel<false><<<samp_blocks(n_rays, sample_pdf_kernel<false>), kSampWarps * 32, 0, as_stream(stream)>>>(
      bins, weights, w_stride, nullptr, u, u_stride, n_rays, nb, n_samples, samples, inds, cdf);
  count_launch();
  return check_launch("sample_pdf_kernel");
}

extern "C" int pn_sample_from_cdf(const float *cdf, const float *bins, const float *u, int64_t u_stride,
                                  int64_t n_rays, int nb, int n_samples, float *samples, int32_t *inds,
                                  pn_stream_t stream) {
  PN_REQUIRE(cdf && bins && u && samples, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(nb >= 2 && nb <= kMaxBins, PN_ESHAPE, "nb %d outside 2..%d", nb, kMaxBins);
  if (n_rays <= 0) return 0;
  sample_pdf_kernel<true><<<samp_blocks(n_rays, sample_pdf_kernel<true>), kSampWarps * 32, 0, as_stream(stream)>>>(
      bins, nullptr, 0, cdf, u, u_stride, n_rays, nb, n_samples, samples, inds, nullptr);
  count_launch();
  return check_launch("sample_pdf_kernel<from_cdf>");
}

extern "C" int pn_sort_merge(const float *a, int sa, const float *b, int sb, int64_t n_rays, float *out,
                             pn_stream_t stream) {
  PN_REQUIRE(a && b && out, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(sa >= 0 && sb >= 0 && sa + sb >= 1 && sa + sb <= kMaxBins, PN_ESHAPE, "sa+sb = %d outside 1..%d",
             sa + sb, kMaxBins);
  if (n_rays <= 0) return 0;
  sort_merge_kernel<<<samp_blocks(n_rays, sort_merge_kernel), kSampWarps * 32, 0, as_stream(stream)>>>(a, sa, b, sb, n_rays, out);
  count_launch();
  return check_launch("sort_merge_kernel");
}
