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

// Bitonic sort of n <= 512 keys per ray in shared memory (ascending, values only).
__global__ void __launch_bounds__(kSampWarps * 32)
sort_merge_kernel(const float *__restrict__ a, int sa, const float *__restrict__ b, int sb, int64_t N,
                  float *__restrict__ out) {
  __shared__ float s_key[kSampWarps][kMaxBins];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *key = s_key[warp];
  const int n = sa + sb;
  int np2 = 32;
  while (np2 < n) np2 <<= 1;
  for (int64_t r = (int64_t)blockIdx.x * kSampWarps + warp; r < N; r += (int64_t)gridDim.x * kSampWarps) {
    for (int i = lane; i < np2; i += 32)
      key[i] = (i < sa) ? a[r * sa + i] : ((i < n) ? b[r * sb + (i - sa)] : INFINITY);
    __syncwarp();
    for (int k = 2; k <= np2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < np2; i += 32) {
          const int partner = i ^ j;
          if (partner > i) {
            const float x = key[i], y = key[partner];
            const bool up = ((i & k) == 0);
            if ((x > y) == up) { key[i] = y; key[partner] = x; }
          }
        }
        __syncwarp();
      }
    }
    for (int i = lane; i < n; i += 32) out[r * n + i] = key[i];
    __syncwarp();
  }
}

static int samp_blocks(int64_t N) {
  const int64_t need = ceil_div(N, kSampWarps);
  const int64_t cap = (int64_t)sm_count() * 16;
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
  sample_pdf_kernel<false><<<samp_blocks(n_rays), kSampWarps * 32, 0, as_stream(stream)>>>(
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
  sample_pdf_kernel<true><<<samp_blocks(n_rays), kSampWarps * 32, 0, as_stream(stream)>>>(
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
  sort_merge_kernel<<<samp_blocks(n_rays), kSampWarps * 32, 0, as_stream(stream)>>>(a, sa, b, sb, n_rays, out);
  count_launch();
  return check_launch("sort_merge_kernel");
}
