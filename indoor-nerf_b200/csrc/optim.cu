// Train-step surroundings that would otherwise cost ~1000 tiny framework launches per iteration
// (SURVEY.md §8f-1): the total-variation regulariser over a random cube of every hash level
// (loss.py:11-43) as one forward and one backward kernel, and the RAdam update (radam.py:28-94) as one
// elementwise pass over a flat parameter / gradient / moment buffer.
#include "hash_core.cuh"

namespace pn {

struct TvArgs {
  const float2 *tables[PN_MAX_LEVELS];
  float2 *dtables[PN_MAX_LEVELS];
  int cube[PN_MAX_LEVELS];        // cube_size per level (vertices per axis = cube + 1)
  int n_levels;
  uint32_t mask;
};

__device__ __forceinline__ uint32_t tv_hash(uint32_t x, uint32_t y, uint32_t z, uint32_t mask) {
  return (x ^ (y * PN_PRIME_Y) ^ (z * PN_PRIME_Z)) & mask;
}

// loss[l] += sum over the cube of the squared forward differences along x, y, z, divided by cube_size
__global__ void __launch_bounds__(256)
tv_fwd_kernel(const __grid_constant__ TvArgs A, const int64_t *__restrict__ min_vertex, float *__restrict__ loss) {
  const int l = blockIdx.y;
  const int c = A.cube[l], n = c + 1;
  const int total = n * n * n;
  const float2 *__restrict__ tab = A.tables[l];
  const uint32_t mx = (uint32_t)min_vertex[3 * l], my = (uint32_t)min_vertex[3 * l + 1], mz = (uint32_t)min_vertex[3 * l + 2];
  float acc = 0.f;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
    const int k = v % n, j = (v / n) % n, i = v / (n * n);
    const float2 e0 = __ldg(tab + tv_hash(mx + i, my + j, mz + k, A.mask));
    if (i < c) { const float2 e = __ldg(tab + tv_hash(mx + i + 1, my + j, mz + k, A.mask)); const float a = e.x - e0.x, b = e.y - e0.y; acc += a * a + b * b; }
    if (j < c) { const float2 e = __ldg(tab + tv_hash(mx + i, my + j + 1, mz + k, A.mask)); const float a = e.x - e0.x, b = e.y - e0.y; acc += a * a + b * b; }
    if (k < c) { const float2 e = __ldg(tab + tv_hash(mx + i, my + j, mz + k + 1, A.mask)); const float a = e.x - e0.x, b = e.y - e0.y; acc += a * a + b * b; }
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(loss + l, s / (float)c);
  }
}

// dtables[l][h(v)] += dloss[l] * 2/c * sum over the (up to 6) cube neighbours u of v of (e_v - e_u)
__global__ void __launch_bounds__(256)
tv_bwd_kernel(const __grid_constant__ TvArgs A, const int64_t *__restrict__ min_vertex, const float *__restrict__ dloss) {
  const int l = blockIdx.y;
  const int c = A.cube[l], n = c + 1;
  const int total = n * n * n;
  const float g = dloss[l] * 2.0f / (float)c;
  if (g == 0.f) return;
  const float2 *__restrict__ tab = A.tables[l];
  const uint32_t mx = (uint32_t)min_vertex[3 * l], my = (uint32_t)min_vertex[3 * l + 1], mz = (uint32_t)min_vertex[3 * l + 2];
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
    const int k = v % n, j = (v / n) % n, i = v / (n * n);
    const uint32_t h0 = tv_hash(mx + i, my + j, mz + k, A.mask);
    const float2 e0 = __ldg(tab + h0);
    float ax = 0.f, ay = 0.f;
    auto nb = [&](int di, int dj, int dk) {
      const float2 e = __ldg(tab + tv_hash(mx + i + di, my + j + dj, mz + k + dk, A.mask));
      ax += e0.x - e.x; ay += e0.y - e.y;
    };
    if (i < c) nb(1, 0, 0);
    if (i > 0) nb(-1, 0, 0);
    if (j < c) nb(0, 1, 0);
    if (j > 0) nb(0, -1, 0);
    if (k < c) nb(0, 0, 1);
    if (k > 0) nb(0, 0, -1);
    atomicAdd(A.dtables[l] + h0, make_float2(g * ax, g * ay));
  }
}

// radam.py:55-88 on n contiguous elements.  mode 2: rectified adaptive step, 1: SGD-with-momentum step
// (degenerated_to_sgd), 0: moments only.
// `dyn` (optional): the two per-step scalars read from device memory — {wd*lr, step_size*lr} — so that a launch recorded
// in a CUDA graph follows the learning-rate schedule and the rectification term when the graph is replayed.
__global__ void __launch_bounds__(256)
radam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
             int64_t n, float beta1, float beta2, float eps, float wd_lr, float step_lr, int mode,
             const float *__restrict__ dyn) {
  if (dyn != nullptr) { wd_lr = dyn[0]; step_lr = dyn[1]; }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 P = *reinterpret_cast<float4 *>(p + i);
      const float4 G = *reinterpret_cast<const float4 *>(g + i);
      float4 M = *reinterpret_cast<float4 *>(m + i), V = *reinterpret_cast<float4 *>(v + i);
      float *pp = &P.x, *mm = &M.x, *vv = &V.x;
      const float *gg = &G.x;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        vv[q] = vv[q] * beta2 + (1.0f - beta2) * gg[q] * gg[q];
        mm[q] = mm[q] * beta1 + (1.0f - beta1) * gg[q];
        if (mode != 0) {
          if (wd_lr != 0.f) pp[q] = pp[q] + (-wd_lr) * pp[q];
          pp[q] = (mode == 2) ? pp[q] + (-step_lr) * (mm[q] / (sqrtf(vv[q]) + eps)) : pp[q] + (-step_lr) * mm[q];
        }
      }
      *reinterpret_cast<float4 *>(m + i) = M;
      *reinterpret_cast<float4 *>(v + i) = V;
      if (mode != 0) *reinterpret_cast<float4 *>(p + i) = P;
    } else {
      for (int64_t e = i; e < n; ++e) {
        v[e] = v[e] * beta2 + (1.0f - beta2) * g[e] * g[e];
        m[e] = m[e] * beta1 + (1.0f - beta1) * g[e];
        if (mode != 0) {
          float x = p[e];
          if (wd_lr != 0.f) x = x + (-wd_lr) * x;
          p[e] = (mode == 2) ? x + (-step_lr) * (m[e] / (sqrtf(v[e]) + eps)) : x + (-step_lr) * m[e];
        }
      }
    }
  }
}

struct ScalarPack {
  float v[16];
};
__global__ void store_floats_kernel(float *__restrict__ dst, const ScalarPack pack, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = pack.v[threadIdx.x];
}

}  // namespace pn

using namespace pn;

static int fill_tv(TvArgs &A, const float *const *tables, float *const *dtables, const int32_t *cube, int n_levels,
                   int log2_hashmap_size) {
  PN_REQUIRE(cube && n_levels >= 1 && n_levels <= PN_MAX_LEVELS, PN_EINVAL, "bad TV arguments");
  PN_REQUIRE(log2_hashmap_size >= 1 && log2_hashmap_size <= 30, PN_ESHAPE, "log2_hashmap_size %d", log2_hashmap_size);
  for (int l = 0; l < PN_MAX_LEVELS; ++l) {
    const int s = l < n_levels ? l : 0;
    A.tables[l] = reinterpret_cast<const float2 *>(tables[s]);
    A.dtables[l] = dtables ? reinterpret_cast<float2 *>(dtables[s]) : nullptr;
    A.cube[l] = cube[s];
    PN_REQUIRE(tables[s] != nullptr && cube[s] >= 1 && cube[s] <= 255, PN_EINVAL, "level %d: table/cube", s);
  }
  A.n_levels = n_levels;
  A.mask = (1u << log2_hashmap_size) - 1u;
  return 0;
}

extern "C" int pn_tv_loss_fwd(const float *const *tables, int n_levels, int log2_hashmap_size, const int32_t *cube,
                              const int64_t *min_vertex, float *loss, pn_stream_t stream) {
  PN_REQUIRE(tables && min_vertex && loss, PN_EINVAL, "NULL pointer argument");
  TvArgs A;
  if (int e = fill_tv(A, tables, nullptr, cube, n_levels, log2_hashmap_size)) return e;
  tv_fwd_kernel<<<dim3(64, n_levels), 256, 0, as_stream(stream)>>>(A, min_vertex, loss);
  count_launch();
  return check_launch("tv_fwd_kernel");
}

extern "C" int pn_tv_loss_bwd(const float *const *tables, float *const *dtables, int n_levels, int log2_hashmap_size,
                              const int32_t *cube, const int64_t *min_vertex, const float *dloss, pn_stream_t stream) {
  PN_REQUIRE(tables && dtables && min_vertex && dloss, PN_EINVAL, "NULL pointer argument");
  TvArgs A;
  if (int e = fill_tv(A, tables, dtables, cube, n_levels, log2_hashmap_size)) return e;
  tv_bwd_kernel<<<dim3(64, n_levels), 256, 0, as_stream(stream)>>>(A, min_vertex, dloss);
  count_launch();
  return check_launch("tv_bwd_kernel");
}

extern "C" int pn_radam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float beta1,
                             float beta2, float eps, float weight_decay_times_lr, float step_size_times_lr, int mode,
                             pn_stream_t stream) {
  PN_REQUIRE(param && grad && exp_avg && exp_avg_sq, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(mode >= 0 && mode <= 2, PN_EINVAL, "mode %d", mode);
  PN_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, PN_EINVAL,
             "buffers must be 16-byte aligned");
  if (n <= 0) return 0;
  const int64_t need = ceil_div(n, 256 * 4);
  const int64_t cap = (int64_t)sm_count() * 8;
  radam_kernel<<<(unsigned)(need < cap ? need : cap), 256, 0, as_stream(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, weight_decay_times_lr, step_size_times_lr, mode, nullptr);
  count_launch();
  return check_launch("radam_kernel");
}

extern "C" int pn_radam_step_dyn(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float beta1,
                                 float beta2, float eps, const float *dyn, int mode, pn_stream_t stream) {
  PN_REQUIRE(param && grad && exp_avg && exp_avg_sq && dyn, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(mode >= 0 && mode <= 2, PN_EINVAL, "mode %d", mode);
  PN_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, PN_EINVAL,
             "buffers must be 16-byte aligned");
  if (n <= 0) return 0;
  const int64_t need = ceil_div(n, 256 * 4);
  const int64_t cap = (int64_t)sm_count() * 8;
  radam_kernel<<<(unsigned)(need < cap ? need : cap), 256, 0, as_stream(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, 0.f, 0.f, mode, dyn);
  count_launch();
  return check_launch("radam_kernel<dyn>");
}

extern "C" int pn_store_floats(float *dst, const float *values, int n, pn_stream_t stream) {
  PN_REQUIRE(dst && values, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(n >= 1 && n <= 16, PN_EINVAL, "n %d outside 1..16", n);
  ScalarPack pack;
  for (int i = 0; i < 16; ++i) pack.v[i] = i < n ? values[i] : 0.f;     // host values travel as kernel arguments
  store_floats_kernel<<<1, 32, 0, as_stream(stream)>>>(dst, pack, n);
  count_launch();
  return check_launch("store_floats_kernel");
}
