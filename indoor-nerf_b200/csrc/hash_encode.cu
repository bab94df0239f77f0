// K1 / K1b — multiresolution hash-grid encoder, forward gather and backward scatter-add.
// Replaces HashEmbedder.forward (hash_encoding.py:82-107), get_voxel_vertices (utils.py:95-117),
// hash (utils.py:13-24) and, in backward, the 16 aten::embedding_dense_backward calls.
//
// Work decomposition (B200): one thread per point, levels in an unrolled loop.  A warp therefore
// holds 32 consecutive points — consecutive samples of one ray in the render path — which share
// voxels at the coarse levels (same-address lanes coalesce into one L1 wavefront) and walk the same
// 4 MiB table at the same time at every level.  The 64 MiB (T=2^19) table set is L2-resident on a
// 126 MB L2; the [P,32] output is staged through shared memory so that each warp writes whole
// 128-byte lines.  HBM traffic per point is then the algorithmic minimum 12 B in + 128 B out.
#include "hash_core.cuh"
#include "hash_scatter.cuh"
#include "io_core.cuh"

namespace pn {


constexpr int kHashThreads = 128;
constexpr int kHashWarps = kHashThreads / 32;
constexpr int kStagePitch = 33;   // 32 floats + 1: conflict-free column writes and row reads

__device__ __forceinline__ void load_point(const float *__restrict__ x, int64_t p, float xv[3]) {
  xv[0] = __ldg(x + 3 * p + 0);
  xv[1] = __ldg(x + 3 * p + 1);
  xv[2] = __ldg(x + 3 * p + 2);
}

template <bool QUANT, bool PAIR>
__global__ void __launch_bounds__(kHashThreads)
hash_fwd_kernel(const __grid_constant__ HashGridDev G, const __grid_constant__ TablePtrs T, const float *__restrict__ qparams,
                const float *__restrict__ x, int64_t P, float *__restrict__ feat,
                uint8_t *__restrict__ keep) {
  __shared__ float stage[kHashWarps][32 * kStagePitch];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int F = 2 * G.n_levels;
  float *st = stage[warp];
  for (int64_t base = ((int64_t)blockIdx.x * kHashWarps + warp) * 32; base < P;
       base += (int64_t)gridDim.x * kHashThreads) {
    const int64_t p = base + lane;
    const bool valid = p < P;
    float xv[3] = {0.f, 0.f, 0.f};
    if (valid) load_point(x, p, xv);
#pragma unroll 4
    for (int l = 0; l < G.n_levels; ++l) {
      Cell c;
      point_cell(G, l, xv, c);
      float e0[8], e1[8];
      gather8<PAIR>(G, T.t[l], c, e0, e1);
      if (QUANT) {
        const float *q = qparams + l * PN_QROW;
        if (q[5] != 0.f) {
          const float scale = q[0], denom = q[1], zp = q[2], qmin = q[3], qmax = q[4];
          const bool train_form = q[6] != 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            e0[k] = fake_quant(e0[k], scale, denom, zp, qmin, qmax, train_form);
            e1[k] = fake_quant(e1[k], scale, denom, zp, qmin, qmax, train_form);
          }
        }
      }
      st[lane * kStagePitch + 2 * l + 0] = trilerp(e0, c.w);
      st[lane * kStagePitch + 2 * l + 1] = trilerp(e1, c.w);
    }
    if (valid && keep) keep[p] = point_keep(G, xv) ? 1 : 0;
    __syncwarp();
    // 32 rows x F floats -> global, whole lines per warp
    const int64_t rows = (P - base) < 32 ? (P - base) : 32;
    if (F == 32) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int i4 = it * 32 + lane, row = i4 >> 3, c4 = (i4 & 7) * 4;
        if (row < rows) {
          const float *s = st + row * kStagePitch + c4;
          float4 v = make_float4(s[0], s[1], s[2], s[3]);
          *reinterpret_cast<float4 *>(feat + (base + row) * 32 + c4) = v;
        }
      }
    } else {
      for (int i = lane; i < 32 * F; i += 32) {
        const int row = i / F, col = i - row * F;
        if (row < rows) feat[(base + row) * F + col] = st[row * kStagePitch + col];
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kHashThreads)
hash_bwd_kernel(const __grid_constant__ HashGridDev G, const __grid_constant__ GradPtrs D, const float *__restrict__ x,
                const float *__restrict__ dfeat, int64_t P) {
  __shared__ float stage[kHashWarps][32 * kStagePitch];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int F = 2 * G.n_levels;
  float *st = stage[warp];
  for (int64_t base = ((int64_t)blockIdx.x * kHashWarps + warp) * 32; base < P;
       base += (int64_t)gridDim.x * kHashThreads) {
    const int64_t p = base + lane;
    const bool valid = p < P;
    const int64_t rows = (P - base) < 32 ? (P - base) : 32;
    // dfeat rows -> shared, whole lines per warp
    for (int i = lane; i < 32 * F; i += 32) {
      const int row = i / F, col = i - row * F;
      st[row * kStagePitch + col] = (row < rows) ? __ldg(dfeat + (base + row) * F + col) : 0.f;
    }
    __syncwarp();
    float xv[3] = {0.f, 0.f, 0.f};
    if (valid) load_point(x, p, xv);
#pragma unroll 2
    for (int l = 0; l < G.n_levels; ++l) {
      const float g0 = valid ? st[lane * kStagePitch + 2 * l + 0] : 0.f;
      const float g1 = valid ? st[lane * kStagePitch + 2 * l + 1] : 0.f;
      scatter_level(G, D.t[l], l, xv, g0, g1, lane);
    }
    __syncwarp();
  }
}

__global__ void hash_indices_kernel(const __grid_constant__ HashGridDev G, const float *__restrict__ x, int64_t P,
                                    int32_t *__restrict__ idx) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float xv[3];
  load_point(x, p, xv);
  for (int l = 0; l < G.n_levels; ++l) {
    Cell c;
    point_cell(G, l, xv, c);
    for (int k = 0; k < 8; ++k) idx[(p * G.n_levels + l) * 8 + k] = (int32_t)corner_index(G, c, k);
  }
}

__global__ void hash_coords_kernel(const int64_t *__restrict__ coords, int64_t n, int dim, uint32_t mask,
                                   int64_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t primes[3] = {1u, PN_PRIME_Y, PN_PRIME_Z};
  uint32_t h = 0;
  for (int d = 0; d < dim; ++d) h ^= (uint32_t)coords[i * dim + d] * primes[d];
  out[i] = (int64_t)(h & mask);
}

__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float *addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256)
hash_minmax_kernel(const __grid_constant__ HashGridDev G, const __grid_constant__ TablePtrs T, const float *__restrict__ x, int64_t P,
                   float *__restrict__ minmax) {
  const int lane = threadIdx.x & 31;
  float lo[PN_MAX_LEVELS], hi[PN_MAX_LEVELS];
#pragma unroll
  for (int l = 0; l < PN_MAX_LEVELS; ++l) { lo[l] = INFINITY; hi[l] = -INFINITY; }
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    float xv[3];
    load_point(x, p, xv);
#pragma unroll
    for (int l = 0; l < PN_MAX_LEVELS; ++l) {
      if (l < G.n_levels) {
        Cell c;
        point_cell(G, l, xv, c);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float2 e = __ldg(T.t[l] + corner_index(G, c, k));
          lo[l] = fminf(lo[l], fminf(e.x, e.y));
          hi[l] = fmaxf(hi[l], fmaxf(e.x, e.y));
        }
      }
    }
  }
#pragma unroll
  for (int l = 0; l < PN_MAX_LEVELS; ++l) {
    if (l < G.n_levels) {
      float a = lo[l], b = hi[l];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
      }
      if (lane == 0) {
        if (a != INFINITY) atomic_min_float(minmax + 2 * l + 0, a);
        if (b != -INFINITY) atomic_max_float(minmax + 2 * l + 1, b);
      }
    }
  }
}

int check_grid_args(const pn_hash_grid *g) {
  PN_REQUIRE(g != nullptr, PN_EINVAL, "grid is NULL");
  PN_REQUIRE(g->n_levels >= 1 && g->n_levels <= PN_MAX_LEVELS, PN_ESHAPE, "n_levels %d outside 1..%d",
             g->n_levels, PN_MAX_LEVELS);
  PN_REQUIRE(g->log2_hashmap_size >= 1 && g->log2_hashmap_size <= 30, PN_ESHAPE,
             "log2_hashmap_size %d outside 1..30", g->log2_hashmap_size);
  for (int l = 0; l < g->n_levels; ++l)
    PN_REQUIRE(g->resolution[l] >= 1.f && g->resolution[l] < 1e6f, PN_EINVAL, "resolution[%d]=%g", l,
               (double)g->resolution[l]);
  return 0;
}

// The same forward with the tables held as integer codes of the A-CAQ quantisers (eval form): each gathered
// entry is 2 or 4 bytes instead of 8, decoded as (code + qmin - zp) * scale — bit-identical to fake-quantising the
// fp32 entry (quantization.py:183-186), so feat equals the quantised reference's eval-mode output.  At T = 2^22
// the 512 MiB fp32 table set shrinks to 128 MiB (8-bit levels), which the 126 MB L2 nearly holds.
__global__ void __launch_bounds__(kHashThreads)
hash_fwd_packed_kernel(const __grid_constant__ HashGridDev G, const __grid_constant__ PackedDev T,
                       const float *__restrict__ x, int64_t P, float *__restrict__ feat, uint8_t *__restrict__ keep) {
  __shared__ float stage[kHashWarps][32 * kStagePitch];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int F = 2 * G.n_levels;
  float *st = stage[warp];
  for (int64_t base = ((int64_t)blockIdx.x * kHashWarps + warp) * 32; base < P;
       base += (int64_t)gridDim.x * kHashThreads) {
    const int64_t p = base + lane;
    const bool valid = p < P;
    float xv[3] = {0.f, 0.f, 0.f};
    if (valid) load_point(x, p, xv);
#pragma unroll 2
    for (int l = 0; l < G.n_levels; ++l) {
      Cell c;
      point_cell(G, l, xv, c);
      float e0[8], e1[8];
packed_gather8<true>(T, l, G, c, e0, e1);
      st[lane * kStagePitch + 2 * l + 0] = trilerp(e0, c.w);
      st[lane * kStagePitch + 2 * l + 1] = trilerp(e1, c.w);
    }
    if (valid && keep) keep[p] = point_keep(G, xv) ? 1 : 0;
    __syncwarp();
    const int64_t rows = (P - base) < 32 ? (P - base) : 32;
    for (int i = lane; i < 32 * F; i += 32) {
      const int row = i / F, col = i - row * F;
      if (row < rows) feat[(base + row) * F + col] = st[row * kStagePitch + col];
    }
    __syncwarp();
  }
}

int fill_packed(PackedDev &T, const pn_hash_grid *grid, const pn_packed_tables *packed) {
  PN_REQUIRE(packed != nullptr, PN_EINVAL, "packed is NULL");
  for (int l = 0; l < PN_MAX_LEVELS; ++l) {
    const int s = l < grid->n_levels ? l : 0;
    PN_REQUIRE(packed->codes[s] != nullptr, PN_EINVAL, "packed->codes[%d] is NULL", s);
    const int eb = packed->entry_bytes[s];
    PN_REQUIRE(eb == 2 || eb == 4 || eb == 8, PN_ESHAPE, "entry_bytes[%d] = %d (2, 4 or 8)", s, eb);
    PN_REQUIRE(((uintptr_t)packed->codes[s] & (uintptr_t)(eb - 1)) == 0, PN_EINVAL, "packed->codes[%d] misaligned", s);
    T.t[l] = packed->codes[s];
    T.eb[l] = (uint8_t)eb;
    T.scale[l] = packed->scale[s];
    T.sub[l] = packed_sub_const(packed->zero_point[s], packed->qmin[s]);
    if (eb != 8) {
      const float off = packed->qmin[s] - packed->zero_point[s];
      PN_REQUIRE(off == rintf(off) && fabsf(off) < 4194304.f, PN_EINVAL, "level %d: qmin - zero_point = %g is not a small integer",
                 s, (double)off);
    }
  }
  return 0;
}

// The per-level fake-quant applied to the TABLES instead of to every gathered corner: the quantiser is an elementwise
// function of the entry with one parameter row per level (hash_encoding.py:97-101), so quantising each of the L x T x 2
// entries once per call gives the gather bit-identical values at 1/250 of the quantiser evaluations (a 65536-ray pass
// gathers 2.1 G corner values from 16.8 M entries).  Levels whose quantiser is off are copied.
struct QuantOut {
  float4 *o[PN_MAX_LEVELS];
};

__global__ void __launch_bounds__(256)
table_fake_quant_kernel(const __grid_constant__ TablePtrs T, const __grid_constant__ QuantOut O,
                        const float *__restrict__ qparams, int64_t n4) {
  const int l = blockIdx.y;
  const float *q = qparams + l * PN_QROW;
  const bool on = q[5] != 0.f;
  const float scale = q[0], denom = q[1], zp = q[2], qmin = q[3], qmax = q[4];
  const bool train_form = q[6] != 0.f;
  const float4 *__restrict__ in = reinterpret_cast<const float4 *>(T.t[l]);
  float4 *__restrict__ out = O.o[l];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(in + i);
    if (on) {
      v.x = fake_quant(v.x, scale, denom, zp, qmin, qmax, train_form);
      v.y = fake_quant(v.y, scale, denom, zp, qmin, qmax, train_form);
      v.z = fake_quant(v.z, scale, denom, zp, qmin, qmax, train_form);
      v.w = fake_quant(v.w, scale, denom, zp, qmin, qmax, train_form);
    }
    out[i] = v;
  }
}

static int hash_blocks(int64_t P, int threads, int per_sm) {
  const int64_t need = ceil_div(P, threads);
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace pn

using namespace pn;

extern "C" int pn_hash_encode_fwd(const pn_hash_grid *grid, const float *const *tables, const float *qparams,
                                  const float *x, int64_t n_points, float *feat, uint8_t *keep,
                                  pn_stream_t stream) {
  if (int e = check_grid_args(grid)) return e;
  PN_REQUIRE(tables && x && feat, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(n_points >= 0, PN_EINVAL, "n_points < 0");
  if (n_points == 0) return 0;
  const HashGridDev G = make_grid_dev(*grid);
  TablePtrs T;
  for (int l = 0; l < PN_MAX_LEVELS; ++l)
    T.t[l] = reinterpret_cast<const float2 *>(tables[l < grid->n_levels ? l : 0]);
  for (int l = 0; l < grid->n_levels; ++l)
    PN_REQUIRE(tables[l] != nullptr && ((uintptr_t)tables[l] & 15) == 0, PN_EINVAL,
               "tables[%d] is NULL or not 16-byte aligned (x-adjacent corner rows are fetched as one 16-byte pair)", l);
  const int blocks = hash_blocks(n_points, kHashThreads, 16);
  if (qparams && G.pair_gather)
    hash_fwd_kernel<true, true><<<blocks, kHashThreads, 0, as_stream(stream)>>>(G, T, qparams, x, n_points, feat, keep);
  else if (qparams)
    hash_fwd_kernel<true, false><<<blocks, kHashThreads, 0, as_stream(stream)>>>(G, T, qparams, x, n_points, feat, keep);
  else if (G.pair_gather)
    hash_fwd_kernel<false, true><<<blocks, kHashThreads, 0, as_stream(stream)>>>(G, T, nullptr, x, n_points, feat, keep);
  else
    hash_fwd_kernel<false, false><<<blocks, kHashThreads, 0, as_stream(stream)>>>(G, T, nullptr, x, n_points, feat, keep);
  count_launch();
  return check_launch("hash_fwd_kernel");
}

extern "C" int pn_table_fake_quant(const float *const *tables, float *const *out, int n_levels, int64_t entries_per_level,
                                   const float *qparams, pn_stream_t stream) {
  PN_REQUIRE(tables && out && qparams, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(n_levels >= 1 && n_levels <= PN_MAX_LEVELS, PN_ESHAPE, "n_levels %d outside 1..%d", n_levels, PN_MAX_LEVELS);
  PN_REQUIRE(entries_per_level >= 2 && entries_per_level % 2 == 0, PN_ESHAPE, "entries_per_level must be even");
  TablePtrs T;
  QuantOut O;
  for (int l = 0; l < PN_MAX_LEVELS; ++l) {
    const int k = l < n_levels ? l : 0;
    PN_REQUIRE(tables[k] && out[k] && (((uintptr_t)tables[k] | (uintptr_t)out[k]) & 15) == 0, PN_EINVAL,
               "tables[%d] / out[%d] is NULL or not 16-byte aligned", k, k);
    T.t[l] = reinterpret_cast<const float2 *>(tables[k]);
    O.o[l] = reinterpret_cast<float4 *>(out[k]);
  }
  const int64_t n4 = entries_per_level / 2;
  const int64_t need = ceil_div(n4, 256);
  const int64_t cap = (int64_t)sm_count() * 2;
  dim3 grid((unsigned)(need < cap ? need : cap), (unsigned)n_levels);
  table_fake_quant_kernel<<<grid, 256, 0, as_stream(stream)>>>(T, O, qparams, n4);
  count_launch();
  return check_launch("table_fake_quant_kernel");
}

extern "C" int pn_hash_encode_fwd_packed(const pn_hash_grid *grid, const pn_packed_tables *packed, const float *x,
                                         int64_t n_points, float *feat, uint8_t *keep, pn_stream_t stream) {
  if (int e = check_grid_args(grid)) return e;
  PN_REQUIRE(x && feat, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(n_points >= 0, PN_EINVAL, "n_points < 0");
  PackedDev T;
  if (int e = fill_packed(T, grid, packed)) return e;
  if (n_points == 0) return 0;
  const HashGridDev G = make_grid_dev(*grid);
  const int blocks = hash_blocks(n_points, kHashThreads, 16);
  hash_fwd_packed_kernel<<<blocks, kHashThreads, 0, as_stream(stream)>>>(G, T, x, n_points, feat, keep);
  count_launch();
  return check_launch("hash_fwd_packed_kernel");
}

extern "C" int pn_hash_encode_bwd(const pn_hash_grid *grid, float *const *dtables, const float *x,
                                  const float *dfeat, int64_t n_points, pn_stream_t stream) {
  if (int e = check_grid_args(grid)) return e;
  PN_REQUIRE(dtables && x && dfeat, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(n_points >= 0, PN_EINVAL, "n_points < 0");
  if (n_points == 0) return 0;
  const HashGridDev G = make_grid_dev(*grid);
  GradPtrs D;
  for (int l = 0; l < PN_MAX_LEVELS; ++l) D.t[l] = reinterpret_cast<float2 *>(dtables[l < grid->n_levels ? l : 0]);
  for (int l = 0; l < grid->n_levels; ++l)
    PN_REQUIRE(dtables[l] != nullptr && ((uintptr_t)dtables[l] & 15) == 0, PN_EINVAL,
               "dtables[%d] is NULL or not 16-byte aligned", l);
  const int blocks = hash_blocks(n_points, kHashThreads, 16);
  hash_bwd_kernel<<<blocks, kHashThreads, 0, as_stream(stream)>>>(G, D, x, dfeat, n_points);
  count_launch();
  return check_launch("hash_bwd_kernel");
}

extern "C" int pn_hash_indices(const pn_hash_grid *grid, const float *x, int64_t n_points, int32_t *idx,
                               pn_stream_t stream) {
  if (int e = check_grid_args(grid)) return e;
  PN_REQUIRE(x && idx, PN_EINVAL, "NULL pointer argument");
  if (n_points <= 0) return 0;
  const HashGridDev G = make_grid_dev(*grid);
  hash_indices_kernel<<<(unsigned)ceil_div(n_points, 128), 128, 0, as_stream(stream)>>>(G, x, n_points, idx);
  count_launch();
  return check_launch("hash_indices_kernel");
}

extern "C" int pn_hash_coords(const int64_t *coords, int64_t n, int dim, int log2_hashmap_size, int64_t *out,
                              pn_stream_t stream) {
  PN_REQUIRE(coords && out, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(dim >= 1 && dim <= 3, PN_ESHAPE, "dim %d outside 1..3", dim);
  PN_REQUIRE(log2_hashmap_size >= 1 && log2_hashmap_size <= 30, PN_ESHAPE, "log2_hashmap_size %d",
             log2_hashmap_size);
  if (n <= 0) return 0;
  hash_coords_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
      coords, n, dim, (1u << log2_hashmap_size) - 1u, out);
  count_launch();
  return check_launch("hash_coords_kernel");
}

extern "C" int pn_hash_gather_minmax(const pn_hash_grid *grid, const float *const *tables, const float *x,
                                     int64_t n_points, float *minmax, pn_stream_t stream) {
  if (int e = check_grid_args(grid)) return e;
  PN_REQUIRE(tables && x && minmax, PN_EINVAL, "NULL pointer argument");
  if (n_points <= 0) return 0;
  const HashGridDev G = make_grid_dev(*grid);
  TablePtrs T;
  for (int l = 0; l < PN_MAX_LEVELS; ++l)
    T.t[l] = reinterpret_cast<const float2 *>(tables[l < grid->n_levels ? l : 0]);
  const int blocks = hash_blocks(n_points, 256, 4);
  hash_minmax_kernel<<<blocks, 256, 0, as_stream(stream)>>>(G, T, x, n_points, minmax);
  count_launch();
  return check_launch("hash_minmax_kernel");
}
