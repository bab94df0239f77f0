// K6 — the data formats either side of the hot path:
//   * training batches generated on the fly from (image, pixel) ids instead of the precomputed
//     [N*H*W, 3, 3] rays_rgb tensor (run_nerf.py:896-920, 962-973) or the full-image get_rays + gather of
//     the no_batching branch (run_nerf.py:976-1004);
//   * test-set evaluation without host round trips: squared error (PSNR), SSIM and to8b on the device
//     (run_nerf.py:154-215, evaluation_utils.py:22-74);
//   * export of hash tables / weights as integer codes bit-packed at the learned A-CAQ width
//     (quantization.py:126-187 only ever fake-quantises; run_nerf.py:1345-1362 saves fp32).
// All of it is HBM-bound byte work: one pass, coalesced, no tensor cores.
#include <algorithm>

#include "io_core.cuh"

namespace pn {

// ---- ray bank ---------------------------------------------------------------------------------------------------
// ids[b] = (slot * H + j) * W + i  with slot indexing the TRAINING images in i_train order (the row order of the
// reference's rays_rgb before shuffling).  image_index[slot] = row of that image in poses / images.
template <bool F64, typename PIX>
__global__ void ray_bank_kernel(const int64_t *__restrict__ ids, int64_t B, int H, int W, CamF64 cam,
                                const float *__restrict__ poses, int64_t pose_stride,
                                const int32_t *__restrict__ image_index, const PIX *__restrict__ images,
                                float *__restrict__ rays, float *__restrict__ target) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t id = ids[b];
  const int64_t hw = (int64_t)H * W;
  const int64_t slot = id / hw, pix = id - slot * hw;
  const int j = (int)(pix / W), i = (int)(pix - (int64_t)j * W);
  const int64_t img = image_index ? image_index[slot] : slot;
  const float *c2w = poses + img * pose_stride;          // row-major [3,4] (or the top of a [4,4])
  float d[3];
  if (F64) ray_dir_f64(cam, c2w, 4, i, j, d);
  else ray_dir_f32(cam, c2w, 4, i, j, d);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rays[3 * b + c] = c2w[4 * c + 3];
    rays[3 * (B + b) + c] = d[c];
  }
  if (target) {
    const PIX *px = images + (img * hw + pix) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (sizeof(PIX) == 1) target[3 * b + c] = (float)pn_ddiv((double)px[c], 255.0);   // load_blender.py:62
      else target[3 * b + c] = (float)px[c];
    }
  }
}

// ---- evaluation ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void block_atomic_add(double v, double *out) {
  __shared__ double part[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(lane < nw ? part[lane] : 0.0);
    if (lane == 0) atomicAdd(out, v);
  }
}

// sum += (a - b)^2, each difference and square rounded in fp32 like np.square(rgb - gt) (run_nerf.py:186),
// accumulated in fp64.
__global__ void sqerr_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t n, double *sum) {
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const float df = pn_sub(a[k], b[k]);
    acc += (double)pn_mul(df, df);
  }
  block_atomic_add(acc, sum);
}

__global__ void to8b_kernel(const float *__restrict__ x, int64_t n, uint8_t *__restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = to8b_one(x[k]);
}

// SSIM of two [H,W,C] images: sum over the interior pixels (pad = win/2 on each side, the crop scikit-image
// applies before averaging) and channels of S.  Block = 32x8 output pixels of one channel; the (32+win-1) x
// (8+win-1) input window of both images is staged in shared memory.
constexpr int SSIM_WIN = 7, SSIM_TX = 32, SSIM_TY = 8;
__global__ void ssim_kernel(const float *__restrict__ a, const float *__restrict__ b, int H, int W, int C,
                            double data_range, double *sum) {
  constexpr int SW = SSIM_TX + SSIM_WIN - 1, SH = SSIM_TY + SSIM_WIN - 1;
  __shared__ float sa[SH][SW], sb[SH][SW];
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * SSIM_TX, y0 = blockIdx.y * SSIM_TY;      // top-left input pixel of the tile
  const int tid = threadIdx.y * SSIM_TX + threadIdx.x;
  for (int k = tid; k < SW * SH; k += SSIM_TX * SSIM_TY) {
    const int yy = k / SW, xx = k - yy * SW;
    const int gy = y0 + yy, gx = x0 + xx;
    const bool in = gy < H && gx < W;
    const int64_t off = ((int64_t)gy * W + gx) * C + c;
    sa[yy][xx] = in ? a[off] : 0.f;
    sb[yy][xx] = in ? b[off] : 0.f;
  }
  __syncthreads();
  double s = 0.0;
  const int ox = x0 + threadIdx.x, oy = y0 + threadIdx.y;              // window origin = output pixel - pad
  if (ox + SSIM_WIN <= W && oy + SSIM_WIN <= H) {
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
    for (int dy = 0; dy < SSIM_WIN; ++dy)
#pragma unroll
      for (int dx = 0; dx < SSIM_WIN; ++dx) {
        const double u = sa[threadIdx.y + dy][threadIdx.x + dx], v = sb[threadIdx.y + dy][threadIdx.x + dx];
        sx += u; sy += v; sxx += u * u; syy += v * v; sxy += u * v;
      }
    s = ssim_from_sums(sx, sy, sxx, syy, sxy, SSIM_WIN * SSIM_WIN, data_range);
  }
  // block reduction (256 threads, 2-D block flattened)
  __shared__ double part[8];
  s = warp_sum(s);
  if ((tid & 31) == 0) part[tid >> 5] = s;
  __syncthreads();
  if (tid < 32) {
    s = warp_sum(tid < 8 ? part[tid] : 0.0);
    if (tid == 0) atomicAdd(sum, s);
  }
}

// ---- A-CAQ export --------------------------------------------------------------------------------------------------
// thread = 32 consecutive values -> `bits` words
__global__ void quant_pack_kernel(const float *__restrict__ x, int64_t groups, const float *__restrict__ qrow, int bits,
                                  uint32_t *__restrict__ words) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  const float denom = qrow[1], zp = qrow[2], qmin = qrow[3], qmax = qrow[4];
  uint32_t code[32];
  const float4 *src = reinterpret_cast<const float4 *>(x + g * 32);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 v = src[q];
    code[4 * q] = quant_code(v.x, denom, zp, qmin, qmax);
    code[4 * q + 1] = quant_code(v.y, denom, zp, qmin, qmax);
    code[4 * q + 2] = quant_code(v.z, denom, zp, qmin, qmax);
    code[4 * q + 3] = quant_code(v.w, denom, zp, qmin, qmax);
  }
  uint32_t w[32];
  pack32(code, bits, w);
  for (int k = 0; k < bits; ++k) words[g * bits + k] = w[k];
}

// thread = one value
__global__ void quant_unpack_kernel(const uint32_t *__restrict__ words, int64_t n, const float *__restrict__ qrow,
                                    int bits, float *__restrict__ x) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float scale = qrow[0], zp = qrow[2], qmin = qrow[3];
  const int64_t g = k >> 5;
  const uint32_t code = unpack_one(words + g * bits, (int)(k & 31), bits);
  x[k] = quant_value(code, scale, zp, qmin);
}

// thread = one value -> one u8 / u16 container code (tables resident as codes for inference)
template <typename CODE>
__global__ void quant_codes_kernel(const float *__restrict__ x, int64_t n, const float *__restrict__ qrow,
                                   CODE *__restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = (CODE)quant_code(x[k], qrow[1], qrow[2], qrow[3], qrow[4]);
}

template <typename CODE>
__global__ void quant_unpack_codes_kernel(const uint32_t *__restrict__ words, int64_t n, int bits, CODE *__restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = (CODE)unpack_one(words + (k >> 5) * bits, (int)(k & 31), bits);
}

}  // namespace pn

using namespace pn;

extern "C" int pn_ray_bank_batch(const int64_t *ids, int64_t n_rays, int height, int width, const double *K,
                                 const float *poses, int64_t pose_stride, const int32_t *image_index,
                                 const void *images, int image_dtype, int f64_dirs, float *batch_rays, float *target,
                                 pn_stream_t stream) {
  PN_REQUIRE(ids && K && poses && batch_rays, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(height > 0 && width > 0 && K[0] != 0.0 && K[4] != 0.0, PN_EINVAL, "bad camera");
  PN_REQUIRE(pose_stride >= 12, PN_EINVAL, "pose_stride %lld < 12", (long long)pose_stride);
  PN_REQUIRE((target == nullptr) == (images == nullptr), PN_EINVAL, "images and target go together");
  PN_REQUIRE(image_dtype == 0 || image_dtype == 1, PN_EINVAL, "image_dtype %d (0 = fp32, 1 = u8)", image_dtype);
  if (n_rays <= 0) return 0;
  CamF64 cam;
  cam.fx = K[0]; cam.cx = K[2]; cam.fy = K[4]; cam.cy = K[5];
  const unsigned grid = (unsigned)ceil_div(n_rays, 256);
  cudaStream_t s = as_stream(stream);
#define PN_BANK(F64, PIX)                                                                                     \
  ray_bank_kernel<F64, PIX><<<grid, 256, 0, s>>>(ids, n_rays, height, width, cam, poses, pose_stride, image_index, \
                                                 (const PIX *)images, batch_rays, target)
  if (f64_dirs) {
    if (image_dtype == 1) PN_BANK(true, uint8_t); else PN_BANK(true, float);
  } else {
    if (image_dtype == 1) PN_BANK(false, uint8_t); else PN_BANK(false, float);
  }
#undef PN_BANK
  count_launch();
  return check_launch("ray_bank_kernel");
}

extern "C" int pn_image_sqerr(const float *a, const float *b, int64_t n, double *sum, pn_stream_t stream) {
  PN_REQUIRE(a && b && sum, PN_EINVAL, "NULL pointer argument");
  if (n <= 0) return 0;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256 * 4), (int64_t)sm_count() * 8);
  sqerr_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, n, sum);
  count_launch();
  return check_launch("sqerr_kernel");
}

extern "C" int pn_image_ssim(const float *a, const float *b, int height, int width, int channels, double data_range,
                             double *sum, pn_stream_t stream) {
  PN_REQUIRE(a && b && sum, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(height >= SSIM_WIN && width >= SSIM_WIN, PN_ESHAPE, "image %dx%d smaller than the %d-pixel window", height,
             width, SSIM_WIN);
  PN_REQUIRE(channels >= 1 && channels <= 65535, PN_ESHAPE, "channels %d", channels);
  dim3 grid((unsigned)ceil_div(width - SSIM_WIN + 1, SSIM_TX), (unsigned)ceil_div(height - SSIM_WIN + 1, SSIM_TY),
            (unsigned)channels);
  PN_REQUIRE(grid.y <= 65535, PN_ESHAPE, "image too tall (%d rows)", height);
  ssim_kernel<<<grid, dim3(SSIM_TX, SSIM_TY), 0, as_stream(stream)>>>(a, b, height, width, channels, data_range, sum);
  count_launch();
  return check_launch("ssim_kernel");
}

extern "C" int pn_to8b(const float *x, int64_t n, uint8_t *out, pn_stream_t stream) {
  PN_REQUIRE(x && out, PN_EINVAL, "NULL pointer argument");
  if (n <= 0) return 0;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16);
  to8b_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, n, out);
  count_launch();
  return check_launch("to8b_kernel");
}

extern "C" int pn_quant_pack(const float *x, int64_t n, const float *qrow, int bits, uint32_t *words,
                             pn_stream_t stream) {
  PN_REQUIRE(x && qrow && words, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(bits >= 1 && bits <= 24, PN_ESHAPE, "bits %d outside 1..24 (wider levels are stored as fp32)", bits);
  PN_REQUIRE(n >= 0 && n % 32 == 0, PN_ESHAPE, "n %lld is not a multiple of 32", (long long)n);
  PN_REQUIRE(((uintptr_t)x & 15) == 0, PN_EINVAL, "x must be 16-byte aligned");
  if (n == 0) return 0;
  const int64_t groups = n / 32;
  quant_pack_kernel<<<(unsigned)ceil_div(groups, 128), 128, 0, as_stream(stream)>>>(x, groups, qrow, bits, words);
  count_launch();
  return check_launch("quant_pack_kernel");
}

extern "C" int pn_quant_unpack(const uint32_t *words, int64_t n, const float *qrow, int bits, float *x,
                               pn_stream_t stream) {
  PN_REQUIRE(x && qrow && words, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(bits >= 1 && bits <= 24, PN_ESHAPE, "bits %d outside 1..24", bits);
  PN_REQUIRE(n >= 0 && n % 32 == 0, PN_ESHAPE, "n %lld is not a multiple of 32", (long long)n);
  if (n == 0) return 0;
  quant_unpack_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(words, n, qrow, bits, x);
  count_launch();
  return check_launch("quant_unpack_kernel");
}

extern "C" int pn_quant_codes(const float *x, int64_t n, const float *qrow, int code_bytes, void *out,
                              pn_stream_t stream) {
  PN_REQUIRE(x && qrow && out, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(code_bytes == 1 || code_bytes == 2, PN_ESHAPE, "code_bytes %d (1 or 2)", code_bytes);
  if (n <= 0) return 0;
  const unsigned grid = (unsigned)ceil_div(n, 256);
  if (code_bytes == 1) quant_codes_kernel<uint8_t><<<grid, 256, 0, as_stream(stream)>>>(x, n, qrow, (uint8_t *)out);
  else quant_codes_kernel<uint16_t><<<grid, 256, 0, as_stream(stream)>>>(x, n, qrow, (uint16_t *)out);
  count_launch();
  return check_launch("quant_codes_kernel");
}

extern "C" int pn_quant_unpack_codes(const uint32_t *words, int64_t n, int bits, int code_bytes, void *out,
                                     pn_stream_t stream) {
  PN_REQUIRE(words && out, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(code_bytes == 1 || code_bytes == 2, PN_ESHAPE, "code_bytes %d (1 or 2)", code_bytes);
  PN_REQUIRE(bits >= 1 && bits <= 8 * code_bytes, PN_ESHAPE, "bits %d do not fit %d-byte codes", bits, code_bytes);
  PN_REQUIRE(n >= 0 && n % 32 == 0, PN_ESHAPE, "n %lld is not a multiple of 32", (long long)n);
  if (n == 0) return 0;
  const unsigned grid = (unsigned)ceil_div(n, 256);
  if (code_bytes == 1) quant_unpack_codes_kernel<uint8_t><<<grid, 256, 0, as_stream(stream)>>>(words, n, bits, (uint8_t *)out);
  else quant_unpack_codes_kernel<uint16_t><<<grid, 256, 0, as_stream(stream)>>>(words, n, bits, (uint16_t *)out);
  count_launch();
  return check_launch("quant_unpack_codes_kernel");
}
