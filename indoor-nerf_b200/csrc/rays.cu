// K5 — ray generation and the small per-ray / per-sample elementwise pieces of render_rays, each a
// fixed sequence of individually rounded fp32 ops so that the sample positions (and therefore the
// hash indices downstream) are the reference's bit for bit.
#include "hash_core.cuh"
#include "ray_core.cuh"

namespace pn {


// get_rays (run_nerf_helpers.py:311-320): dirs = [(i-cx)/fx, -(j-cy)/fy, -1]; rays_d = sum_k dirs[k]*c2w[c][k]
__global__ void gen_rays_kernel(Cam cam, int H, int W, float *__restrict__ rays_o, float *__restrict__ rays_d) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (int64_t)H * W) return;
  const int j = (int)(p / W), i = (int)(p - (int64_t)j * W);
  float d[3];
  ray_dir(cam, i, j, d);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rays_d[3 * p + c] = d[c];
    rays_o[3 * p + c] = cam.t[c];
  }
}

// pts = o + d*z   (run_nerf.py:490,513)
__global__ void make_points_kernel(const float *__restrict__ o, int64_t os, const float *__restrict__ d,
                                   int64_t ds, const float *__restrict__ z, int64_t N, int S,
                                   float *__restrict__ pts) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N * S) return;
  const int64_t r = p / S;
  const float zz = z[p];
#pragma unroll
  for (int c = 0; c < 3; ++c) pts[3 * p + c] = point_at(o[r * os + c], d[r * ds + c], zz);
}

// run_nerf.py:466-488
__global__ void coarse_z_kernel(const float *__restrict__ near, const float *__restrict__ far, int64_t nfs,
                                const float *__restrict__ t_vals, const float *__restrict__ t_rand, int64_t N,
                                int S, int lindisp, float *__restrict__ z) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N * S) return;
  const int64_t r = p / S;
  const int s = (int)(p - r * S);
  z[p] = coarse_z_at(near[r * nfs], far[r * nfs], t_vals, s, S, lindisp != 0, t_rand ? &t_rand[p] : nullptr);
}

__global__ void ndc_rays_kernel(float cw, float ch, float near, float two_near, const float *__restrict__ o,
                                const float *__restrict__ d, int64_t n, float *__restrict__ oo,
                                float *__restrict__ od) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float ov[3] = {o[3 * p], o[3 * p + 1], o[3 * p + 2]}, dv[3] = {d[3 * p], d[3 * p + 1], d[3 * p + 2]};
  float a[3], b[3];
  ndc_ray(cw, ch, near, two_near, ov, dv, a, b);
#pragma unroll
  for (int c = 0; c < 3; ++c) { oo[3 * p + c] = a[c]; od[3 * p + c] = b[c]; }
}

__global__ void sh_encode_kernel(const float *__restrict__ dirs, int64_t n, float *__restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float o[16];
  sh4(dirs[3 * p], dirs[3 * p + 1], dirs[3 * p + 2], o);
  float4 *dst = reinterpret_cast<float4 *>(out + 16 * p);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}

}  // namespace pn

using namespace pn;

extern "C" int pn_gen_rays(int height, int width, const float *K, const float *c2w, float *rays_o, float *rays_d,
                           pn_stream_t stream) {
  PN_REQUIRE(K && c2w && rays_o && rays_d, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(height > 0 && width > 0, PN_EINVAL, "image %dx%d", height, width);
  Cam cam;
  cam.fx = K[0]; cam.cx = K[2]; cam.fy = K[4]; cam.cy = K[5];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) cam.R[r][c] = c2w[4 * r + c];
    cam.t[r] = c2w[4 * r + 3];
  }
  const int64_t n = (int64_t)height * width;
  gen_rays_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(cam, height, width, rays_o, rays_d);
  count_launch();
  return check_launch("gen_rays_kernel");
}

extern "C" int pn_make_points(const float *rays_o, int64_t o_stride, const float *rays_d, int64_t d_stride,
                              const float *z, int64_t n_rays, int n_samples, float *pts, pn_stream_t stream) {
  PN_REQUIRE(rays_o && rays_d && z && pts, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(n_samples >= 1, PN_EINVAL, "n_samples %d", n_samples);
  if (n_rays <= 0) return 0;
  const int64_t n = n_rays * n_samples;
  make_points_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(rays_o, o_stride, rays_d, d_stride,
                                                                                z, n_rays, n_samples, pts);
  count_launch();
  return check_launch("make_points_kernel");
}

extern "C" int pn_coarse_z(const float *near, const float *far, int64_t nf_stride, const float *t_vals,
                           const float *t_rand, int64_t n_rays, int n_samples, int lindisp, float *z,
                           pn_stream_t stream) {
  PN_REQUIRE(near && far && t_vals && z, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(n_samples >= 1, PN_EINVAL, "n_samples %d", n_samples);
  if (n_rays <= 0) return 0;
  const int64_t n = n_rays * n_samples;
  coarse_z_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(near, far, nf_stride, t_vals, t_rand,
                                                                             n_rays, n_samples, lindisp, z);
  count_launch();
  return check_launch("coarse_z_kernel");
}

extern "C" int pn_sh_encode(const float *dirs, int64_t n, float *out, pn_stream_t stream) {
  PN_REQUIRE(dirs && out, PN_EINVAL, "NULL pointer argument");
  if (n <= 0) return 0;
  sh_encode_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(dirs, n, out);
  count_launch();
  return check_launch("sh_encode_kernel");
}

extern "C" int pn_ndc_rays(int height, int width, double focal, double near, const float *rays_o, const float *rays_d,
                           int64_t n, float *out_o, float *out_d, pn_stream_t stream) {
  PN_REQUIRE(rays_o && rays_d && out_o && out_d, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(height > 0 && width > 0 && focal != 0.0, PN_EINVAL, "bad camera");
  if (n <= 0) return 0;
  const float cw = (float)(-1.0 / (width / (2.0 * focal))), ch = (float)(-1.0 / (height / (2.0 * focal)));
  ndc_rays_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(cw, ch, (float)near, (float)(2.0 * near),
                                                                             rays_o, rays_d, n, out_o, out_d);
  count_launch();
  return check_launch("ndc_rays_kernel");
}
