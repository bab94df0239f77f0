/*
 * pocketnerf.h — C ABI of libpocketnerf.so: the B200 (sm_100a) implementation of PocketNeRF's
 * HashNeRF hot path  HashEmbedder -> SHEncoder -> NeRFSmall -> raw2outputs -> sample_pdf.
 *
 * The reference (ryanjsuh/indoor-nerf, PocketNeRF/) is pure Python on PyTorch and has no FFI of
 * its own; the "operator API" a replacement must serve is the set of Python callables named in
 * SURVEY.md §8b.  Each entry point below states the reference lines whose arithmetic it replaces
 * (paths relative to PocketNeRF/).  The Python host layer (indoor-nerf_b200/) binds these with
 * ctypes and re-exports the reference's names; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every pointer marked "device" is caller-owned device memory on the CURRENT device; the library
 *    never allocates, frees or synchronises (exceptions are noted); pointers marked "host" are read
 *    during the call only.
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*).
 *  - return 0 on success, a negative PN_E* code otherwise; pn_last_error() (thread-local) says why.
 *  - floats are IEEE fp32; where the reference's result is an integer (hash indices, sample bins)
 *    or a fixed sequence of individually rounded fp32 ops (voxel weights, trilinear interpolation,
 *    SH, fake-quant, o+d*z), the kernels reproduce it bit for bit (no FMA contraction there).
 *  - no entry point has a CPU implementation: without a CUDA device they fail with PN_ECUDA.
 */
#ifndef POCKETNERF_H
#define POCKETNERF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN_ABI_VERSION 11
#define PN_MAX_LEVELS 16

#define PN_EINVAL (-1)   /* bad argument */
#define PN_ECUDA  (-2)   /* CUDA runtime error (launch failure, no device, ...) */
#define PN_ESHAPE (-3)   /* shape outside what the kernels are built for */

typedef void *pn_stream_t; /* cudaStream_t */

int pn_abi_version(void);
const char *pn_last_error(void);
/* Number of kernels this library has launched since load (all threads); bench.py's gpu_launches. */
int64_t pn_launch_count(void);

/* ---- hash grid ---------------------------------------------------------------------------- */

/* Geometry of the multiresolution hash grid; replaces the per-call tensor arithmetic of
 * HashEmbedder.forward (hash_encoding.py:82-107) and get_voxel_vertices (utils.py:95-117).
 * `resolution[l]` must be the value of torch.floor(base_resolution * b**l) evaluated with the
 * reference's own float32 expression (hash_encoding.py:28,89) — the host layer does that. */
typedef struct {
  float box_min[3];
  float box_max[3];
  float resolution[PN_MAX_LEVELS];
  int32_t n_levels;          /* 1..16 */
  int32_t log2_hashmap_size; /* 1..30; table l has 2^log2_hashmap_size rows of 2 floats */
} pn_hash_grid;

/* Per-level fake-quant of the gathered corner embeddings (hash_encoding.py:97-101 calling
 * LearnedBitwidthQuantizer.forward, quantization.py:144-187).  One row of 8 floats per level in
 * DEVICE memory so that no host sync is needed when soft_bits move:
 *   [0] scale  [1] scale+1e-8 (the divisor)  [2] zero_point  [3] qmin  [4] qmax
 *   [5] enabled (0/1)  [6] mode: 1 = training form x+(dq-x), 0 = eval form dq  [7] unused */
#define PN_QROW 8

/* feat[P, 2*n_levels], keep[P] (1 = inside the box on all three axes) from points x[P,3].
 *   tables: host array of n_levels device pointers (the reference's nn.Embedding weights [T,2]).
 *   qparams: device [n_levels][PN_QROW] or NULL.   Reference: hash_encoding.py:82-107. */
int pn_hash_encode_fwd(const pn_hash_grid *grid, const float *const *tables, const float *qparams,
                       const float *x, int64_t n_points, float *feat, uint8_t *keep,
                       pn_stream_t stream);

/* dtables[l][h] += sum over points/corners of the trilinear weight times dfeat — the dense
 * gradient aten::embedding_dense_backward gives the reference (hash_encoding.py:94 under
 * autograd).  Accumulates (atomics) into caller-zeroed dtables; the fake-quant is a
 * straight-through estimator, so it does not appear here. */
int pn_hash_encode_bwd(const pn_hash_grid *grid, float *const *dtables, const float *x,
                       const float *dfeat, int64_t n_points, pn_stream_t stream);

/* The 8 hashed corner indices per level, idx[P, n_levels, 8] (int32) in the reference's corner
 * order (utils.py:9, 13-24, 114-115).  Test/diagnostic entry for the bit-exact index claim. */
int pn_hash_indices(const pn_hash_grid *grid, const float *x, int64_t n_points, int32_t *idx,
                    pn_stream_t stream);

/* utils.hash (utils.py:13-24) on int64 coordinates coords[n, dim], dim <= 3 — used by the TV loss
 * (loss.py:29).  out[n] int64. */
int pn_hash_coords(const int64_t *coords, int64_t n, int dim, int log2_hashmap_size, int64_t *out,
                   pn_stream_t stream);

/* out[l][i] = LearnedBitwidthQuantizer.forward(tables[l][i]) with level l's row of qparams (quantization.py:177-187), or
 * a copy where the row is disabled: the fake-quant of hash_encoding.py:97-101 hoisted from the gathered corner values to
 * the table entries.  Bit-identical to passing `qparams` to pn_hash_encode_fwd / pn_field_fwd_bf16 on the original tables
 * (in the exact IEEE-division form for both arithmetic modes), at L x T x 2 quantiser evaluations per call instead of
 * 16 x 8 x 2 per point.  tables[l], out[l]: device [entries_per_level, 2] fp32, 16-byte aligned; may not alias. */
int pn_table_fake_quant(const float *const *tables, float *const *out, int n_levels, int64_t entries_per_level,
                        const float *qparams, pn_stream_t stream);

/* Per-level min and max of the gathered corner values, minmax[n_levels][2] (caller initialises
 * to +inf/-inf) — the batch statistics LearnedBitwidthQuantizer.calibrate reads
 * (quantization.py:97-119) on the first quantised training call. */
int pn_hash_gather_minmax(const pn_hash_grid *grid, const float *const *tables, const float *x,
                          int64_t n_points, float *minmax, pn_stream_t stream);

/* ---- view-direction encoding ---------------------------------------------------------------- */

/* out[n,16] = degree-4 real spherical harmonics of dirs[n,3] (hash_encoding.py:153-191). */
int pn_sh_encode(const float *dirs, int64_t n, float *out, pn_stream_t stream);

/* ---- NeRFSmall ---------------------------------------------------------------------------------- */

/* Weights in the reference's nn.Linear layout [out,in], row-major (run_nerf_helpers.py:187-263 with
 * the create_nerf shapes, run_nerf.py:240-247): s0[64,32] s1[16,64] c0[64,31] c1[64,64] c2[3,64];
 * optional normal head n0w[32,15] n0b[32] n2w[3,32] n2b[3] (all four NULL when absent). */
typedef struct {
  const float *s0, *s1, *c0, *c1, *c2;
  const float *n0w, *n0b, *n2w, *n2b;
} pn_mlp_weights;

typedef struct {
  float *s0, *s1, *c0, *c1, *c2;
  float *n0w, *n0b, *n2w, *n2b;
} pn_mlp_grads;

/* Inputs of one NeRFSmall evaluation over n_points rows.
 *   feat   device [n_points, 32], row stride feat_stride floats.
 *   view   either sh (device [n_points,16], row stride sh_stride) or dirs (device [n_rays,3]) with
 *          samples_per_ray: point p uses dirs[p / samples_per_ray] and the SH basis is evaluated in
 *          the kernel (run_nerf.py:59-63 + hash_encoding.py:153-191).  Exactly one is non-NULL.
 *   act_q  device [PN_QROW] fake-quant of the hidden activation after the first ReLU
 *          (run_nerf_helpers.py:281-284) or NULL.
 *   keep   device [n_points] or NULL; where 0 the LAST output channel is set to 0 and receives no
 *          gradient (run_nerf.py:66 — sigma for 4 channels, normal_z for 7, as the reference does). */
typedef struct {
  const float *feat;
  int64_t feat_stride;
  const float *sh;
  int64_t sh_stride;
  const float *dirs;
  int32_t samples_per_ray;
  const float *act_q;
  const uint8_t *keep;
  int64_t n_points;
} pn_mlp_input;

/* out[n_points, 4] = (r,g,b,sigma) or [n_points,7] with the normalised normal appended when the
 * normal head is present (run_nerf_helpers.py:265-306). */
int pn_mlp_fwd(const pn_mlp_weights *w, const pn_mlp_input *in, float *out, pn_stream_t stream);

/* Backward of pn_mlp_fwd for the cotangent dout[n_points, 4|7]:
 *   dfeat [n_points,32] (stride dfeat_stride) written; dsh [n_points,16] written when non-NULL (only
 *   valid with in->sh); dw accumulated (atomics) into caller-zeroed buffers. */
int pn_mlp_bwd(const pn_mlp_weights *w, const pn_mlp_input *in, const float *dout, float *dfeat,
               int64_t dfeat_stride, float *dsh, int64_t dsh_stride, const pn_mlp_grads *dw,
               pn_stream_t stream);

/* The same two operations in the bf16 tensor-core mode (tcgen05.mma with TMEM accumulators: inputs,
 * weights and hidden activations rounded to bf16, fp32 accumulation, fp32 outputs and gradients).  Same
 * arguments, same semantics; results agree with the fp32 mode to ~2e-3 relative (BASELINE north_star). */
int pn_mlp_fwd_bf16(const pn_mlp_weights *w, const pn_mlp_input *in, float *out, pn_stream_t stream);
int pn_mlp_bwd_bf16(const pn_mlp_weights *w, const pn_mlp_input *in, const float *dout, float *dfeat,
                    int64_t dfeat_stride, float *dsh, int64_t dsh_stride, const pn_mlp_grads *dw,
                    pn_stream_t stream);

/* run_network (run_nerf.py:53-68) as ONE kernel in the bf16 tensor-core mode: hash-grid encode (16 levels) ->
 * SH of the ray's view direction -> NeRFSmall -> keep mask, for pts[n_points,3] with point p belonging to ray
 * p / samples_per_ray.  The [P,32] feature tensor never reaches HBM in fp32: the kernel computes each tile's
 * features straight into the MMA operand tile in shared memory and (when feat_tiles != NULL) saves that tile —
 * bf16, 8 KB per 128 points, already in operand layout — for the backward.
 *   out [n_points, 4|7], keep [n_points] (may be NULL), feat_tiles ceil(n_points/128)*8192 bytes (may be NULL). */
int pn_field_fwd_bf16(const pn_hash_grid *grid, const float *const *tables, const float *qparams,
                      const pn_mlp_weights *w, const float *pts, const float *dirs, int samples_per_ray,
                      const float *act_q, int64_t n_points, float *out, uint8_t *keep, void *feat_tiles,
                      pn_stream_t stream);
/* Backward of pn_field_fwd_bf16 as ONE kernel: NeRFSmall backward on the saved feature tiles and, in the same launch,
 * the run-aggregated scatter-add of the feature gradient into dtables (caller-zeroed, accumulated, 16-byte aligned);
 * weight gradients accumulated into dw.  The [P,32] feature gradient never reaches HBM as a tensor: the kernel's MLP
 * warps hand each 128-point tile of it to the kernel's scatter warps through a small per-CTA ring in `workspace`
 * (caller-owned scratch of pn_field_bwd_workspace_bytes() bytes, 16-byte aligned, contents undefined before and
 * after; it stays L2-resident).  dout must be 16-byte aligned when the network has 4 output channels. */
int pn_field_bwd_bf16(const pn_hash_grid *grid, float *const *dtables, const pn_mlp_weights *w,
                      const void *feat_tiles, const float *pts, const float *dirs, int samples_per_ray,
                      const float *act_q, const uint8_t *keep, const float *dout, int64_t n_points,
                      const pn_mlp_grads *dw, void *workspace, int64_t workspace_bytes, pn_stream_t stream);
/* Scratch size pn_field_bwd_bf16 needs on the current device (depends on its SM count only). */
int64_t pn_field_bwd_workspace_bytes(void);

/* Diagnostic: one tcgen05 GEMM with caller-chosen shared-memory descriptor fields (cfg[15], see
 * mlp_tc.cu) — proves the K-major / MN-major operand readings and the TMEM accumulator layouts on the
 * hardware.  D is the raw [128 lanes][N] accumulator. */
int pn_tc_selftest(const int32_t *cfg, const float *A, const float *B, float *D, pn_stream_t stream);

/* Diagnostic: install (buf != NULL) or remove a device buffer of 3 * cap_per_thread int64 into which CTA 0 of the
 * single-role fused backward (PN_FIELD_BWD=v1) writes clock64() marks at the boundaries of its MMA -> epilogue rounds
 * (thread 0 = the MMA issuer in [0, cap), thread 160 in [cap, 2 cap); the third row is spare): scripts/timeline_rounds.py
 * turns them into a per-round latency table.  Never set in production; the kernel tests one pointer per mark. */
int pn_debug_timeline(int64_t *buf, int64_t cap_per_thread);

/* ---- volume rendering ---------------------------------------------------------------------------- */

/* raw2outputs (run_nerf.py:347-411).  raw[N,S,C] C = 4 or 7, z[N,S], rays_d[N,3], noise[N,S] or NULL
 * (already multiplied by raw_noise_std).  Outputs: rgb[N,3] disp[N] acc[N] weights[N,S] depth[N]
 * sparsity[N] normal[N,3] (normal only when C == 7 and non-NULL). S <= 512. */
int pn_composite_fwd(const float *raw, int channels, const float *z, const float *rays_d,
                     const float *noise, int64_t n_rays, int n_samples, int white_bkgd, float *rgb,
                     float *disp, float *acc, float *weights, float *depth, float *sparsity,
                     float *normal, pn_stream_t stream);

/* Backward of pn_composite_fwd.  Any cotangent pointer may be NULL (treated as zeros).
 * draw[N,S,C] is written (not accumulated). */
int pn_composite_bwd(const float *raw, int channels, const float *z, const float *rays_d,
                     const float *noise, int64_t n_rays, int n_samples, int white_bkgd,
                     const float *d_rgb, const float *d_disp, const float *d_acc,
                     const float *d_weights, const float *d_depth, const float *d_sparsity,
                     const float *d_normal, float *draw, pn_stream_t stream);

/* sample_pdf (run_nerf_helpers.py:354-397).  bins[N,nb], weights: nb-1 values per ray starting at
 * weights + r*w_stride; u: row r at u + r*u_stride (u_stride = 0 broadcasts one row, which is how
 * the host passes torch.linspace(0,1,M) for det=True).  samples[N,M]; inds[N,M] (int32, the
 * searchsorted(right=True) result) and cdf[N,nb] are optional outputs.  nb <= 512. */
int pn_sample_pdf(const float *bins, const float *weights, int64_t w_stride, const float *u,
                  int64_t u_stride, int64_t n_rays, int nb, int n_samples, float *samples,
                  int32_t *inds, float *cdf, pn_stream_t stream);

/* The inversion step alone, from a caller-supplied cdf[N,nb] (run_nerf_helpers.py:381-397): the
 * entry whose bin indices are bit-exact by construction. */
int pn_sample_from_cdf(const float *cdf, const float *bins, const float *u, int64_t u_stride,
                       int64_t n_rays, int nb, int n_samples, float *samples, int32_t *inds,
                       pn_stream_t stream);

/* out[N, sa+sb] = ascending sort of the concatenation of a[N,sa] and b[N,sb]
 * (torch.sort(torch.cat(...)) at run_nerf.py:512).  sa+sb <= 512. */
int pn_sort_merge(const float *a, int sa, const float *b, int sb, int64_t n_rays, float *out,
                  pn_stream_t stream);

/* ---- rays ------------------------------------------------------------------------------------------ */

/* get_rays (run_nerf_helpers.py:311-320): K host[9] row-major intrinsics, c2w host[12] row-major
 * [3,4]; rays_o, rays_d device [H,W,3]. */
int pn_gen_rays(int height, int width, const float *K, const float *c2w, float *rays_o,
                float *rays_d, pn_stream_t stream);

/* ndc_rays (run_nerf_helpers.py:333-350) on n rays: rays_o/rays_d device [n,3] -> out_o/out_d [n,3].
 * focal and near are host scalars (python floats in the reference). */
int pn_ndc_rays(int height, int width, double focal, double near, const float *rays_o,
                const float *rays_d, int64_t n, float *out_o, float *out_d, pn_stream_t stream);

/* pts[N,S,3] = rays_o[N,3] + rays_d[N,3] * z[N,S]  (run_nerf.py:490,513), mul then add, each
 * rounded.  o_stride/d_stride are the row strides (in floats) of rays_o / rays_d. */
int pn_make_points(const float *rays_o, int64_t o_stride, const float *rays_d, int64_t d_stride,
                   const float *z, int64_t n_rays, int n_samples, float *pts, pn_stream_t stream);

/* Stratified coarse depths (run_nerf.py:466-488): z[N,S] from near[N], far[N] (row stride
 * nf_stride floats each), t_vals[S] (the host passes torch.linspace(0,1,S) so that its rounding
 * is the reference's), optional t_rand[N,S] (NULL = no perturbation), lindisp flag. */
int pn_coarse_z(const float *near, const float *far, int64_t nf_stride, const float *t_vals,
                const float *t_rand, int64_t n_rays, int n_samples, int lindisp, float *z,
                pn_stream_t stream);

/* ---- train-step surroundings ------------------------------------------------------ */

/* total_variation_loss (loss.py:11-43) for all levels in one launch.  cube: host[n_levels] cube sizes;
 * min_vertex: device int64 [n_levels,3] (the torch.randint draws of loss.py:26); loss: device [n_levels],
 * caller-zeroed, receives (tv_x+tv_y+tv_z)/cube_size per level. */
int pn_tv_loss_fwd(const float *const *tables, int n_levels, int log2_hashmap_size, const int32_t *cube,
                   const int64_t *min_vertex, float *loss, pn_stream_t stream);
/* dtables[l] += dloss[l] * d loss_l / d table_l  (atomics into caller-owned dense buffers). */
int pn_tv_loss_bwd(const float *const *tables, float *const *dtables, int n_levels, int log2_hashmap_size,
                   const int32_t *cube, const int64_t *min_vertex, const float *dloss, pn_stream_t stream);

/* One RAdam update (radam.py:55-88) over n contiguous fp32 elements: moments always; mode 2 = rectified
 * adaptive step p += -wd*lr*p; p += -step_size*lr * m/(sqrt(v)+eps); mode 1 = p += -step_size*lr*m
 * (degenerated_to_sgd); mode 0 = moments only (N_sma < 5).  The scalars are computed by the host exactly as
 * radam.py:63-78 does. */
int pn_radam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float beta1,
                  float beta2, float eps, float weight_decay_times_lr, float step_size_times_lr, int mode,
                  pn_stream_t stream);
/* The same update with the two per-step scalars read from DEVICE memory: dyn[0] = weight_decay*lr, dyn[1] =
 * step_size*lr.  For a training step recorded in a CUDA graph (Trainer(cuda_graph=True)): the launch is recorded once,
 * pn_store_floats refreshes `dyn` before every replay, so the replayed update follows the learning-rate decay
 * (run_nerf.py:1289-1293) and the rectification term (radam.py:63-78) exactly as the host computes them. */
int pn_radam_step_dyn(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float beta1,
                      float beta2, float eps, const float *dyn, int mode, pn_stream_t stream);
/* dst[0..n) (device) = values[0..n) (HOST, n <= 16), stream-ordered: the values travel as kernel arguments, so the host
 * array may be reused as soon as the call returns. */
int pn_store_floats(float *dst, const float *values, int n, pn_stream_t stream);

/* ---- data formats either side of the path -------------------------------------- */

/* A training batch straight from (image, pixel) ids: replaces the precomputed rays_rgb tensor of the
 * use_batching branch (run_nerf.py:896-920 builds [N*H*W, ro+rd+rgb, 3] fp32 = 36 B/ray; :962-966 slices it)
 * and the full-image get_rays + gather of the no_batching branch (run_nerf.py:976-1004).
 *   ids          device int64 [n_rays]; ids[b] = (slot*H + j)*W + i, slot = position of the image in i_train
 *                (the row order of rays_rgb before shuffling), (j, i) = pixel row / column.
 *   K            host double[9], row-major intrinsics (fx = K[0], cx = K[2], fy = K[4], cy = K[5]).
 *   poses        device fp32, camera-to-world [3,4] row-major at poses + image*pose_stride (pose_stride = 12 for
 *                [N,3,4], 16 for [N,4,4]).
 *   image_index  device int32 [n_slots] mapping slot -> image row, or NULL for identity.
 *   images       device [n_images, H, W, 3], fp32 (image_dtype 0) or uint8 (image_dtype 1, converted as
 *                load_blender.py:62 does: float32(u8 / 255.0) with the division in double); NULL with target NULL.
 *   f64_dirs     1: ray directions in the arithmetic of get_rays_np as train() calls it (run_nerf_helpers.py:323-330
 *                with a float64 K: float64 ops, rounded to fp32 by run_nerf.py:905); 0: the fp32 arithmetic of
 *                get_rays (run_nerf_helpers.py:311-320).  Both bit-exact.
 *   batch_rays   device fp32 [2, n_rays, 3] (origins, then directions); target device fp32 [n_rays, 3]. */
int pn_ray_bank_batch(const int64_t *ids, int64_t n_rays, int height, int width, const double *K,
                      const float *poses, int64_t pose_stride, const int32_t *image_index, const void *images,
                      int image_dtype, int f64_dirs, float *batch_rays, float *target, pn_stream_t stream);

/* *sum += sum_k (a[k]-b[k])^2 (differences and squares rounded in fp32 as np.square(rgb - gt) does, accumulated in
 * fp64; sum is a caller-zeroed DEVICE double) — the numerator of render_path's PSNR (run_nerf.py:186) and
 * ComprehensiveEvaluator.compute_metrics (evaluation_utils.py:24-25) without the image leaving the GPU. */
int pn_image_sqerr(const float *a, const float *b, int64_t n, double *sum, pn_stream_t stream);

/* *sum += sum over interior pixels and channels of the SSIM map of two [H,W,C] fp32 images, as
 * skimage.metrics.structural_similarity(channel_axis=2, data_range=...) computes it with its defaults
 * (evaluation_utils.py:33; scikit-image 0.25.2: 7x7 uniform window, sample covariance, K1=0.01, K2=0.03,
 * mean over the map cropped by 3 pixels per side).  mean SSIM = *sum / ((H-6)*(W-6)*C). */
int pn_image_ssim(const float *a, const float *b, int height, int width, int channels, double data_range,
                  double *sum, pn_stream_t stream);

/* out[k] = uint8(255*clip(x[k],0,1))  — to8b (run_nerf_helpers.py:13), truncating like numpy's astype. */
int pn_to8b(const float *x, int64_t n, uint8_t *out, pn_stream_t stream);

/* Integer codes of LearnedBitwidthQuantizer's eval form (quantization.py:157-187), bit-packed at the learned
 * width: code = clamp(round(x/(scale+1e-8) + zp), qmin, qmax) - qmin, `bits` bits per value, 32 values per
 * `bits` little-endian 32-bit words.  qrow: device [PN_QROW] (the eval-form row).  n % 32 == 0, 1 <= bits <= 24.
 * pn_quant_unpack writes (code + qmin - zp) * scale, bit-identical to the quantiser's eval output on x. */
int pn_quant_pack(const float *x, int64_t n, const float *qrow, int bits, uint32_t *words, pn_stream_t stream);
int pn_quant_unpack(const uint32_t *words, int64_t n, const float *qrow, int bits, float *x, pn_stream_t stream);

/* Hash tables resident as integer codes (inference with a trained A-CAQ model).  Level l holds 2^log2_hashmap_size
 * entries; an entry is the pair of feature codes as two u8 (entry_bytes 2: learned width <= 8 bits), two u16
 * (entry_bytes 4: <= 16 bits), little endian, feature 0 first — or two fp32 values already dequantised (entry_bytes 8).
 * A code c stands for (c + qmin - zero_point) * scale, the value LearnedBitwidthQuantizer's eval forward gives the
 * fp32 entry (quantization.py:183-186), so gathering codes reproduces hash_encoding.py:97-101 in eval mode bit for
 * bit while moving 2-4 bytes per corner instead of 8. */
typedef struct {
  const void *codes[PN_MAX_LEVELS];     /* device */
  int32_t entry_bytes[PN_MAX_LEVELS];   /* 2, 4 or 8 */
  float scale[PN_MAX_LEVELS];
  float zero_point[PN_MAX_LEVELS];
  float qmin[PN_MAX_LEVELS];
} pn_packed_tables;

/* pn_hash_encode_fwd on code tables (no gradient: inference). */
int pn_hash_encode_fwd_packed(const pn_hash_grid *grid, const pn_packed_tables *packed, const float *x,
                              int64_t n_points, float *feat, uint8_t *keep, pn_stream_t stream);
/* pn_field_fwd_bf16 on code tables (no feature tiles are saved: inference). */
int pn_field_fwd_bf16_packed(const pn_hash_grid *grid, const pn_packed_tables *packed, const pn_mlp_weights *w,
                             const float *pts, const float *dirs, int samples_per_ray, const float *act_q,
                             int64_t n_points, float *out, uint8_t *keep, pn_stream_t stream);

/* Container codes for pn_packed_tables: out[k] = code of x[k] as u8 (code_bytes 1) or u16 (code_bytes 2), from the
 * fp32 values (pn_quant_codes, qrow = the quantiser's eval-form row) or from a pn_quant_pack bit stream
 * (pn_quant_unpack_codes; bits <= 8*code_bytes, n % 32 == 0). */
int pn_quant_codes(const float *x, int64_t n, const float *qrow, int code_bytes, void *out, pn_stream_t stream);
int pn_quant_unpack_codes(const uint32_t *words, int64_t n, int bits, int code_bytes, void *out, pn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POCKETNERF_H */
