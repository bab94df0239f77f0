// K2 / K2b (fp32 mode) — fused SH + NeRFSmall forward and backward (run_nerf_helpers.py:265-306 with the
// create_nerf shapes 32->64->16 | [SH16,geo15]->64->64->3, optional normal head 15->32->3).
//
// One persistent CTA per SM works on tiles of 128 points.  All weights (40 KB) and every activation of
// the tile live in shared memory; nothing but the tile's inputs and outputs touches HBM.  The forward
// and the dgrad steps run "thread = (point, half of the outputs)" with the weight row broadcast from
// shared memory; the wgrad steps are small register-tiled GEMMs over the tile whose accumulators stay
// in registers across all tiles of the CTA and are flushed once with atomics.  This is the fp32
// (1e-5) mode; the bf16 tensor-core mode lives in mlp_tc.cu.
#include "hash_core.cuh"

namespace pn {

constexpr int kTile = 128;
constexpr int kMlpThreads = 256;
constexpr int P64 = 68, P32 = 36, P16 = 20, P8 = 8;   // row pitches (floats): = 4 mod 32 -> LDS.128 conflict-free

// shared-memory layout (float offsets)
struct Smem {
  // weights, nn.Linear layout [out][in], in padded to a multiple of 4
  static constexpr int S0 = 0;                   // [64][32]
  static constexpr int S1 = S0 + 64 * 32;        // [16][64]
  static constexpr int C0 = S1 + 16 * 64;        // [64][32]  (col 31 = 0)
  static constexpr int C1 = C0 + 64 * 32;        // [64][64]
  static constexpr int C2 = C1 + 64 * 64;        // [4][64]   (row 3 = 0)
  static constexpr int N0 = C2 + 4 * 64;         // [32][16]  (col 15 = 0)
  static constexpr int N0B = N0 + 32 * 16;       // [32]
  static constexpr int N2 = N0B + 32;            // [4][32]   (row 3 = 0)
  static constexpr int N2B = N2 + 4 * 32;        // [4]
  static constexpr int WEND = N2B + 4;
  // activations of the tile
  static constexpr int X = WEND;                 // [128][P32]  hash features, later dX
  static constexpr int H1 = X + kTile * P32;     // [128][P64]  relu(.) (quantised if act_q), later dH1pre
  static constexpr int CIN = H1 + kTile * P64;   // [128][P32]  [sh16, geo15, 0]
  static constexpr int A1 = CIN + kTile * P32;   // [128][P64]  colour hidden 1, later its pre-activation grad
  static constexpr int A2 = A1 + kTile * P64;    // [128][P64]  colour hidden 2, later its pre-activation grad
  static constexpr int FWD_END = A2 + kTile * P64;
  static constexpr int NH = FWD_END;             // [128][P32]  normal-head hidden, later its grad
  static constexpr int NR = NH + kTile * P32;    // [128][P8]   raw normal (0..2), d raw normal (4..6)
  static constexpr int FWDN_END = NR + kTile * P8;
  static constexpr int DH2 = FWDN_END;           // [128][P16]  grad of the sigma-net output
  static constexpr int DO = DH2 + kTile * P16;   // [128][P8]   dout tile
  static constexpr int MSK = DO + kTile * P8;    // [128][2] uint32: relu mask of H1
  static constexpr int BWD_END = MSK + kTile * 2;
};

struct MlpArgs {
  pn_mlp_weights w;
  pn_mlp_input in;
  bool normals;
  int C;   // 4 or 7
};

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void sts4(float *p, float a, float b, float c, float d) {
  *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}

__device__ void load_weights(float *sm, const pn_mlp_weights &w, bool normals) {
  const int t = threadIdx.x;
  for (int i = t; i < 64 * 32; i += kMlpThreads) sm[Smem::S0 + i] = __ldg(w.s0 + i);
  for (int i = t; i < 16 * 64; i += kMlpThreads) sm[Smem::S1 + i] = __ldg(w.s1 + i);
  for (int i = t; i < 64 * 32; i += kMlpThreads) {
    const int j = i >> 5, k = i & 31;
    sm[Smem::C0 + i] = (k < 31) ? __ldg(w.c0 + j * 31 + k) : 0.f;
  }
  for (int i = t; i < 64 * 64; i += kMlpThreads) sm[Smem::C1 + i] = __ldg(w.c1 + i);
  for (int i = t; i < 4 * 64; i += kMlpThreads) sm[Smem::C2 + i] = (i < 3 * 64) ? __ldg(w.c2 + i) : 0.f;
  for (int i = t; i < 32 * 16 + 32 + 4 * 32 + 4; i += kMlpThreads) sm[Smem::N0 + i] = 0.f;
  if (normals) {
    __syncthreads();
    for (int i = t; i < 32 * 16; i += kMlpThreads) {
      const int j = i >> 4, k = i & 15;
      if (k < 15) sm[Smem::N0 + i] = __ldg(w.n0w + j * 15 + k);
    }
    for (int i = t; i < 32; i += kMlpThreads) sm[Smem::N0B + i] = __ldg(w.n0b + i);
    for (int i = t; i < 3 * 32; i += kMlpThreads) sm[Smem::N2 + i] = __ldg(w.n2w + i);
    for (int i = t; i < 3; i += kMlpThreads) sm[Smem::N2B + i] = __ldg(w.n2b + i);
  }
  __syncthreads();
}

// out[j0 .. j0+NJ) of one row: act(sum_k x[k] * W[j][k] + b[j]).  x in registers, W broadcast from smem.
template <int K, int NJ, bool RELU>
__device__ __forceinline__ void dense_row(const float *x, const float *W, const float *bias, int j0, float *out_row) {
#pragma unroll 1
  for (int jj = 0; jj < NJ; jj += 4) {
    float acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = bias ? bias[j0 + jj + i] : 0.f;
#pragma unroll
    for (int k4 = 0; k4 < K; k4 += 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 wv = lds4(W + (j0 + jj + i) * K + k4);
        acc[i] = fmaf(x[k4 + 0], wv.x, acc[i]);
        acc[i] = fmaf(x[k4 + 1], wv.y, acc[i]);
        acc[i] = fmaf(x[k4 + 2], wv.z, acc[i]);
        acc[i] = fmaf(x[k4 + 3], wv.w, acc[i]);
      }
    }
    if (RELU) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
    sts4(out_row + j0 + jj, acc[0], acc[1], acc[2], acc[3]);
  }
}

template <int K>
__device__ __forceinline__ void load_row(const float *row, float *x) {
#pragma unroll
  for (int k4 = 0; k4 < K; k4 += 4) {
    const float4 v = lds4(row + k4);
    x[k4] = v.x; x[k4 + 1] = v.y; x[k4 + 2] = v.z; x[k4 + 3] = v.w;
  }
}

// acc[k - k0] = sum_j dy[j] * W[j][k]  for k in [k0, k0+NK);  W pitch = KW.
template <int NJ, int KW, int NK>
__device__ __forceinline__ void dgrad_row(const float *dy, const float *W, int k0, float *acc) {
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
#pragma unroll
    for (int k4 = 0; k4 < NK; k4 += 4) {
      const float4 wv = lds4(W + j * KW + k0 + k4);
      acc[k4 + 0] = fmaf(dy[j], wv.x, acc[k4 + 0]);
      acc[k4 + 1] = fmaf(dy[j], wv.y, acc[k4 + 1]);
      acc[k4 + 2] = fmaf(dy[j], wv.z, acc[k4 + 2]);
      acc[k4 + 3] = fmaf(dy[j], wv.w, acc[k4 + 3]);
    }
  }
}

// acc[a][b] += sum_p dY[p][tj*TJ + a] * X[p][tk*TK + b]   over the 128 rows of the tile.
template <int N, int K, int TJ, int TK>
__device__ __forceinline__ void wgrad_tile(const float *dY, int pitchY, const float *X, int pitchX, float *acc) {
  constexpr int KT = K / TK;
  const int tj = threadIdx.x / KT, tk = threadIdx.x - tj * KT;
  if (tj >= N / TJ) return;
  const float *py = dY + tj * TJ, *px = X + tk * TK;
#pragma unroll 4
  for (int p = 0; p < kTile; ++p) {
    float dy[TJ], xv[TK];
    if constexpr (TJ == 4) {
      const float4 v = lds4(py + p * pitchY);
      dy[0] = v.x; dy[1] = v.y; dy[2] = v.z; dy[3] = v.w;
    } else {
#pragma unroll
      for (int a = 0; a < TJ; ++a) dy[a] = py[p * pitchY + a];
    }
    if constexpr (TK == 4) {
      const float4 v = lds4(px + p * pitchX);
      xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
    } else if constexpr (TK == 2) {
      const float2 v = *reinterpret_cast<const float2 *>(px + p * pitchX);
      xv[0] = v.x; xv[1] = v.y;
    } else {
#pragma unroll
      for (int b = 0; b < TK; ++b) xv[b] = px[p * pitchX + b];
    }
#pragma unroll
    for (int a = 0; a < TJ; ++a)
#pragma unroll
      for (int b = 0; b < TK; ++b) acc[a * TK + b] = fmaf(dy[a], xv[b], acc[a * TK + b]);
  }
}

template <int N, int K, int TJ, int TK>
__device__ __forceinline__ void wgrad_flush(float *dst, int dst_pitch, int n_valid, int k_valid, const float *acc) {
  constexpr int KT = K / TK;
  const int tj = threadIdx.x / KT, tk = threadIdx.x - tj * KT;
  if (tj >= N / TJ || dst == nullptr) return;
#pragma unroll
  for (int a = 0; a < TJ; ++a)
#pragma unroll
    for (int b = 0; b < TK; ++b) {
      const int j = tj * TJ + a, k = tk * TK + b;
      if (j < n_valid && k < k_valid) atomicAdd(dst + j * dst_pitch + k, acc[a * TK + b]);
    }
}

// ---- tile input: features, SH, (dout) ------------------------------------------------------------
__device__ void load_tile(float *sm, const MlpArgs &A, int64_t base, int rows) {
  const int t = threadIdx.x;
  for (int i = t; i < kTile * 8; i += kMlpThreads) {
    const int row = i >> 3, c4 = (i & 7) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) v = __ldg(reinterpret_cast<const float4 *>(A.in.feat + (base + row) * A.in.feat_stride + c4));
    *reinterpret_cast<float4 *>(sm + Smem::X + row * P32 + c4) = v;
  }
  if (A.in.sh) {
    for (int i = t; i < kTile * 4; i += kMlpThreads) {
      const int row = i >> 2, c4 = (i & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows) v = __ldg(reinterpret_cast<const float4 *>(A.in.sh + (base + row) * A.in.sh_stride + c4));
      *reinterpret_cast<float4 *>(sm + Smem::CIN + row * P32 + c4) = v;
    }
  } else if (t < kTile) {
    float o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = 0.f;
    if (t < rows) {
      const int64_t r = (base + t) / A.in.samples_per_ray;
      sh4(__ldg(A.in.dirs + 3 * r), __ldg(A.in.dirs + 3 * r + 1), __ldg(A.in.dirs + 3 * r + 2), o);
    }
    float *dst = sm + Smem::CIN + t * P32;
#pragma unroll
    for (int q = 0; q < 4; ++q) sts4(dst + 4 * q, o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  }
}

// ---- forward phases (shared by the forward kernel and the backward's recompute) --------------------
// Returns sigma (valid in half-0 threads).
template <bool BWD>
__device__ __forceinline__ float forward_tile(float *sm, const MlpArgs &A, const float *qrow) {
  const int p = threadIdx.x & (kTile - 1), half = threadIdx.x >> 7;
  float sigma = 0.f;
  {  // F1: H1 = relu(X S0^T), optional activation fake-quant (run_nerf_helpers.py:276-284)
    float x[32];
    load_row<32>(sm + Smem::X + p * P32, x);
    float *h = sm + Smem::H1 + p * P64;
    dense_row<32, 32, true>(x, sm + Smem::S0, nullptr, half * 32, h);
    if (BWD || qrow) {
      uint32_t m = 0;
      const bool quant = qrow && qrow[5] != 0.f;
      float q0 = 0.f, q1 = 1.f, q2 = 0.f, q3 = 0.f, q4 = 0.f;
      bool qtrain = false;
      if (quant) { q0 = qrow[0]; q1 = qrow[1]; q2 = qrow[2]; q3 = qrow[3]; q4 = qrow[4]; qtrain = qrow[6] != 0.f; }
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) {
        const float4 hv = lds4(h + half * 32 + j4);
        float v[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (v[i] > 0.f) m |= (1u << (j4 + i));
          if (quant) v[i] = fake_quant(v[i], q0, q1, q2, q3, q4, qtrain);
        }
        if (quant) sts4(h + half * 32 + j4, v[0], v[1], v[2], v[3]);
      }
      if (BWD) reinterpret_cast<uint32_t *>(sm + Smem::MSK)[p * 2 + half] = m;
    }
  }
  __syncthreads();
  {  // F2: [sigma, geo] = H1 S1^T  -> CIN[16..31)
    float x[64];
    load_row<64>(sm + Smem::H1 + p * P64, x);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    // 8 outputs per half, accumulate in registers (dense_row writes to smem; here outputs are remapped)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float *wr = sm + Smem::S1 + (half * 8 + i) * 64;
      float acc = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < 64; k4 += 4) {
        const float4 wv = lds4(wr + k4);
        acc = fmaf(x[k4], wv.x, acc); acc = fmaf(x[k4 + 1], wv.y, acc);
        acc = fmaf(x[k4 + 2], wv.z, acc); acc = fmaf(x[k4 + 3], wv.w, acc);
      }
      o[i] = acc;
    }
    float *c = sm + Smem::CIN + p * P32 + 16;
    if (half == 0) {
      sigma = o[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) c[i - 1] = o[i];            // geo 0..6  -> CIN[16..22]
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) c[7 + i] = o[i];            // geo 7..14 -> CIN[23..30]
      c[15] = 0.f;                                            // CIN[31] pad
    }
  }
  __syncthreads();
  {  // F3: A1 = relu(CIN C0^T);  normal head hidden NH = relu(geo N0^T + b)
    float x[32];
    load_row<32>(sm + Smem::CIN + p * P32, x);
    dense_row<32, 32, true>(x, sm + Smem::C0, nullptr, half * 32, sm + Smem::A1 + p * P64);
    if (A.normals) dense_row<16, 16, true>(x + 16, sm + Smem::N0, sm + Smem::N0B, half * 16, sm + Smem::NH + p * P32);
  }
  __syncthreads();
  {  // F4: A2 = relu(A1 C1^T);  raw normal = NH N2^T + b
    float x[64];
    load_row<64>(sm + Smem::A1 + p * P64, x);
    dense_row<64, 32, true>(x, sm + Smem::C1, nullptr, half * 32, sm + Smem::A2 + p * P64);
    if (A.normals && half == 1) {
      float y[32];
      load_row<32>(sm + Smem::NH + p * P32, y);
      dense_row<32, 4, false>(y, sm + Smem::N2, sm + Smem::N2B, 0, sm + Smem::NR + p * P8);
    }
  }
  __syncthreads();
  return sigma;
}

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fwd_kernel(const MlpArgs A, float *__restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  load_weights(sm, A.w, A.normals);
  const int p = threadIdx.x & (kTile - 1), half = threadIdx.x >> 7;
  const int64_t n_tiles = (A.in.n_points + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kTile;
    const int rows = (int)((A.in.n_points - base) < kTile ? (A.in.n_points - base) : kTile);
    load_tile(sm, A, base, rows);
    __syncthreads();
    const float sigma = forward_tile<false>(sm, A, A.in.act_q);
    if (half == 0 && p < rows) {
      // rgb = A2 C2^T   (3 outputs)
      float x[64];
      load_row<64>(sm + Smem::A2 + p * P64, x);
      float rgb[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float *wr = sm + Smem::C2 + i * 64;
        float acc = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 64; k4 += 4) {
          const float4 wv = lds4(wr + k4);
          acc = fmaf(x[k4], wv.x, acc); acc = fmaf(x[k4 + 1], wv.y, acc);
          acc = fmaf(x[k4 + 2], wv.z, acc); acc = fmaf(x[k4 + 3], wv.w, acc);
        }
        rgb[i] = acc;
      }
      const bool kept = A.in.keep ? (A.in.keep[base + p] != 0) : true;
      float *o = out + (base + p) * A.C;
      if (A.C == 4) {
        *reinterpret_cast<float4 *>(o) = make_float4(rgb[0], rgb[1], rgb[2], kept ? sigma : 0.f);   // run_nerf.py:66
      } else {
        const float *nr = sm + Smem::NR + p * P8;
        const float nn = fmaxf(sqrtf(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]), 1e-12f);
        o[0] = rgb[0]; o[1] = rgb[1]; o[2] = rgb[2]; o[3] = sigma;
        o[4] = nr[0] / nn; o[5] = nr[1] / nn; o[6] = kept ? nr[2] / nn : 0.f;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_bwd_kernel(const MlpArgs A, const float *__restrict__ dout, float *__restrict__ dfeat, int64_t dfeat_stride,
               float *__restrict__ dsh, int64_t dsh_stride, const pn_mlp_grads G) {
  extern __shared__ __align__(16) float sm[];
  load_weights(sm, A.w, A.normals);
  const int t = threadIdx.x, p = t & (kTile - 1), half = t >> 7;
  // weight-gradient accumulators, live across all tiles of this CTA
  float g_c1[16], g_s0[8], g_c0[8], g_s1[4], g_c2[1], g_n2[1], g_n0[2], g_nb = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) g_c1[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { g_s0[i] = 0.f; g_c0[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < 4; ++i) g_s1[i] = 0.f;
  g_c2[0] = g_n2[0] = g_n0[0] = g_n0[1] = 0.f;

  const int64_t n_tiles = (A.in.n_points + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kTile;
    const int rows = (int)((A.in.n_points - base) < kTile ? (A.in.n_points - base) : kTile);
    load_tile(sm, A, base, rows);
    // dout tile -> DO[p][0..C), keep mask applied to the last channel
    for (int i = t; i < kTile * P8; i += kMlpThreads) {
      const int row = i >> 3, c = i & 7;
      float v = 0.f;
      if (row < rows && c < A.C) {
        v = __ldg(dout + (base + row) * A.C + c);
        if (c == A.C - 1 && A.in.keep && A.in.keep[base + row] == 0) v = 0.f;
      }
      sm[Smem::DO + row * P8 + c] = v;
    }
    __syncthreads();
    forward_tile<true>(sm, A, A.in.act_q);

    // B0: normal head: gradient of normalize -> d raw normal in NR[4..6]
    if (A.normals && half == 0) {
      float *nr = sm + Smem::NR + p * P8;
      const float *d = sm + Smem::DO + p * P8;
      const float nn = sqrtf(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
      float r0, r1, r2;
      if (nn > 1e-12f) {
        const float m0 = nr[0] / nn, m1 = nr[1] / nn, m2 = nr[2] / nn;
        const float dot = m0 * d[4] + m1 * d[5] + m2 * d[6];
        r0 = (d[4] - m0 * dot) / nn; r1 = (d[5] - m1 * dot) / nn; r2 = (d[6] - m2 * dot) / nn;
      } else {
        r0 = d[4] / 1e-12f; r1 = d[5] / 1e-12f; r2 = d[6] / 1e-12f;
      }
      nr[4] = r0; nr[5] = r1; nr[6] = r2; nr[7] = 0.f;
    }
    __syncthreads();
    // B1: dC2w += drgb^T A2 ;  dN2w += dnraw^T NH ; dN2b
    wgrad_tile<4, 64, 1, 1>(sm + Smem::DO, P8, sm + Smem::A2, P64, g_c2);
    if (A.normals) {
      wgrad_tile<4, 32, 1, 1>(sm + Smem::NR + 4, P8, sm + Smem::NH, P32, g_n2);
      if (t >= 128 && t < 131) {
        float s = 0.f;
        for (int q = 0; q < kTile; ++q) s += sm[Smem::NR + q * P8 + 4 + (t - 128)];
        g_nb += s;
      }
    }
    __syncthreads();
    // B2: dA2pre = (drgb C2w) * [A2 > 0] in place;  dNHpre = (dnraw N2w) * [NH > 0] in place
    {
      float dy[4];
      const float4 v = lds4(sm + Smem::DO + p * P8);
      dy[0] = v.x; dy[1] = v.y; dy[2] = v.z; dy[3] = 0.f;
      float acc[32];
      dgrad_row<3, 64, 32>(dy, sm + Smem::C2, half * 32, acc);
      float *a2 = sm + Smem::A2 + p * P64 + half * 32;
#pragma unroll
      for (int k4 = 0; k4 < 32; k4 += 4) {
        const float4 a = lds4(a2 + k4);
        sts4(a2 + k4, a.x > 0.f ? acc[k4] : 0.f, a.y > 0.f ? acc[k4 + 1] : 0.f, a.z > 0.f ? acc[k4 + 2] : 0.f,
             a.w > 0.f ? acc[k4 + 3] : 0.f);
      }
      if (A.normals) {
        const float4 u = lds4(sm + Smem::NR + p * P8 + 4);
        float dn[3] = {u.x, u.y, u.z};
        float an[16];
        dgrad_row<3, 32, 16>(dn, sm + Smem::N2, half * 16, an);
        float *nh = sm + Smem::NH + p * P32 + half * 16;
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
          const float4 a = lds4(nh + k4);
          sts4(nh + k4, a.x > 0.f ? an[k4] : 0.f, a.y > 0.f ? an[k4 + 1] : 0.f, a.z > 0.f ? an[k4 + 2] : 0.f,
               a.w > 0.f ? an[k4 + 3] : 0.f);
        }
      }
    }
    __syncthreads();
    // B3: dC1w += dA2pre^T A1 ;  dN0w += dNHpre^T geo ; dN0b
    wgrad_tile<64, 64, 4, 4>(sm + Smem::A2, P64, sm + Smem::A1, P64, g_c1);
    if (A.normals) {
      wgrad_tile<32, 16, 1, 2>(sm + Smem::NH, P32, sm + Smem::CIN + 16, P32, g_n0);
      if (t < 32) {
        float s = 0.f;
        for (int q = 0; q < kTile; ++q) s += sm[Smem::NH + q * P32 + t];
        g_nb += s;
      }
    }
    __syncthreads();
    // B4: dA1pre = (dA2pre C1w) * [A1 > 0] in place
    {
      float dy[64];
      load_row<64>(sm + Smem::A2 + p * P64, dy);
      float acc[32];
      dgrad_row<64, 64, 32>(dy, sm + Smem::C1, half * 32, acc);
      float *a1 = sm + Smem::A1 + p * P64 + half * 32;
#pragma unroll
      for (int k4 = 0; k4 < 32; k4 += 4) {
        const float4 a = lds4(a1 + k4);
        sts4(a1 + k4, a.x > 0.f ? acc[k4] : 0.f, a.y > 0.f ? acc[k4 + 1] : 0.f, a.z > 0.f ? acc[k4 + 2] : 0.f,
             a.w > 0.f ? acc[k4 + 3] : 0.f);
      }
    }
    __syncthreads();
    // B5: dC0w += dA1pre^T CIN
    wgrad_tile<64, 32, 4, 2>(sm + Smem::A1, P64, sm + Smem::CIN, P32, g_c0);
    __syncthreads();
    // B6: dCIN = dA1pre C0w (+ normal head into geo) -> DH2 = [dsigma, dgeo];  dsh out
    {
      float dy[64];
      load_row<64>(sm + Smem::A1 + p * P64, dy);
      float acc[16];
      dgrad_row<64, 32, 16>(dy, sm + Smem::C0, half * 16, acc);
      float *dh = sm + Smem::DH2 + p * P16;
      if (half == 0) {
        if (dsh && p < rows) {
          float *o = dsh + (base + p) * dsh_stride;
#pragma unroll
          for (int k4 = 0; k4 < 16; k4 += 4)
            *reinterpret_cast<float4 *>(o + k4) = make_float4(acc[k4], acc[k4 + 1], acc[k4 + 2], acc[k4 + 3]);
        }
        dh[0] = sm[Smem::DO + p * P8 + 3];                      // dsigma (keep mask already applied when C == 4)
      } else {
        if (A.normals) {
          float dn[32];
          load_row<32>(sm + Smem::NH + p * P32, dn);
          float an[16];
          dgrad_row<32, 16, 16>(dn, sm + Smem::N0, 0, an);
#pragma unroll
          for (int k = 0; k < 15; ++k) acc[k] += an[k];
        }
#pragma unroll
        for (int k = 0; k < 15; ++k) dh[1 + k] = acc[k];          // dgeo = dCIN[16..31)
      }
    }
    __syncthreads();
    // B7: dS1w += DH2^T H1
    wgrad_tile<16, 64, 1, 4>(sm + Smem::DH2, P16, sm + Smem::H1, P64, g_s1);
    __syncthreads();
    // B8: dH1pre = (DH2 S1w) * relu-mask in place (fake-quant is a straight-through estimator)
    {
      float dy[16];
      load_row<16>(sm + Smem::DH2 + p * P16, dy);
      float acc[32];
      dgrad_row<16, 64, 32>(dy, sm + Smem::S1, half * 32, acc);
      const uint32_t m = reinterpret_cast<const uint32_t *>(sm + Smem::MSK)[p * 2 + half];
      float *h1 = sm + Smem::H1 + p * P64 + half * 32;
#pragma unroll
      for (int k4 = 0; k4 < 32; k4 += 4)
        sts4(h1 + k4, (m >> k4) & 1u ? acc[k4] : 0.f, (m >> (k4 + 1)) & 1u ? acc[k4 + 1] : 0.f,
             (m >> (k4 + 2)) & 1u ? acc[k4 + 2] : 0.f, (m >> (k4 + 3)) & 1u ? acc[k4 + 3] : 0.f);
    }
    __syncthreads();
    // B9: dS0w += dH1pre^T X
    wgrad_tile<64, 32, 4, 2>(sm + Smem::H1, P64, sm + Smem::X, P32, g_s0);
    __syncthreads();
    // B10: dX = dH1pre S0w -> X in place -> global
    {
      float dy[64];
      load_row<64>(sm + Smem::H1 + p * P64, dy);
      float acc[16];
      dgrad_row<64, 32, 16>(dy, sm + Smem::S0, half * 16, acc);
      float *xr = sm + Smem::X + p * P32 + half * 16;
#pragma unroll
      for (int k4 = 0; k4 < 16; k4 += 4) sts4(xr + k4, acc[k4], acc[k4 + 1], acc[k4 + 2], acc[k4 + 3]);
    }
    __syncthreads();
    for (int i = t; i < kTile * 8; i += kMlpThreads) {
      const int row = i >> 3, c4 = (i & 7) * 4;
      if (row < rows)
        *reinterpret_cast<float4 *>(dfeat + (base + row) * dfeat_stride + c4) = lds4(sm + Smem::X + row * P32 + c4);
    }
    __syncthreads();
  }
  // flush the weight gradients
  wgrad_flush<64, 64, 4, 4>(G.c1, 64, 64, 64, g_c1);
  wgrad_flush<64, 32, 4, 2>(G.s0, 32, 64, 32, g_s0);
  wgrad_flush<64, 32, 4, 2>(G.c0, 31, 64, 31, g_c0);
  wgrad_flush<16, 64, 1, 4>(G.s1, 64, 16, 64, g_s1);
  wgrad_flush<4, 64, 1, 1>(G.c2, 64, 3, 64, g_c2);
  if (A.normals) {
    wgrad_flush<4, 32, 1, 1>(G.n2w, 32, 3, 32, g_n2);
    wgrad_flush<32, 16, 1, 2>(G.n0w, 15, 32, 15, g_n0);
    if (t < 32 && G.n0b) atomicAdd(G.n0b + t, g_nb);
    if (t >= 128 && t < 131 && G.n2b) atomicAdd(G.n2b + (t - 128), g_nb);
  }
}

int check_mlp_args(const pn_mlp_weights *w, const pn_mlp_input *in, bool *normals) {
  PN_REQUIRE(w && in, PN_EINVAL, "NULL argument");
  PN_REQUIRE(w->s0 && w->s1 && w->c0 && w->c1 && w->c2, PN_EINVAL, "NULL weight pointer");
  const int n_normal = (w->n0w != nullptr) + (w->n0b != nullptr) + (w->n2w != nullptr) + (w->n2b != nullptr);
  PN_REQUIRE(n_normal == 0 || n_normal == 4, PN_EINVAL, "normal head needs all of n0w,n0b,n2w,n2b");
  *normals = n_normal == 4;
  PN_REQUIRE(in->feat != nullptr, PN_EINVAL, "feat is NULL");
  PN_REQUIRE((in->sh != nullptr) != (in->dirs != nullptr), PN_EINVAL, "exactly one of sh / dirs must be given");
  PN_REQUIRE(in->feat_stride >= 32 && in->feat_stride % 4 == 0, PN_EINVAL, "feat_stride %lld (>=32, multiple of 4)",
             (long long)in->feat_stride);
  PN_REQUIRE(((uintptr_t)in->feat & 15) == 0, PN_EINVAL, "feat must be 16-byte aligned");
  if (in->sh) {
    PN_REQUIRE(in->sh_stride >= 16 && in->sh_stride % 4 == 0 && ((uintptr_t)in->sh & 15) == 0, PN_EINVAL,
               "sh stride/alignment");
  } else {
    PN_REQUIRE(in->samples_per_ray >= 1, PN_EINVAL, "samples_per_ray %d", in->samples_per_ray);
  }
  PN_REQUIRE(in->n_points >= 0, PN_EINVAL, "n_points < 0");
  return 0;
}

}  // namespace pn

using namespace pn;

extern "C" int pn_mlp_fwd(const pn_mlp_weights *w, const pn_mlp_input *in, float *out, pn_stream_t stream) {
  bool normals = false;
  if (int e = check_mlp_args(w, in, &normals)) return e;
  PN_REQUIRE(out != nullptr, PN_EINVAL, "out is NULL");
  if (in->n_points == 0) return 0;
  MlpArgs A;
  A.w = *w; A.in = *in; A.normals = normals; A.C = normals ? 7 : 4;
  PN_REQUIRE(A.C == 7 || ((uintptr_t)out & 15) == 0, PN_EINVAL, "out must be 16-byte aligned");
  const size_t smem = (size_t)(normals ? Smem::FWDN_END : Smem::FWD_END) * sizeof(float);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(Smem::FWDN_END * sizeof(float)));
    PN_REQUIRE(e == cudaSuccess, PN_ECUDA, "cudaFuncSetAttribute(mlp_fwd): %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  const int64_t tiles = ceil_div(in->n_points, kTile);
  const int blocks = (int)(tiles < sm_count() ? tiles : sm_count());
  mlp_fwd_kernel<<<blocks, kMlpThreads, smem, as_stream(stream)>>>(A, out);
  count_launch();
  return check_launch("mlp_fwd_kernel");
}

extern "C" int pn_mlp_bwd(const pn_mlp_weights *w, const pn_mlp_input *in, const float *dout, float *dfeat,
                          int64_t dfeat_stride, float *dsh, int64_t dsh_stride, const pn_mlp_grads *dw,
                          pn_stream_t stream) {
  bool normals = false;
  if (int e = check_mlp_args(w, in, &normals)) return e;
  PN_REQUIRE(dout && dfeat && dw, PN_EINVAL, "NULL pointer argument");
  PN_REQUIRE(dfeat_stride >= 32 && dfeat_stride % 4 == 0 && ((uintptr_t)dfeat & 15) == 0, PN_EINVAL,
             "dfeat stride/alignment");
  PN_REQUIRE(dsh == nullptr || (in->sh != nullptr && dsh_stride >= 16 && dsh_stride % 4 == 0 &&
                                ((uintptr_t)dsh & 15) == 0),
             PN_EINVAL, "dsh needs in->sh, stride >= 16 (multiple of 4) and 16-byte alignment");
  if (in->n_points == 0) return 0;
  MlpArgs A;
  A.w = *w; A.in = *in; A.normals = normals; A.C = normals ? 7 : 4;
  const size_t smem = (size_t)Smem::BWD_END * sizeof(float);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    PN_REQUIRE(e == cudaSuccess, PN_ECUDA, "cudaFuncSetAttribute(mlp_bwd): %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  const int64_t tiles = ceil_div(in->n_points, kTile);
  const int blocks = (int)(tiles < sm_count() ? tiles : sm_count());
  mlp_bwd_kernel<<<blocks, kMlpThreads, smem, as_stream(stream)>>>(A, dout, dfeat, dfeat_stride, dsh, dsh_stride,
                                                                    *dw);
  count_launch();
  return check_launch("mlp_bwd_kernel");
}
