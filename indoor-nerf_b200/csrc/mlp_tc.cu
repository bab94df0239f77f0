// K2 / K2b (bf16 tensor-core mode) — NeRFSmall on tcgen05 with TMEM accumulators.
//   thread = point = TMEM lane; a CTA owns a tile of 128 points; every layer is a 128 x N x K UMMA whose
//   A operand (the previous layer's activations, bf16) and B operand (the weights, bf16) sit in shared
//   memory in the tile layout of tc_common.cuh; the epilogue of each layer (tcgen05.ld -> ReLU -> bf16 ->
//   st.shared) produces the next layer's A operand.  Nothing but the tile's inputs and outputs touches HBM.
#include <stdlib.h>

#include "hash_core.cuh"
#include "hash_scatter.cuh"
#include "io_core.cuh"
#include "tc_common.cuh"

namespace pn {
using namespace tc;

// ------------------------------------------------------------------------------------------------------
// Descriptor self-test: D = A(.)B with host-chosen descriptor fields, so that the K-major / MN-major
// readings of the tile layout and the M=64 / M=128 accumulator layouts are proven on hardware, not assumed.
// ------------------------------------------------------------------------------------------------------
struct SelfTestArgs {
  int ra, ca, rb, cb;              // tile shapes (rows, cols) of the two operands as stored
  int M, N, ksteps;                // instruction shape and number of K=16 steps
  int a_mn, b_mn;                  // 1 = read the operand MN-major
  uint32_t lbo_a, sbo_a, adv_a;    // descriptor fields (bytes) and per-k-step start-address advance
  uint32_t lbo_b, sbo_b, adv_b;
};

__global__ void __launch_bounds__(128) tc_selftest_kernel(SelfTestArgs a, const float *__restrict__ A,
                                                           const float *__restrict__ B, float *__restrict__ D) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  uint8_t *ta = smem;
  uint8_t *tb = smem + ((a.ra * a.ca * 2 + 127) / 128) * 128;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < a.ra * (a.ca / 8); i += 128) {
    const int r = i / (a.ca / 8), c8 = i % (a.ca / 8);
    float v[8];
    for (int j = 0; j < 8; ++j) v[j] = A[r * a.ca + c8 * 8 + j];
    st_chunk(ta, chunk_off(r, c8, a.ca / 8), v);
  }
  for (int i = t; i < a.rb * (a.cb / 8); i += 128) {
    const int r = i / (a.cb / 8), c8 = i % (a.cb / 8);
    float v[8];
    for (int j = 0; j < 8; ++j) v[j] = B[r * a.cb + c8 * 8 + j];
    st_chunk(tb, chunk_off(r, c8, a.cb / 8), v);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    const uint32_t idesc = instr_desc(a.M, a.N, a.a_mn, a.b_mn);
    for (int k = 0; k < a.ksteps; ++k) {
      const uint64_t da = smem_desc(smem_u32(ta) + k * a.adv_a, a.lbo_a, a.sbo_a);
      const uint64_t db = smem_desc(smem_u32(tb) + k * a.adv_b, a.lbo_b, a.sbo_b);
      mma_f16(tmem, da, db, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c = 0; c < a.N; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[t * a.N + c + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}


// ------------------------------------------------------------------------------------------------------
// NeRFSmall on the tensor cores
// ------------------------------------------------------------------------------------------------------
constexpr int kTcTile = 128;     // points per CTA tile = TMEM lanes = threads

// shared-memory map (bytes).  All tiles in the tile layout of tc_common.cuh.
struct TS {
  static constexpr int W_S0 = 0;                  // [64 x 32]
  static constexpr int W_S1 = W_S0 + 64 * 32 * 2; // [16 x 64]
  static constexpr int W_C0 = W_S1 + 16 * 64 * 2; // [64 x 32] (col 31 = 0)
  static constexpr int W_C1 = W_C0 + 64 * 32 * 2; // [64 x 64]
  static constexpr int W_C2 = W_C1 + 64 * 64 * 2; // [16 x 64] (rows 3.. = 0)
  static constexpr int W_N0 = W_C2 + 16 * 64 * 2; // [32 x 16] (col 15 = 0)
  static constexpr int W_N2 = W_N0 + 32 * 16 * 2; // [16 x 32] (rows 3.. = 0)
  static constexpr int BIAS = W_N2 + 16 * 32 * 2; // fp32: n0b[32], n2b[4]
  static constexpr int A0 = BIAS + 256;           // [128 x 32] hash features (bias block padded to 256 B)
  static constexpr int A1 = A0 + 128 * 32 * 2;    // [128 x 64] hidden (fwd: reused by every 64-wide layer; bwd: H1)
  static constexpr int CIN = A1 + 128 * 64 * 2;   // [128 x 32] [sh16, geo15, 0]
  static constexpr int NH = CIN + 128 * 32 * 2;   // [128 x 32] normal-head hidden
  static constexpr int FWD_END = NH + 128 * 32 * 2;
  // backward only
  static constexpr int A1C = FWD_END;             // [128 x 64] colour hidden 1 -> its pre-activation grad
  static constexpr int A2C = A1C + 128 * 64 * 2;  // [128 x 64] colour hidden 2 -> its pre-activation grad
  static constexpr int DOUT = A2C + 128 * 64 * 2; // [128 x 16] (drgb, 0...)
  static constexpr int DH2 = DOUT + 128 * 16 * 2; // [128 x 16] (dsigma, dgeo)
  static constexpr int DNR = DH2 + 128 * 16 * 2;  // [128 x 16] (d raw normal, 0...)
  static constexpr int BWD_END = DNR + 128 * 16 * 2;
};
static_assert(TS::A0 % 128 == 0 && TS::A1 % 128 == 0 && TS::CIN % 128 == 0 && TS::A1C % 128 == 0, "tile alignment");

// TMEM columns
constexpr uint32_t TM_D1 = 0;      // 64 columns: 64-wide layer outputs / input gradients
constexpr uint32_t TM_D2 = 64;     // 32 columns: 16-wide outputs (sigma+geo, rgb, raw normal, dgeo from the normal head)
constexpr uint32_t TM_DN = 64;     // normal-head hidden / its gradient: aliases D2 (never live in the same round)
constexpr uint32_t TM_FWD_COLS = 128;
constexpr uint32_t TM_GC1 = 96;    // wgrad accumulators (M = 64 layout: row m -> lane m%16 + 32*(m/16)), live across tiles
constexpr uint32_t TM_GS0 = 160;
constexpr uint32_t TM_GC0 = 192;
constexpr uint32_t TM_GS1 = 224;   // [64 x 16] = dS1^T
constexpr uint32_t TM_GC2 = 240;   // [64 x 16] = dC2^T
constexpr uint32_t TM_BWD_COLS = 256;

struct TcArgs {
  pn_mlp_weights w;
  pn_mlp_input in;
  int normals;
  int C;
};

// Extra arguments of the fused field kernels (hash grid + SH + NeRFSmall in one launch).
struct FieldArgs {
  HashGridDev G;
  TablePtrs T;          // forward: tables
  GradPtrs D;           // backward: table gradients
  const float *pts;     // [P,3]
  const float *qparams; // per-level fake-quant rows or NULL
  uint8_t *keep_out;    // forward: [P]
  uint4 *featb;         // bf16 feature tiles, 8 KB per 128-point tile, tile layout (forward writes, backward reads)
  int scatter_split;    // backward: half 0 scatters levels [0, split), half 1 the rest
  int debug;            // PN_DEBUG_FLAGS (measurement only): 1 = skip the scatter work, 2 = skip the gather work
  int sc_direct, sc_maxlen;   // scatter_level tuning (PN_SCATTER_DIRECT, PN_SCATTER_MAXLEN)
  long long *tlog;      // pn_debug_timeline buffer: [2 threads][tlog_cap] (+1 spare row) clock64 marks, or NULL
  int tlog_cap;
  PackedDev PK;         // SRC_PACKED: tables as integer codes (inference)
};

// Diagnostic timeline (pn_debug_timeline): when a log buffer is installed, two threads of CTA 0 of the single-role
// backward — thread 0 (the MMA issuer) and thread 160 (an ordinary epilogue thread) — write clock64() at the marked
// points of every round of their first tiles.  One predictable branch per mark otherwise; never set in production runs.
struct TLog {
  long long *p;
  int i, cap;
  __device__ __forceinline__ void mark() {
    if (p && i < cap) p[i++] = clock64();
  }
};
static long long *g_tlog_buf = nullptr;
static int g_tlog_cap = 0;

// input source of a tile's hash features
enum { SRC_F32 = 0, SRC_HASH = 1, SRC_TILE = 2, SRC_PACKED = 3 };

// fp32 [N x K] row-major global weight (ld = Kvalid) -> bf16 tile [NT x KT], zero padded
__device__ void load_w_tile(uint8_t *tile, const float *__restrict__ W, int NT, int KT, int Nvalid, int Kvalid) {
  const int C8 = KT / 8;
  for (int i = threadIdx.x; i < NT * C8; i += blockDim.x) {
    const int r = i / C8, c8 = i - r * C8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      v[j] = (r < Nvalid && c < Kvalid) ? __ldg(W + r * Kvalid + c) : 0.f;
    }
    st_chunk(tile, chunk_off(r, c8, C8), v);
  }
}

__device__ void load_all_weights(uint8_t *sm, const TcArgs &A) {
  load_w_tile(sm + TS::W_S0, A.w.s0, 64, 32, 64, 32);
  load_w_tile(sm + TS::W_S1, A.w.s1, 16, 64, 16, 64);
  load_w_tile(sm + TS::W_C0, A.w.c0, 64, 32, 64, 31);
  load_w_tile(sm + TS::W_C1, A.w.c1, 64, 64, 64, 64);
  load_w_tile(sm + TS::W_C2, A.w.c2, 16, 64, 3, 64);
  float *bias = reinterpret_cast<float *>(sm + TS::BIAS);
  if (A.normals) {
    load_w_tile(sm + TS::W_N0, A.w.n0w, 32, 16, 32, 15);
    load_w_tile(sm + TS::W_N2, A.w.n2w, 16, 32, 3, 32);
    for (int i = threadIdx.x; i < 36; i += blockDim.x)
      bias[i] = (i < 32) ? __ldg(A.w.n0b + i) : (i < 35 ? __ldg(A.w.n2b + (i - 32)) : 0.f);
  }
}

// operand descriptors of a tile: K-major (c = K) or MN-major (r = K)
struct Opnd {
  uint32_t addr, lbo, sbo, adv;
};
__device__ __forceinline__ Opnd k_major(const uint8_t *tile, int cols, int col0 = 0) {
  return Opnd{smem_u32(tile) + (uint32_t)(col0 / 8) * 128u, 128u, (uint32_t)(cols / 8) * 128u, 256u};
}
__device__ __forceinline__ Opnd mn_major(const uint8_t *tile, int cols) {
  const uint32_t rg = (uint32_t)(cols / 8) * 128u;
  return Opnd{smem_u32(tile), rg, 128u, 2u * rg};
}
__device__ __forceinline__ void issue(uint32_t d, const Opnd &a, const Opnd &b, uint32_t idesc, int ksteps, bool accumulate) {
  for (int k = 0; k < ksteps; ++k)
    mma_f16(d, smem_desc(a.addr + k * a.adv, a.lbo, a.sbo), smem_desc(b.addr + k * b.adv, b.lbo, b.sbo), idesc,
            (accumulate || k > 0) ? 1u : 0u);
}

struct DescW {
  uint32_t lo, hi, adv;                                   // descriptor words; adv = start-address step per K=16, in 16 B units
};
__device__ __forceinline__ DescW desc_words(const Opnd &o) {
  DescW d;
  d.lo = ((o.addr & 0x3FFFFu) >> 4) | (((o.lbo >> 4) & 0x3FFFu) << 16);
  d.hi = ((o.sbo >> 4) & 0x3FFFu) | (1u << 14);
  d.adv = o.adv >> 4;
  return d;
}
__device__ __forceinline__ void issue3(uint32_t d, const Opnd &a, const Opnd &b, uint32_t idesc, int ksteps, bool accumulate) {
  const DescW da = desc_words(a), db = desc_words(b);
#pragma unroll
  for (int k = 0; k < ksteps; ++k)
    mma_f16(d, ((uint64_t)da.hi << 32) | (uint64_t)(da.lo + k * da.adv), ((uint64_t)db.hi << 32) | (uint64_t)(db.lo + k * db.adv),
            idesc, (accumulate || k > 0) ? 1u : 0u);
}

// Who issues the MMAs in the single-role kernels: warp 0 enters the issue block as a whole (a warp-uniform branch, so the
// descriptor arithmetic stays on the uniform datapath — from a divergent `threadIdx.x == 0` branch each tcgen05.mma
// cost ~80 cycles of R2UR traffic, clock64 timeline of round 2) and one elected lane issues.
struct Issuer {
  bool warp0, lead;
};
__device__ __forceinline__ Issuer make_issuer() {
  Issuer iw;
  iw.warp0 = uniform_warp_idx() == 0;
  iw.lead = elect_one() && iw.warp0;
  return iw;
}
#define PN_ISSUE_BEGIN(iw) if ((iw).warp0) { fence_after_sync(); if ((iw).lead) {
#define PN_ISSUE_END(iw, bar) mma_commit(bar); } __syncwarp(); }

// Thread mapping: 256 threads per 128-point tile.  Thread (p = tid & 127, half = tid >> 7) owns row p and
// half of the columns of every epilogue; warps w and w+4 share TMEM lane quarter w (a warp may only touch
// lanes 32*(warp%4) .. +31), so both halves read the same accumulator rows, different columns.
constexpr int kTcThreads = 256;

// tile inputs: features -> A0 (16 columns = 8 levels per thread), SH -> CIN[0..16) (half 0).
//   SRC_F32 : fp32 rows from global (pn_mlp_*_bf16)
//   SRC_HASH: evaluated here from the point coordinates — 8 levels x 8 gathers per thread — and, when
//             F.featb is set, also saved as bf16 tiles for the backward (pn_field_fwd_bf16)
//   SRC_TILE: bf16 tile saved by the forward, copied as is (pn_field_bwd_bf16)
//   SRC_PACKED: as SRC_HASH with the tables held as u8/u16 codes, decoded in the gather (pn_field_fwd_bf16_packed)
template <int SRC, bool GQ = true>
__device__ __forceinline__ void tc_load_inputs(uint8_t *sm, const TcArgs &A, const FieldArgs *F, int64_t tile,
                                               int64_t base, int p, int half, bool valid) {
  if (SRC == SRC_F32) {
    float v[8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int c8 = half * 2 + c;
      if (valid) {
        const float4 *src = reinterpret_cast<const float4 *>(A.in.feat + (base + p) * A.in.feat_stride + c8 * 8);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      st_chunk(sm + TS::A0, chunk_off(p, c8, 4), v);
    }
  } else if (SRC == SRC_TILE) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t off = chunk_off(p, half * 2 + c, 4);
      *reinterpret_cast<uint4 *>(sm + TS::A0 + off) = __ldg(F->featb + tile * 512 + (off >> 4));
    }
  } else {
    float xv[3] = {0.f, 0.f, 0.f};
    if (valid) { xv[0] = __ldg(F->pts + 3 * (base + p)); xv[1] = __ldg(F->pts + 3 * (base + p) + 1); xv[2] = __ldg(F->pts + 3 * (base + p) + 2); }
    // rolled (x2) loop over this thread's 8 levels, each level's two features stored straight into the operand tile
    // (4 bytes at chunk l/4, slot l%4): keeps one level's state live instead of 16 features + 8 levels of code
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
      const int l = half * 8 + i;
      float f0 = 0.f, f1 = 0.f;
      if (l < F->G.n_levels && valid && !(F->debug & 2)) {
        Cell c;
        point_cell<false>(F->G, l, xv, c);
        float e0[8], e1[8];
        if (SRC == SRC_PACKED) {
packed_gather8<false>(F->PK, l, F->G, c, e0, e1);
        } else {
          gather8<false>(F->G, F->T.t[l], c, e0, e1);
        }
        if (GQ && SRC == SRC_HASH && F->qparams) {
          const float *q = F->qparams + l * PN_QROW;
          if (q[5] != 0.f) {
            const float scale = q[0], rdenom = 1.0f / q[1], zp = q[2], qmin = q[3], qmax = q[4];
            const bool train_form = q[6] != 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              e0[k] = fake_quant_fast(e0[k], scale, rdenom, zp, qmin, qmax, train_form);
              e1[k] = fake_quant_fast(e1[k], scale, rdenom, zp, qmin, qmax, train_form);
            }
          }
        }
        f0 = trilerp_fast(e0, c.w);
        f1 = trilerp_fast(e1, c.w);
        if (SRC == SRC_PACKED && F->PK.eb[l] != 8) {        // interpolated in the code domain: one scale per feature
          f0 *= F->PK.scale[l];
          f1 *= F->PK.scale[l];
        }
      }
      *reinterpret_cast<uint32_t *>(sm + TS::A0 + chunk_off(p, l >> 2, 4) + (l & 3) * 4) = pack_bf16(f0, f1);
    }
    if (SRC == SRC_HASH && F->featb) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t off = chunk_off(p, half * 2 + c, 4);
        F->featb[tile * 512 + (off >> 4)] = *reinterpret_cast<const uint4 *>(sm + TS::A0 + off);
      }
    }
    if (half == 0 && valid && F->keep_out) F->keep_out[base + p] = point_keep(F->G, xv) ? 1 : 0;
  }
  if (half == 0) {
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = 0.f;
    if (valid) {
      if (A.in.sh) {
        const float4 *src = reinterpret_cast<const float4 *>(A.in.sh + (base + p) * A.in.sh_stride);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = __ldg(src + q);
          o[4 * q] = a.x; o[4 * q + 1] = a.y; o[4 * q + 2] = a.z; o[4 * q + 3] = a.w;
        }
      } else {
        const int64_t r = (base + p) / A.in.samples_per_ray;
        sh4(__ldg(A.in.dirs + 3 * r), __ldg(A.in.dirs + 3 * r + 1), __ldg(A.in.dirs + 3 * r + 2), o);
      }
    }
    st_chunk(sm + TS::CIN, chunk_off(p, 0, 4), o);
    st_chunk(sm + TS::CIN, chunk_off(p, 1, 4), o + 8);
  }
}

// epilogue of a 64-wide hidden layer, this thread's 32 columns: TMEM -> ReLU -> (fake-quant) -> bf16 tile row;
// returns the ReLU mask of those columns.  The activation fake-quant (IEEE division, ~40 instructions per element) sits
// behind ONE branch: written as `if (qrow)` per element it was if-converted, and the clock64 timeline showed the first
// epilogue of every tile taking 2.4-2.8 k cycles (the others 0.2-0.6 k) with quantisation off.
// RCP: quantise with fake_quant_rcp (no IEEE division off the rounding ties; same bits).  In the three-role backward the
// extra code of its 32 fallback branches cost the (register-tight) epilogue role 0.3 ms per fine pass even with
// quantisation OFF (5.48 -> 5.79 ms, A/B on one box): hence that kernel's ACTQ template parameter — builds without an
// activation quantiser carry no quantiser code at all, builds with one use this form.
template <bool RCP = false>
__device__ __forceinline__ uint32_t epi_hidden32(uint32_t taddr, uint8_t *tile, int p, int half, const float *qrow) {
  uint32_t mask = 0;
  float v[32];
  tmem_ld16(taddr + half * 32, v);
  tmem_ld16(taddr + half * 32 + 16, v + 16);
  tmem_ld_wait();
  // four dependent shift-or chains of 8 bits: written as 32 independent `mask |= bit << j` terms, ptxas kept the terms
  // live and spilled 24 of them to local memory (which, with 2 x 107 KB of shared memory per SM, has almost no L1)
  uint32_t m4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      const bool pos = v[8 * c + j] > 0.f;
      m4[c] = (m4[c] << 1) | (pos ? 1u : 0u);
      v[8 * c + j] = fmaxf(v[8 * c + j], 0.f);
    }
  }
  mask = m4[0] | (m4[1] << 8) | (m4[2] << 16) | (m4[3] << 24);
  if (qrow != nullptr) {
    const float scale = qrow[0], denom = qrow[1], zp = qrow[2], qmin = qrow[3], qmax = qrow[4];
    const bool train_form = qrow[6] != 0.f;
    if (RCP) {
      const float rdenom = pn_div(1.0f, denom);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fake_quant_rcp(v[j], scale, denom, rdenom, zp, qmin, qmax, train_form);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fake_quant(v[j], scale, denom, zp, qmin, qmax, train_form);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) st_chunk(tile, chunk_off(p, half * 4 + c, 8), v + 8 * c);
  return mask;
}

// barrier over the 256 threads that run the MLP rounds: the whole CTA in the plain kernels, warps 0-7 in the
// warp-specialised ones (whose scatter / gather warps never take part)
__device__ __forceinline__ void mlp_sync() { named_bar_sync<1, 256>(); }
#define PN_ROUND_SYNC() do { fence_async_smem(); fence_before_sync(); mlp_sync(); } while (0)

// forward rounds shared by the forward kernel and the backward's recompute.
// BWD = false: every 64-wide activation goes to A1.   BWD = true: H1 -> A1, colour hidden 1 -> A1C, 2 -> A2C.
template <bool BWD>
__device__ __forceinline__ void tc_forward(uint8_t *sm, const TcArgs &A, uint32_t tmem, uint32_t lane_addr, uint64_t *bar,
                                           uint32_t &ph, int p, int half, const float *qrow, float &sigma,
                                           float nraw[3], uint32_t &h1_mask, const Issuer &iw, TLog *tl = nullptr) {
  TLog none = {nullptr, 0, 0};
  TLog &T = tl ? *tl : none;
  uint8_t *a1c = BWD ? sm + TS::A1C : sm + TS::A1;
  uint8_t *a2c = BWD ? sm + TS::A2C : sm + TS::A1;
  uint8_t *a0 = sm + TS::A0, *cin = sm + TS::CIN;
  // R1: H1 = relu(X S0^T)
  T.mark();
  PN_ISSUE_BEGIN(iw)
    issue3(tmem + TM_D1, k_major(a0, 32), k_major(sm + TS::W_S0, 32), instr_desc(128, 64, 0, 0), 2, false);
  PN_ISSUE_END(iw, bar)
  T.mark();
  mbar_wait(bar, ph); ph ^= 1; fence_after_sync();
  T.mark();
  h1_mask = epi_hidden32<!BWD>(lane_addr + TM_D1, sm + TS::A1, p, half, qrow);
  T.mark();
  PN_ROUND_SYNC();
  // R2: [sigma, geo] = H1 S1^T
  T.mark();
  PN_ISSUE_BEGIN(iw)
    issue3(tmem + TM_D2, k_major(sm + TS::A1, 64), k_major(sm + TS::W_S1, 64), instr_desc(128, 16, 0, 0), 4, false);
  PN_ISSUE_END(iw, bar)
  T.mark();
  mbar_wait(bar, ph); ph ^= 1; fence_after_sync();
  T.mark();
  if (half == 0) {
    float v[17];
    tmem_ld16(lane_addr + TM_D2, v);
    tmem_ld_wait();
    sigma = v[0];
    v[16] = 0.f;
    st_chunk(cin, chunk_off(p, 2, 4), v + 1);      // geo 0..7
    st_chunk(cin, chunk_off(p, 3, 4), v + 9);      // geo 8..14, 0
  }
  T.mark();
  PN_ROUND_SYNC();
  // R3: A1c = relu(CIN C0^T);  NH = relu(geo N0^T + b)
  T.mark();
  PN_ISSUE_BEGIN(iw)
    issue3(tmem + TM_D1, k_major(cin, 32), k_major(sm + TS::W_C0, 32), instr_desc(128, 64, 0, 0), 2, false);
    if (A.normals)
      issue3(tmem + TM_DN, k_major(cin, 32, 16), k_major(sm + TS::W_N0, 16), instr_desc(128, 32, 0, 0), 1, false);
  PN_ISSUE_END(iw, bar)
  T.mark();
  mbar_wait(bar, ph); ph ^= 1; fence_after_sync();
  T.mark();
  epi_hidden32(lane_addr + TM_D1, a1c, p, half, nullptr);
  if (A.normals) {
    const float *bias = reinterpret_cast<const float *>(sm + TS::BIAS) + half * 16;
    float v[16];
    tmem_ld16(lane_addr + TM_DN + half * 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + bias[j], 0.f);
    st_chunk(sm + TS::NH, chunk_off(p, half * 2, 4), v);
    st_chunk(sm + TS::NH, chunk_off(p, half * 2 + 1, 4), v + 8);
  }
  T.mark();
  PN_ROUND_SYNC();
  // R4: A2c = relu(A1c C1^T);  raw normal = NH N2^T + b
  T.mark();
  PN_ISSUE_BEGIN(iw)
    issue3(tmem + TM_D1, k_major(a1c, 64), k_major(sm + TS::W_C1, 64), instr_desc(128, 64, 0, 0), 4, false);
    if (A.normals)
      issue3(tmem + TM_D2, k_major(sm + TS::NH, 32), k_major(sm + TS::W_N2, 32), instr_desc(128, 16, 0, 0), 2, false);
  PN_ISSUE_END(iw, bar)
  T.mark();
  mbar_wait(bar, ph); ph ^= 1; fence_after_sync();
  T.mark();
  if (A.normals && half == 0) {
    const float *bias = reinterpret_cast<const float *>(sm + TS::BIAS) + 32;
    float v[16];
    tmem_ld16(lane_addr + TM_D2, v);
    tmem_ld_wait();
    nraw[0] = v[0] + bias[0]; nraw[1] = v[1] + bias[1]; nraw[2] = v[2] + bias[2];
  }
  epi_hidden32(lane_addr + TM_D1, a2c, p, half, nullptr);
  T.mark();
  PN_ROUND_SYNC();
}

// QUANT = false: a build without any fake-quant code (no per-level rows in the gather, no activation quantiser) — what
// every unquantised model runs, and quantised ones too unless the rows are passed to the gather (PN_QUANT_IN_GATHER)
// or the MLP has an activation quantiser.
template <int MIN_CTAS, int SRC, bool QUANT = true>
__global__ void __launch_bounds__(kTcThreads, MIN_CTAS)
mlp_tc_fwd_kernel(const TcArgs A, const __grid_constant__ FieldArgs F, float *__restrict__ out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, p = tid & 127, half = tid >> 7, warp = tid >> 5;
  load_all_weights(sm, A);
  if (warp == 0) tmem_alloc(&tmem_slot, TM_FWD_COLS);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const Issuer iw = make_issuer();
  uint32_t ph = 0;
  float q[8];
  const float *qrow = nullptr;
  if (QUANT && A.in.act_q) {
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = __ldg(A.in.act_q + i);
    if (q[5] != 0.f) qrow = q;
  }
  const int64_t n_tiles = (A.in.n_points + kTcTile - 1) / kTcTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kTcTile;
    const bool valid = base + p < A.in.n_points;
    // the next tile's positions -> L2 while this one runs (SRC_HASH / SRC_PACKED read them right at the start of a tile)
    if ((SRC == SRC_HASH || SRC == SRC_PACKED) && iw.lead && tile + gridDim.x < n_tiles) {
      const int64_t nb = (tile + gridDim.x) * kTcTile;
      const int64_t rows = (A.in.n_points - nb) < kTcTile ? (A.in.n_points - nb) : kTcTile;
      const uint32_t pb = (uint32_t)(rows * 12) & ~15u;
      if (pb && (((uintptr_t)(F.pts + nb * 3)) & 15) == 0) prefetch_l2(F.pts + nb * 3, pb);
    }
    tc_load_inputs<SRC, QUANT>(sm, A, &F, tile, base, p, half, valid);
    PN_ROUND_SYNC();
    float sigma = 0.f, nraw[3] = {0.f, 0.f, 0.f};
    uint32_t m;
    tc_forward<false>(sm, A, tmem, lane_addr, &bar, ph, p, half, qrow, sigma, nraw, m, iw);
    // R5: rgb = A2c C2^T
    PN_ISSUE_BEGIN(iw)
      issue3(tmem + TM_D2, k_major(sm + TS::A1, 64), k_major(sm + TS::W_C2, 64), instr_desc(128, 16, 0, 0), 4, false);
    PN_ISSUE_END(iw, &bar)
    mbar_wait(&bar, ph); ph ^= 1; fence_after_sync();
    if (half == 0) {
      float v[16];
      tmem_ld16(lane_addr + TM_D2, v);
      tmem_ld_wait();
      if (valid) {
        bool kept = A.in.keep ? (A.in.keep[base + p] != 0) : true;
        if (SRC == SRC_HASH || SRC == SRC_PACKED) {
          const float xv[3] = {__ldg(F.pts + 3 * (base + p)), __ldg(F.pts + 3 * (base + p) + 1), __ldg(F.pts + 3 * (base + p) + 2)};
          kept = point_keep(F.G, xv);
        }
        float *o = out + (base + p) * A.C;
        if (A.C == 4) {
          *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], kept ? sigma : 0.f);
        } else {
          const float nn = fmaxf(sqrtf(nraw[0] * nraw[0] + nraw[1] * nraw[1] + nraw[2] * nraw[2]), 1e-12f);
          o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = sigma;
          o[4] = nraw[0] / nn; o[5] = nraw[1] / nn; o[6] = kept ? nraw[2] / nn : 0.f;
        }
      }
    }
    fence_before_sync(); __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, TM_FWD_COLS);
}

// masked in-place epilogue of an input-gradient GEMM, this thread's 32 columns:
// tile row <- D * [tile row > 0] (or an explicit mask)
__device__ __forceinline__ void epi_grad32(uint32_t taddr, uint8_t *tile, int p, int half, bool use_mask, uint32_t mask) {
  float v[32];
  tmem_ld16(taddr + half * 32, v);
  tmem_ld16(taddr + half * 32 + 16, v + 16);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t off = chunk_off(p, half * 4 + c, 8);
    if (use_mask) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * c + j] = ((mask >> (c * 8 + j)) & 1u) ? v[8 * c + j] : 0.f;
    } else {
      float a[8];
      ld_chunk(tile, off, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * c + j] = a[j] > 0.f ? v[8 * c + j] : 0.f;
    }
    st_chunk(tile, off, v + 8 * c);
  }
}

__device__ __forceinline__ float tile_elem(const uint8_t *tile, int r, int c, int C8) {
  const uint16_t h = *reinterpret_cast<const uint16_t *>(tile + chunk_off(r, c >> 3, C8) + (c & 7) * 2);
  return __uint_as_float((uint32_t)h << 16);
}

// accumulator rows held by this thread (M = 64 layout): valid when lane < 16, row = 16*warp + lane
__device__ __forceinline__ void flush_acc(uint32_t taddr, int ncols, bool owner, int row, float *dst, int pitch,
                                          int row_valid, int col_valid, bool transposed) {
  for (int c = 0; c < ncols; c += 16) {
    float v[16];
    tmem_ld16(taddr + c, v);
    tmem_ld_wait();
    if (owner && dst && row < row_valid) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c + j;
        if (col < col_valid) atomicAdd(transposed ? dst + col * pitch + row : dst + row * pitch + col, v[j]);
      }
    }
  }
}

template <int SRC>
__global__ void __launch_bounds__(kTcThreads, 2)
mlp_tc_bwd_kernel(const TcArgs A, const __grid_constant__ FieldArgs F, const float *__restrict__ dout,
                  float *__restrict__ dfeat, int64_t dfeat_stride, float *__restrict__ dsh, int64_t dsh_stride,
                  const pn_mlp_grads G) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, p = tid & 127, half = tid >> 7, warp = tid >> 5, lane = tid & 31;
  load_all_weights(sm, A);
  if (warp == 0) tmem_alloc(&tmem_slot, TM_BWD_COLS);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int64_t n_tiles = (A.in.n_points + kTcTile - 1) / kTcTile;
  uint32_t ph = 0, n_done = 0;
  TLog T = {nullptr, 0, F.tlog_cap};
  if (F.tlog && blockIdx.x == 0 && (tid == 0 || tid == 160)) T.p = F.tlog + (tid ? F.tlog_cap : 0);
  float q[8];
  const float *qrow = nullptr;
  if (A.in.act_q) {
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = __ldg(A.in.act_q + i);
    if (q[5] != 0.f) qrow = q;
  }
  // normal-head weight gradients are reduced on the CUDA cores (tiny: 611 numbers), fp32 registers across tiles
  float g_n2 = 0.f, g_n0[4] = {0.f, 0.f, 0.f, 0.f};
  const Issuer iw = make_issuer();
  bool first = true;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kTcTile;
    const bool valid = base + p < A.in.n_points;
    T.mark();
    tc_load_inputs<SRC>(sm, A, &F, tile, base, p, half, valid);
    float d_o[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (valid) {
      for (int c = 0; c < A.C; ++c) d_o[c] = __ldg(dout + (base + p) * A.C + c);
      if (A.in.keep && A.in.keep[base + p] == 0) d_o[A.C - 1] = 0.f;            // run_nerf.py:66
    }
    T.mark();
    PN_ROUND_SYNC();
    float sigma = 0.f, nraw[3] = {0.f, 0.f, 0.f};
    uint32_t h1_mask;
    tc_forward<true>(sm, A, tmem, lane_addr, &bar, ph, p, half, qrow, sigma, nraw, h1_mask, iw, &T);

    // B0: cotangent tiles (half 0 owns the row-level values)
    if (half == 0) {
      float v[8] = {d_o[0], d_o[1], d_o[2], 0.f, 0.f, 0.f, 0.f, 0.f}, z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      st_chunk(sm + TS::DOUT, chunk_off(p, 0, 2), v);
      st_chunk(sm + TS::DOUT, chunk_off(p, 1, 2), z);
      if (A.normals) {
        const float nn = sqrtf(nraw[0] * nraw[0] + nraw[1] * nraw[1] + nraw[2] * nraw[2]);
        float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (nn > 1e-12f) {
          const float m0 = nraw[0] / nn, m1 = nraw[1] / nn, m2 = nraw[2] / nn;
          const float dot = m0 * d_o[4] + m1 * d_o[5] + m2 * d_o[6];
          r[0] = (d_o[4] - m0 * dot) / nn; r[1] = (d_o[5] - m1 * dot) / nn; r[2] = (d_o[6] - m2 * dot) / nn;
        } else {
          r[0] = d_o[4] / 1e-12f; r[1] = d_o[5] / 1e-12f; r[2] = d_o[6] / 1e-12f;
        }
        st_chunk(sm + TS::DNR, chunk_off(p, 0, 2), r);
        st_chunk(sm + TS::DNR, chunk_off(p, 1, 2), z);
      }
    }
    T.mark();
    PN_ROUND_SYNC();
    // B1: dC2^T += A2c^T dOut ; dA2 = dOut C2 ; (normals) dNH = dNraw N2
    T.mark();
    PN_ISSUE_BEGIN(iw)
      issue3(tmem + TM_GC2, mn_major(sm + TS::A2C, 64), mn_major(sm + TS::DOUT, 16), instr_desc(64, 16, 1, 1), 8, !first);
      issue3(tmem + TM_D1, k_major(sm + TS::DOUT, 16), mn_major(sm + TS::W_C2, 64), instr_desc(128, 64, 0, 1), 1, false);
      if (A.normals)
        issue3(tmem + TM_DN, k_major(sm + TS::DNR, 16), mn_major(sm + TS::W_N2, 32), instr_desc(128, 32, 0, 1), 1, false);
    PN_ISSUE_END(iw, &bar)
    if (A.normals && tid < 99) {                   // dN2w[j][k] / dN2b[j] on the CUDA cores (needs NH before E1 overwrites it)
      const int j = tid < 96 ? tid >> 5 : tid - 96, k = tid & 31;
      float s = 0.f;
      for (int r = 0; r < kTcTile; ++r)
        s += tile_elem(sm + TS::DNR, r, j, 2) * (tid < 96 ? tile_elem(sm + TS::NH, r, k, 4) : 1.f);
      g_n2 += s;
    }
    T.mark();
    mbar_wait(&bar, ph); ph ^= 1; fence_after_sync();
    T.mark();
    if (A.normals) mlp_sync();                     // all NH reads above are done
    epi_grad32(lane_addr + TM_D1, sm + TS::A2C, p, half, false, 0);
    if (A.normals) {
      float v[16];
      tmem_ld16(lane_addr + TM_DN + half * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float a[8];
        ld_chunk(sm + TS::NH, chunk_off(p, half * 2 + c, 4), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * c + j] = a[j] > 0.f ? v[8 * c + j] : 0.f;
        st_chunk(sm + TS::NH, chunk_off(p, half * 2 + c, 4), v + 8 * c);
      }
    }
    T.mark();
    PN_ROUND_SYNC();
    // B2: dC1 += dA2pre^T A1c ; dA1 = dA2pre C1
    T.mark();
    PN_ISSUE_BEGIN(iw)
      issue3(tmem + TM_GC1, mn_major(sm + TS::A2C, 64), mn_major(sm + TS::A1C, 64), instr_desc(64, 64, 1, 1), 8, !first);
      issue3(tmem + TM_D1, k_major(sm + TS::A2C, 64), mn_major(sm + TS::W_C1, 64), instr_desc(128, 64, 0, 1), 4, false);
    PN_ISSUE_END(iw, &bar)
    if (A.normals && tid < 128) {                  // dN0w[j][k] (k < 15) and dN0b[j] (k == 15): 4 outputs per thread
      const int j = tid >> 2, k0 = (tid & 3) * 4;
      for (int r = 0; r < kTcTile; ++r) {
        const float d = tile_elem(sm + TS::NH, r, j, 4);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          g_n0[i] += d * ((k0 + i) < 15 ? tile_elem(sm + TS::CIN, r, 16 + k0 + i, 4) : 1.f);
      }
    }
    T.mark();
    mbar_wait(&bar, ph); ph ^= 1; fence_after_sync();
    T.mark();
    epi_grad32(lane_addr + TM_D1, sm + TS::A1C, p, half, false, 0);
    T.mark();
    PN_ROUND_SYNC();
    // B3: dC0 += dA1pre^T CIN ; dCIN = dA1pre C0 ; (normals) dgeo_n = dNHpre N0
    T.mark();
    PN_ISSUE_BEGIN(iw)
      issue3(tmem + TM_GC0, mn_major(sm + TS::A1C, 64), mn_major(sm + TS::CIN, 32), instr_desc(64, 32, 1, 1), 8, !first);
      issue3(tmem + TM_D1, k_major(sm + TS::A1C, 64), mn_major(sm + TS::W_C0, 32), instr_desc(128, 32, 0, 1), 4, false);
      if (A.normals)
        issue3(tmem + TM_D2, k_major(sm + TS::NH, 32), mn_major(sm + TS::W_N0, 16), instr_desc(128, 16, 0, 1), 2, false);
    PN_ISSUE_END(iw, &bar)
    T.mark();
    mbar_wait(&bar, ph); ph ^= 1; fence_after_sync();
    T.mark();
    if (half == 0) {
      if (dsh) {                                       // dCIN[0..16) = d SH
        float v[16];
        tmem_ld16(lane_addr + TM_D1, v);
        tmem_ld_wait();
        if (valid) {
          float4 *o = reinterpret_cast<float4 *>(dsh + (base + p) * dsh_stride);
#pragma unroll
          for (int c = 0; c < 4; ++c) o[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
      }
    } else {
      float g[17];
      tmem_ld16(lane_addr + TM_D1 + 16, g + 1);        // dCIN[16..32) = dgeo[0..15) + pad
      tmem_ld_wait();
      if (A.normals) {
        float v[16];
        tmem_ld16(lane_addr + TM_D2, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 15; ++j) g[1 + j] += v[j];
      }
      g[0] = d_o[3];                                   // dsigma (keep mask applied above when C == 4)
      st_chunk(sm + TS::DH2, chunk_off(p, 0, 2), g);
      st_chunk(sm + TS::DH2, chunk_off(p, 1, 2), g + 8);
    }
    T.mark();
    PN_ROUND_SYNC();
    // B4: dS1^T += H1^T dH2 ; dH1 = dH2 S1
    T.mark();
    PN_ISSUE_BEGIN(iw)
      issue3(tmem + TM_GS1, mn_major(sm + TS::A1, 64), mn_major(sm + TS::DH2, 16), instr_desc(64, 16, 1, 1), 8, !first);
      issue3(tmem + TM_D1, k_major(sm + TS::DH2, 16), mn_major(sm + TS::W_S1, 64), instr_desc(128, 64, 0, 1), 1, false);
    PN_ISSUE_END(iw, &bar)
    T.mark();
    mbar_wait(&bar, ph); ph ^= 1; fence_after_sync();
    T.mark();
    epi_grad32(lane_addr + TM_D1, sm + TS::A1, p, half, true, h1_mask);
    T.mark();
    PN_ROUND_SYNC();
    // B5: dS0 += dH1pre^T X ; dX = dH1pre S0
    T.mark();
    PN_ISSUE_BEGIN(iw)
      issue3(tmem + TM_GS0, mn_major(sm + TS::A1, 64), mn_major(sm + TS::A0, 32), instr_desc(64, 32, 1, 1), 8, !first);
      issue3(tmem + TM_D1, k_major(sm + TS::A1, 64), mn_major(sm + TS::W_S0, 32), instr_desc(128, 32, 0, 1), 4, false);
    PN_ISSUE_END(iw, &bar)
    T.mark();
    mbar_wait(&bar, ph); ph ^= 1; fence_after_sync();
    T.mark();
    if (SRC == SRC_TILE) {
      // Fused scatter.  A warp's 32 lanes are 32 consecutive samples, so the run-aggregated scatter applies
      // unchanged.  Half 0 scatters levels [0, split), half 1 the rest; each level's two gradients are read from
      // TMEM inside the loop so that the loop stays rolled (16 inlined copies of the scatter code thrashed the
      // instruction cache: 1.4x slower; the rolled loop is 5 % faster than registers + 8 copies).  Measured step time
      // vs split on B200: 4: 15.56 ms, 6: 15.43, 7: 15.17, 8: 15.14, 10: 15.59, 12: 15.97 -> default 8.
      const int kSplit = F.scatter_split;
      float xv[3] = {0.f, 0.f, 0.f};
      if (valid) { xv[0] = __ldg(F.pts + 3 * (base + p)); xv[1] = __ldg(F.pts + 3 * (base + p) + 1); xv[2] = __ldg(F.pts + 3 * (base + p) + 2); }
      const int l0 = half ? kSplit : 0,
                l1 = (F.debug & 1) ? 0 : (half ? F.G.n_levels : (kSplit < F.G.n_levels ? kSplit : F.G.n_levels));
#pragma unroll 2
      for (int l = l0; l < l1; ++l) {
        float g0, g1;
        tmem_ld2(lane_addr + TM_D1 + 2 * l, g0, g1);
        tmem_ld_wait();
        scatter_level<false>(F.G, F.D.t[l], l, xv, valid ? g0 : 0.f, valid ? g1 : 0.f, lane);
      }
    } else {
      float v[16];
      tmem_ld16(lane_addr + TM_D1 + half * 16, v);
      tmem_ld_wait();
      if (valid) {
        float4 *o = reinterpret_cast<float4 *>(dfeat + (base + p) * dfeat_stride + half * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
    }
    T.mark();
    first = false;
    ++n_done;
    fence_before_sync(); mlp_sync();
  }
  // flush the weight gradients (every MMA has completed: the last commit was waited on)
  fence_after_sync();
  if (!first && warp < 4) {
    const bool owner = lane < 16;
    const int row = warp * 16 + lane;
    flush_acc(lane_addr + TM_GC1, 64, owner, row, G.c1, 64, 64, 64, false);
    flush_acc(lane_addr + TM_GS0, 32, owner, row, G.s0, 32, 64, 32, false);
    flush_acc(lane_addr + TM_GC0, 32, owner, row, G.c0, 31, 64, 31, false);
    flush_acc(lane_addr + TM_GS1, 16, owner, row, G.s1, 64, 64, 16, true);
    flush_acc(lane_addr + TM_GC2, 16, owner, row, G.c2, 64, 64, 3, true);
  }
  if (!first && A.normals) {
    if (tid < 96) { if (G.n2w) atomicAdd(G.n2w + (tid >> 5) * 32 + (tid & 31), g_n2); }
    else if (tid < 99) { if (G.n2b) atomicAdd(G.n2b + (tid - 96), g_n2); }
    if (tid < 128) {
      const int j = tid >> 2, k0 = (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (k0 + i < 15) { if (G.n0w) atomicAdd(G.n0w + j * 15 + k0 + i, g_n0[i]); }
        else if (G.n0b) atomicAdd(G.n0b + j, g_n0[i]);
      }
    }
  }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TM_BWD_COLS);
}


// ------------------------------------------------------------------------------------------------------
// Helpers of the three-role backward below: a dedicated MMA warp and per-warp mbarrier hand-offs.
// What the clock64 timeline of the single-role kernel above showed per 128-point tile (33 k cycles,
// profiles/r02_timeline_bwd_v1_ws_clock64.txt): 3.0 k waiting for the saved feature tile and 2.4 k inside the first
// epilogue (an if-converted fake-quant, since fixed), 6 k in thread 0 building descriptors and issuing the 78 MMAs of a
// tile (~80 cycles each from a divergent branch), 16 k in the scatter; the MMAs themselves and the epilogues are ~1 k and
// ~2 k.  In the three-role kernel the MMA warp issues from warp-uniform control flow with precomputed descriptor words
// (uniform datapath, ~20-40 cycles per MMA), prefetches the next tile's inputs into L2, and the epilogue warps never
// issue or meet at a CTA barrier: they arrive on `ready` (one elected lane per warp) and go straight to waiting on `done`.
// ------------------------------------------------------------------------------------------------------

// epilogue warps: this round's operand tiles are written -> one arrival per warp on `ready`
__device__ __forceinline__ void epi_arrive(uint64_t *ready, int lane) {
  fence_async_smem();
  fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(ready);
}
__device__ __forceinline__ void epi_wait(uint64_t *done, uint32_t &ph) {
  mbar_wait(done, ph);
  ph ^= 1;
  fence_after_sync();
}


// ------------------------------------------------------------------------------------------------------
// Fused backward (pn_field_bwd_bf16 default): three roles per CTA, 2 CTAs per SM, with the scatter DECOUPLED from the
// round chain through a small ring in global memory (L2-resident: kRingSlots x 16 KB per CTA).
//   warps 0-7   epilogue; after B5 each thread copies its half row of dX (TMEM, fp32) into the CTA's ring slot
//   warp  8     MMA issuer
//   warps 9-15  scatter: 7 warps take (tile, 4 levels) items round-robin from the ring — any warp can take any item
//               (the TMEM lane-quarter rule no longer applies), and they may lag the chain by kRingSlots tiles; a lane
//               walks 4 consecutive samples serially per level (scatter_segment: run aggregation without shuffles).
// Why: with the training step's own gradients nearly every warp of 32 samples has some non-zero rows, so the scatter
// costs its full ~230 instructions per level per warp; 4 scatter warps tied to the chain by a one-tile TMEM hand-off
// (an earlier structure of this round) took ~30 k cycles per tile against ~13 k for the chain.  What the scatter needs is issue slots from many warps
// (the standalone scatter kernel reaches 71 % issue utilisation with 32 warps/SM), not a place in the chain.
// ------------------------------------------------------------------------------------------------------
#ifndef PN_V4_SCATTER_WARPS
#define PN_V4_SCATTER_WARPS 7
#endif
constexpr int kV4ScatterWarps = PN_V4_SCATTER_WARPS;                    // 3 (12 warps per CTA) or 7 (16 warps)
constexpr int kV4Threads = kTcThreads + 32 + 32 * kV4ScatterWarps;
// register budget at 2 CTAs/SM (setmaxnreg per warpgroup): 12 warps -> launch bound 80: 256*88 + 128*64 = 384*80;
//                                                           16 warps -> launch bound 64: 256*80 + 256*48 = 256*72 + 256*56 = 512*64
// The split between the roles depends on the build: with the activation quantiser compiled in, the epilogue role needs its
// 80 registers; without it 72 are enough and the 8 freed per thread take the scatter role from 48 to 56, where it no
// longer spills (fine-pass backward 5.66 -> 5.57 ms kernel-only; 88 / 40 was 7.25 ms).
template <bool ACTQ>
struct V4Regs {
#ifndef PN_V4_EPI_NOQ
#define PN_V4_EPI_NOQ 72
#endif
  static constexpr int epi = kV4ScatterWarps == 3 ? 88 : (ACTQ ? 80 : PN_V4_EPI_NOQ);
  static constexpr int aux = kV4ScatterWarps == 3 ? 64 : (ACTQ ? 48 : 128 - PN_V4_EPI_NOQ);
};
constexpr int kRingSlots = 4;
static_assert(kV4ScatterWarps == 3 || kV4ScatterWarps == 7, "warps 8.. must fill whole warpgroups (setmaxnreg)");

// SEG: which scatter the scatter warps run — true: a lane walks 4 consecutive samples serially per level
// (scatter_segment, ring slot laid out [level][row]); false: a lane owns one sample and runs are summed with the shuffle
// tree (scatter_level, ring slot [row][level]).  Measured in the training step (round 2): with 192 sorted samples per ray
// (fine pass) consecutive samples share voxels at most levels and SEG wins (5.41 -> 5.19 ms); with 64 samples per ray
// (coarse pass) they rarely do and the 4-sample serial loop only costs parallelism (2.25 -> 2.67 ms) — so the launcher
// picks SEG by samples_per_ray.
template <bool NORMALS, bool SEG, bool ACTQ>
__global__ void __launch_bounds__(kV4Threads, 2)
field_bwd4_kernel(const TcArgs A, const __grid_constant__ FieldArgs F, const float *__restrict__ dout, const pn_mlp_grads G,
                  float *__restrict__ ring) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t ready, done, ring_full[kRingSlots], ring_empty[kRingSlots];
  __shared__ uint32_t ring_nz[kRingSlots][4][2];      // per slot / quarter / column half: lanes whose dX half row is non-zero
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = uniform_warp_idx();
  load_all_weights(sm, A);
  if (warp == 0) tmem_alloc(&tmem_slot, TM_BWD_COLS);
  if (tid == 0) {
    mbar_init(&ready, 8); mbar_init(&done, 1);
    for (int i = 0; i < kRingSlots; ++i) { mbar_init(&ring_full[i], 8); mbar_init(&ring_empty[i], 4); }
    mbar_fence_init();
  }
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int64_t n_tiles = (A.in.n_points + kTcTile - 1) / kTcTile;
  const int C = A.C;

  float *const cta_ring = ring + (size_t)blockIdx.x * (kRingSlots * kTcTile * 32);
  if (warp >= 8) setmaxnreg_dec<V4Regs<ACTQ>::aux>();   // warpgroup 8-11 (and 12-15), one instruction
  if (warp == 8) {
    // ---------------- MMA role ----------------
    const bool lead = elect_one();
    uint32_t pr = 0, n = 0;
    bool first = true;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      // next tile's inputs -> L2 (saved feature tile 8 KB, cotangent rows, positions, keep flags)
      const int64_t nt = tile + gridDim.x;
      if (lead && nt < n_tiles) {
        const int64_t nb = nt * kTcTile;
        const int64_t rows = (A.in.n_points - nb) < kTcTile ? (A.in.n_points - nb) : kTcTile;
        prefetch_l2(F.featb + nt * 512, 8192);
        const uint32_t db = (uint32_t)(rows * C * 4) & ~15u, pb = (uint32_t)(rows * 12) & ~15u;
        if (db && (((uintptr_t)(dout + nb * C)) & 15) == 0) prefetch_l2(dout + nb * C, db);
        if (pb && (((uintptr_t)(F.pts + nb * 3)) & 15) == 0) prefetch_l2(F.pts + nb * 3, pb);
      }
#define PN_MMA_ROUND(...)                                         \
  do {                                                            \
    mbar_wait(&ready, pr); pr ^= 1; fence_after_sync();           \
    if (lead) { __VA_ARGS__; mma_commit(&done); }                 \
    __syncwarp();                                                 \
  } while (0)
      PN_MMA_ROUND(issue3(tmem + TM_D1, k_major(sm + TS::A0, 32), k_major(sm + TS::W_S0, 32), instr_desc(128, 64, 0, 0), 2, false));
      PN_MMA_ROUND(issue3(tmem + TM_D2, k_major(sm + TS::A1, 64), k_major(sm + TS::W_S1, 64), instr_desc(128, 16, 0, 0), 4, false));
      PN_MMA_ROUND(issue3(tmem + TM_D1, k_major(sm + TS::CIN, 32), k_major(sm + TS::W_C0, 32), instr_desc(128, 64, 0, 0), 2, false);
                   if (NORMALS) issue3(tmem + TM_DN, k_major(sm + TS::CIN, 32, 16), k_major(sm + TS::W_N0, 16), instr_desc(128, 32, 0, 0), 1, false));
      PN_MMA_ROUND(issue3(tmem + TM_D1, k_major(sm + TS::A1C, 64), k_major(sm + TS::W_C1, 64), instr_desc(128, 64, 0, 0), 4, false);
                   if (NORMALS) issue3(tmem + TM_D2, k_major(sm + TS::NH, 32), k_major(sm + TS::W_N2, 32), instr_desc(128, 16, 0, 0), 2, false));
      // B1: dC2^T += A2c^T dOut ; dA2 = dOut C2 ; (normals) dNH = dNraw N2
      PN_MMA_ROUND(issue3(tmem + TM_GC2, mn_major(sm + TS::A2C, 64), mn_major(sm + TS::DOUT, 16), instr_desc(64, 16, 1, 1), 8, !first);
                   issue3(tmem + TM_D1, k_major(sm + TS::DOUT, 16), mn_major(sm + TS::W_C2, 64), instr_desc(128, 64, 0, 1), 1, false);
                   if (NORMALS) issue3(tmem + TM_DN, k_major(sm + TS::DNR, 16), mn_major(sm + TS::W_N2, 32), instr_desc(128, 32, 0, 1), 1, false));
      // B2: dC1 += dA2pre^T A1c ; dA1 = dA2pre C1
      PN_MMA_ROUND(issue3(tmem + TM_GC1, mn_major(sm + TS::A2C, 64), mn_major(sm + TS::A1C, 64), instr_desc(64, 64, 1, 1), 8, !first);
                   issue3(tmem + TM_D1, k_major(sm + TS::A2C, 64), mn_major(sm + TS::W_C1, 64), instr_desc(128, 64, 0, 1), 4, false));
      // B3: dC0 += dA1pre^T CIN ; dCIN = dA1pre C0 ; (normals) dgeo_n = dNHpre N0
      PN_MMA_ROUND(issue3(tmem + TM_GC0, mn_major(sm + TS::A1C, 64), mn_major(sm + TS::CIN, 32), instr_desc(64, 32, 1, 1), 8, !first);
                   issue3(tmem + TM_D1, k_major(sm + TS::A1C, 64), mn_major(sm + TS::W_C0, 32), instr_desc(128, 32, 0, 1), 4, false);
                   if (NORMALS) issue3(tmem + TM_D2, k_major(sm + TS::NH, 32), mn_major(sm + TS::W_N0, 16), instr_desc(128, 16, 0, 1), 2, false));
      // B4: dS1^T += H1^T dH2 ; dH1 = dH2 S1
      PN_MMA_ROUND(issue3(tmem + TM_GS1, mn_major(sm + TS::A1, 64), mn_major(sm + TS::DH2, 16), instr_desc(64, 16, 1, 1), 8, !first);
                   issue3(tmem + TM_D1, k_major(sm + TS::DH2, 16), mn_major(sm + TS::W_S1, 64), instr_desc(128, 64, 0, 1), 1, false));
      // B5: dS0 += dH1pre^T X ; dX = dH1pre S0
      PN_MMA_ROUND(issue3(tmem + TM_GS0, mn_major(sm + TS::A1, 64), mn_major(sm + TS::A0, 32), instr_desc(64, 32, 1, 1), 8, !first);
                   issue3(tmem + TM_D1, k_major(sm + TS::A1, 64), mn_major(sm + TS::W_S0, 32), instr_desc(128, 32, 0, 1), 4, false));
#undef PN_MMA_ROUND
      first = false;
    }
    __syncthreads();
    return;
  }

  if (warp >= 9) {
    if (SEG) {
    // ---------------- scatter role: kV4ScatterWarps warps; work item = (tile, 4 levels), taken round-robin -------------
    // A lane owns 4 CONSECUTIVE samples of the tile (rows 4*lane .. +3) and walks them serially per level
    // (scatter_segment): consecutive samples of a ray that share a voxel are summed in registers, no shuffles.
    const int sw = warp - 9;
    const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;      // tiles this CTA walks
    for (int64_t item = sw; item < my_tiles * 4; item += kV4ScatterWarps) {
      const int64_t n = item >> 2;
      const int lq = (int)(item & 3), slot = (int)(n % kRingSlots);
      const int64_t base = (blockIdx.x + n * gridDim.x) * kTcTile + 4 * lane;
      const int64_t left = A.in.n_points - base;
      const int ns = left >= 4 ? 4 : (left > 0 ? (int)left : 0);
      mbar_wait(&ring_full[slot], (uint32_t)(n / kRingSlots) & 1);
      // rows of this lane with a gradient (published by the epilogue warps with the tile; rows past n_points are zero)
      const uint32_t mq = ring_nz[slot][lane >> 3][0] | ring_nz[slot][lane >> 3][1];
      const uint32_t mine = (mq >> ((lane & 7) * 4)) & 0xFu;
      if (mine != 0u && !(F.debug & 1)) {
        // ring slot layout [level][row] float2: this lane's 4 rows of one level are 32 contiguous bytes
        const float2 *lv = reinterpret_cast<const float2 *>(cta_ring + (size_t)slot * (kTcTile * 32)) + 4 * lane;
#pragma unroll 1
        for (int li = 0; li < 4; ++li) {
          const int l = lq * 4 + li;
          if (l >= F.G.n_levels) break;
          scatter_segment<false>(F.G, F.D.t[l], l, F.pts + 3 * base, lv + (size_t)l * kTcTile, ns);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ring_empty[slot]);
    }
    } else {
    // ---------------- scatter role: kV4ScatterWarps warps, work item = (tile, quarter), taken round-robin ----------------
    const int sw = warp - 9;
    const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;      // tiles this CTA walks
    for (int64_t item = sw; item < my_tiles * 4; item += kV4ScatterWarps) {
      const int64_t n = item >> 2;
      const int qd = (int)(item & 3), slot = (int)(n % kRingSlots);
      const int64_t pt = (blockIdx.x + n * gridDim.x) * kTcTile + qd * 32 + lane;
      const bool valid = pt < A.in.n_points;
      float xv[3] = {0.f, 0.f, 0.f};
      if (valid) { xv[0] = __ldg(F.pts + 3 * pt); xv[1] = __ldg(F.pts + 3 * pt + 1); xv[2] = __ldg(F.pts + 3 * pt + 2); }
      mbar_wait(&ring_full[slot], (uint32_t)(n / kRingSlots) & 1);
      // rows of this quarter with a gradient (published by the epilogue warps with the tile); 32 gradient-free samples
      // (empty space, occluded samples) cost one shared-memory read
      const uint32_t rows_nz = (ring_nz[slot][qd][0] | ring_nz[slot][qd][1]) & __ballot_sync(0xffffffffu, valid);
      if (rows_nz != 0u && !(F.debug & 1)) {
        const bool act = (rows_nz >> lane) & 1u;
        // two values per level straight from the ring row (L2 / L1 hits), the next level's in flight during this one's
        // scatter: no local-memory staging (the 107 KB x 2 of shared memory leave the SM almost no L1 for stack traffic)
        const float2 *row = reinterpret_cast<const float2 *>(cta_ring + ((size_t)slot * kTcTile + qd * 32 + lane) * 32);
        float2 gc = act ? row[0] : make_float2(0.f, 0.f);
#pragma unroll 1
        for (int l = 0; l < F.G.n_levels; ++l) {
          const float2 gn = (act && l + 1 < F.G.n_levels) ? row[l + 1] : make_float2(0.f, 0.f);
          scatter_level<false>(F.G, F.D.t[l], l, xv, gc.x, gc.y, lane);
          gc = gn;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ring_empty[slot]);
    }
    }
    __syncthreads();
    return;
  }

  // ---------------- epilogue role ----------------
  setmaxnreg_inc<V4Regs<ACTQ>::epi>();
  const int p = tid & 127, half = tid >> 7;
  uint32_t ph = 0;
  // ACTQ = false (no activation quantiser: every unquantised model) compiles the fake-quant out of the epilogue role
  float q[8];
  const float *qrow = nullptr;
  if (ACTQ && A.in.act_q) {
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = __ldg(A.in.act_q + i);
    if (q[5] != 0.f) qrow = q;
  }
  float g_n2 = 0.f, g_n0[4] = {0.f, 0.f, 0.f, 0.f};
  bool first = true;
  uint32_t n_done = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kTcTile;
    const bool valid = base + p < A.in.n_points;
    // the previous tile's B5 (reader of A0 / A1) was waited for at the end of the previous iteration
    tc_load_inputs<SRC_TILE>(sm, A, &F, tile, base, p, half, valid);
    float d_o[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (valid) {
      if (C == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(dout + (base + p) * 4));
        d_o[0] = v.x; d_o[1] = v.y; d_o[2] = v.z; d_o[3] = v.w;
        if (A.in.keep && A.in.keep[base + p] == 0) d_o[3] = 0.f;                  // run_nerf.py:66
      } else {
#pragma unroll
        for (int c = 0; c < 7; ++c) d_o[c] = __ldg(dout + (base + p) * 7 + c);
        if (A.in.keep && A.in.keep[base + p] == 0) d_o[6] = 0.f;
      }
    }
    epi_arrive(&ready, lane);
    // E1: H1 = relu(D1) -> A1
    epi_wait(&done, ph);
    const uint32_t h1_mask = epi_hidden32<true>(lane_addr + TM_D1, sm + TS::A1, p, half, qrow);   // quant code only in ACTQ builds
    epi_arrive(&ready, lane);
    // E2: [sigma, geo] -> CIN[16..32)
    epi_wait(&done, ph);
    if (half == 0) {
      float v[17];
      tmem_ld16(lane_addr + TM_D2, v);
      tmem_ld_wait();
      v[16] = 0.f;
      st_chunk(sm + TS::CIN, chunk_off(p, 2, 4), v + 1);
      st_chunk(sm + TS::CIN, chunk_off(p, 3, 4), v + 9);
    }
    epi_arrive(&ready, lane);
    // E3: colour hidden 1 -> A1C ; (normals) NH
    epi_wait(&done, ph);
    epi_hidden32(lane_addr + TM_D1, sm + TS::A1C, p, half, nullptr);
    if (NORMALS) {
      const float *bias = reinterpret_cast<const float *>(sm + TS::BIAS) + half * 16;
      float v[16];
      tmem_ld16(lane_addr + TM_DN + half * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + bias[j], 0.f);
      st_chunk(sm + TS::NH, chunk_off(p, half * 2, 4), v);
      st_chunk(sm + TS::NH, chunk_off(p, half * 2 + 1, 4), v + 8);
    }
    epi_arrive(&ready, lane);
    // E4: colour hidden 2 -> A2C ; raw normal ; B0: cotangent tiles
    epi_wait(&done, ph);
    float nraw[3] = {0.f, 0.f, 0.f};
    if (NORMALS && half == 0) {
      const float *bias = reinterpret_cast<const float *>(sm + TS::BIAS) + 32;
      float v[16];
      tmem_ld16(lane_addr + TM_D2, v);
      tmem_ld_wait();
      nraw[0] = v[0] + bias[0]; nraw[1] = v[1] + bias[1]; nraw[2] = v[2] + bias[2];
    }
    epi_hidden32(lane_addr + TM_D1, sm + TS::A2C, p, half, nullptr);
    if (half == 0) {
      float v[8] = {d_o[0], d_o[1], d_o[2], 0.f, 0.f, 0.f, 0.f, 0.f}, z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      st_chunk(sm + TS::DOUT, chunk_off(p, 0, 2), v);
      st_chunk(sm + TS::DOUT, chunk_off(p, 1, 2), z);
      if (NORMALS) {
        const float nn = sqrtf(nraw[0] * nraw[0] + nraw[1] * nraw[1] + nraw[2] * nraw[2]);
        float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (nn > 1e-12f) {
          const float m0 = nraw[0] / nn, m1 = nraw[1] / nn, m2 = nraw[2] / nn;
          const float dot = m0 * d_o[4] + m1 * d_o[5] + m2 * d_o[6];
          r[0] = (d_o[4] - m0 * dot) / nn; r[1] = (d_o[5] - m1 * dot) / nn; r[2] = (d_o[6] - m2 * dot) / nn;
        } else {
          r[0] = d_o[4] / 1e-12f; r[1] = d_o[5] / 1e-12f; r[2] = d_o[6] / 1e-12f;
        }
        st_chunk(sm + TS::DNR, chunk_off(p, 0, 2), r);
        st_chunk(sm + TS::DNR, chunk_off(p, 1, 2), z);
      }
    }
    epi_arrive(&ready, lane);
    // normal head weight gradients on the CUDA cores (611 numbers): every row of NH / DNR must be written first
    if (NORMALS) {
      mlp_sync();
      if (tid < 99) {
        const int j = tid < 96 ? tid >> 5 : tid - 96, k = tid & 31;
        float s = 0.f;
        for (int r = 0; r < kTcTile; ++r)
          s += tile_elem(sm + TS::DNR, r, j, 2) * (tid < 96 ? tile_elem(sm + TS::NH, r, k, 4) : 1.f);
        g_n2 += s;
      }
    }
    // E(B1): dA2pre -> A2C (in place, masked) ; (normals) dNHpre -> NH
    epi_wait(&done, ph);
    if (NORMALS) mlp_sync();                     // all NH reads above are done
    epi_grad32(lane_addr + TM_D1, sm + TS::A2C, p, half, false, 0);
    if (NORMALS) {
      float v[16];
      tmem_ld16(lane_addr + TM_DN + half * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float a[8];
        ld_chunk(sm + TS::NH, chunk_off(p, half * 2 + c, 4), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * c + j] = a[j] > 0.f ? v[8 * c + j] : 0.f;
        st_chunk(sm + TS::NH, chunk_off(p, half * 2 + c, 4), v + 8 * c);
      }
    }
    epi_arrive(&ready, lane);
    if (NORMALS) {
      mlp_sync();                                  // dNHpre rows of every thread are written
      if (tid < 128) {                             // dN0w[j][k] (k < 15) and dN0b[j] (k == 15): 4 outputs per thread
        const int j = tid >> 2, k0 = (tid & 3) * 4;
        for (int r = 0; r < kTcTile; ++r) {
          const float d = tile_elem(sm + TS::NH, r, j, 4);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            g_n0[i] += d * ((k0 + i) < 15 ? tile_elem(sm + TS::CIN, r, 16 + k0 + i, 4) : 1.f);
        }
      }
    }
    // E(B2): dA1pre -> A1C
    epi_wait(&done, ph);
    epi_grad32(lane_addr + TM_D1, sm + TS::A1C, p, half, false, 0);
    epi_arrive(&ready, lane);
    // E(B3): [dsigma, dgeo] -> DH2
    epi_wait(&done, ph);
    if (half == 1) {
      float g[17];
      tmem_ld16(lane_addr + TM_D1 + 16, g + 1);        // dCIN[16..32) = dgeo[0..15) + pad
      tmem_ld_wait();
      if (NORMALS) {
        float v[16];
        tmem_ld16(lane_addr + TM_D2, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 15; ++j) g[1 + j] += v[j];
      }
      g[0] = d_o[3];                                   // dsigma (keep mask applied above when C == 4)
      st_chunk(sm + TS::DH2, chunk_off(p, 0, 2), g);
      st_chunk(sm + TS::DH2, chunk_off(p, 1, 2), g + 8);
    }
    epi_arrive(&ready, lane);
    // E(B4): dH1pre -> A1
    epi_wait(&done, ph);
    epi_grad32(lane_addr + TM_D1, sm + TS::A1, p, half, true, h1_mask);
    epi_arrive(&ready, lane);
    // B5 reads A1 and A0: wait for it before the next tile's inputs overwrite them; its dX goes to the ring slot
    epi_wait(&done, ph);
    {
      const int slot = (int)(n_done % kRingSlots);
      if (n_done >= kRingSlots) mbar_wait(&ring_empty[slot], (uint32_t)(n_done / kRingSlots - 1) & 1);
      float v[16];
      tmem_ld16(lane_addr + TM_D1 + half * 16, v);
      tmem_ld_wait();
      bool nzr = false;
      if (SEG) {
        // slot layout [level][row] float2 (a warp stores 256 contiguous bytes per level)
        float2 *col = reinterpret_cast<float2 *>(cta_ring + (size_t)slot * (kTcTile * 32)) + (size_t)(half * 8) * kTcTile + p;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          col[(size_t)i * kTcTile] = make_float2(v[2 * i], v[2 * i + 1]);
          nzr = nzr || v[2 * i] != 0.f || v[2 * i + 1] != 0.f;
        }
      } else {
        // slot layout [row][level] float2: this thread's half row is 64 contiguous bytes
        float4 *row = reinterpret_cast<float4 *>(cta_ring + ((size_t)slot * kTcTile + p) * 32 + half * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          row[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          nzr = nzr || v[4 * c] != 0.f || v[4 * c + 1] != 0.f || v[4 * c + 2] != 0.f || v[4 * c + 3] != 0.f;
        }
      }
      const uint32_t nzm = __ballot_sync(0xffffffffu, nzr && valid);
      fence_before_sync();                             // the tcgen05.ld above precedes the next tile's first MMA (via `ready`)
      if (lane == 0) {
        ring_nz[slot][warp & 3][half] = nzm;
        mbar_arrive(&ring_full[slot]);                 // release: this warp's ring stores (ordered by the ballot) and the mask
      }
    }
    ++n_done;
    first = false;
  }
  // flush the weight gradients (every MMA has completed: the last commit was waited on)
  if (!first && warp < 4) {
    const bool owner = lane < 16;
    const int row = warp * 16 + lane;
    flush_acc(lane_addr + TM_GC1, 64, owner, row, G.c1, 64, 64, 64, false);
    flush_acc(lane_addr + TM_GS0, 32, owner, row, G.s0, 32, 64, 32, false);
    flush_acc(lane_addr + TM_GC0, 32, owner, row, G.c0, 31, 64, 31, false);
    flush_acc(lane_addr + TM_GS1, 16, owner, row, G.s1, 64, 64, 16, true);
    flush_acc(lane_addr + TM_GC2, 16, owner, row, G.c2, 64, 64, 3, true);
  }
  if (!first && NORMALS) {
    if (tid < 96) { if (G.n2w) atomicAdd(G.n2w + (tid >> 5) * 32 + (tid & 31), g_n2); }
    else if (tid < 99) { if (G.n2b) atomicAdd(G.n2b + (tid - 96), g_n2); }
    if (tid < 128) {
      const int j = tid >> 2, k0 = (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (k0 + i < 15) { if (G.n0w) atomicAdd(G.n0w + j * 15 + k0 + i, g_n0[i]); }
        else if (G.n0b) atomicAdd(G.n0b + j, g_n0[i]);
      }
    }
  }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TM_BWD_COLS);
}

}  // namespace pn

using namespace pn;

extern "C" int pn_debug_timeline(int64_t *buf, int64_t cap_per_thread) {
  PN_REQUIRE(cap_per_thread >= 0 && cap_per_thread < (1 << 20), PN_EINVAL, "cap_per_thread %lld", (long long)cap_per_thread);
  pn::g_tlog_buf = reinterpret_cast<long long *>(buf);
  pn::g_tlog_cap = buf ? (int)cap_per_thread : 0;
  return 0;
}

extern "C" int pn_tc_selftest(const int32_t *cfg, const float *A, const float *B, float *D, pn_stream_t stream) {
  PN_REQUIRE(cfg && A && B && D, PN_EINVAL, "NULL pointer argument");
  SelfTestArgs a;
  a.ra = cfg[0]; a.ca = cfg[1]; a.rb = cfg[2]; a.cb = cfg[3]; a.M = cfg[4]; a.N = cfg[5]; a.ksteps = cfg[6];
  a.a_mn = cfg[7]; a.b_mn = cfg[8];
  a.lbo_a = cfg[9]; a.sbo_a = cfg[10]; a.adv_a = cfg[11]; a.lbo_b = cfg[12]; a.sbo_b = cfg[13]; a.adv_b = cfg[14];
  PN_REQUIRE(a.ca % 8 == 0 && a.cb % 8 == 0 && a.ra % 8 == 0 && a.rb % 8 == 0, PN_ESHAPE, "tile dims must be multiples of 8");
  PN_REQUIRE((a.M == 64 || a.M == 128) && a.N >= 16 && a.N <= 64 && a.N % 16 == 0, PN_ESHAPE, "M in {64,128}, N in 16..64");
  const size_t smem = ((size_t)(a.ra * a.ca * 2 + 127) / 128) * 128 + (size_t)a.rb * a.cb * 2 + 4096;
  PN_REQUIRE(smem <= 200 * 1024, PN_ESHAPE, "tiles too large");
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  PN_REQUIRE(e == cudaSuccess, PN_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  tc_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(a, A, B, D);
  count_launch();
  return check_launch("tc_selftest_kernel");
}

namespace pn {
int check_mlp_args(const pn_mlp_weights *w, const pn_mlp_input *in, bool *normals);   // mlp_fp32.cu
}


namespace pn {
int check_grid_args(const pn_hash_grid *g);                                            // hash_encode.cu

static int launch_tc_fwd(const TcArgs &A, const FieldArgs &F, bool fused, float *out, cudaStream_t st, bool packed = false) {
  // without the normal head the NH tile (last block of the forward map) is not allocated: 55.5 KB -> 4 CTAs/SM
  const int smem = A.normals ? TS::FWD_END : TS::NH;
  // the fused variant carries the hash-gather state: at 4 CTAs/SM (64 registers) it spills and measured 25 %
  // slower than at 3 CTAs/SM (76 registers), so it stays at 3
  static int fused_ctas = -1;
  if (fused_ctas < 0) {                                  // tuning knob
    const char *e = getenv("PN_FWD_CTAS");
    fused_ctas = (e && atoi(e) == 4) ? 4 : ((e && atoi(e) == 2) ? 2 : 3);
    if (fused_ctas == 2)     // two resident CTAs and the rest of the SM's 256 KB as L1 (measurement knob)
      cudaFuncSetAttribute(mlp_tc_fwd_kernel<3, SRC_HASH, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
  }
  const int per_sm = (A.normals || packed) ? 3 : (fused ? fused_ctas : 4);   // x 128 TMEM columns each
  const bool quant = F.qparams != nullptr || A.in.act_q != nullptr;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_tc_fwd_kernel<3, SRC_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS::FWD_END);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tc_fwd_kernel<4, SRC_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS::FWD_END);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tc_fwd_kernel<3, SRC_HASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS::FWD_END);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tc_fwd_kernel<4, SRC_HASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS::FWD_END);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tc_fwd_kernel<3, SRC_HASH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS::FWD_END);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tc_fwd_kernel<4, SRC_HASH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS::FWD_END);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tc_fwd_kernel<3, SRC_PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS::FWD_END);
    PN_REQUIRE(e == cudaSuccess, PN_ECUDA, "cudaFuncSetAttribute(mlp_tc_fwd): %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  const int64_t tiles = ceil_div(A.in.n_points, kTcTile);
  const int64_t cap = (int64_t)sm_count() * per_sm;
  const int blocks = (int)(tiles < cap ? tiles : cap);
  if (packed) mlp_tc_fwd_kernel<3, SRC_PACKED><<<blocks, kTcThreads, smem, st>>>(A, F, out);
  else if (fused && per_sm == 4 && quant) mlp_tc_fwd_kernel<4, SRC_HASH><<<blocks, kTcThreads, smem, st>>>(A, F, out);
  else if (fused && per_sm == 4)      // no normal head (per_sm): its hidden tile, the last region of the map, is never touched,
    // so 54 KB per CTA and four really are resident — measured slower (fine pass 2.34 -> 2.72 ms, coarse 1.14 -> 1.96):
    // 4 x 54 KB of shared memory leave the gather almost no L1
    mlp_tc_fwd_kernel<4, SRC_HASH, false><<<blocks, kTcThreads, TS::NH, st>>>(A, F, out);
  else if (fused && quant) mlp_tc_fwd_kernel<3, SRC_HASH><<<blocks, kTcThreads, smem, st>>>(A, F, out);
  else if (fused) mlp_tc_fwd_kernel<3, SRC_HASH, false><<<blocks, kTcThreads, smem, st>>>(A, F, out);
  else if (A.normals) mlp_tc_fwd_kernel<3, SRC_F32><<<blocks, kTcThreads, smem, st>>>(A, F, out);
  else mlp_tc_fwd_kernel<4, SRC_F32><<<blocks, kTcThreads, smem, st>>>(A, F, out);
  count_launch();
  return check_launch("mlp_tc_fwd_kernel");
}

// Which fused backward runs (PN_FIELD_BWD): the three-role kernel with the ring-decoupled scatter (default) or "v1", the
// single-role kernel it replaced — kept as the instrumented baseline (clock64 timeline) the design was derived from.
static int bwd_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("PN_FIELD_BWD");
    v = (e && e[0] == 'v' && e[1] == '1') ? 0 : 3;
  }
  return v;
}
// bytes of the dX ring of the default backward: one ring of kRingSlots 128 x 32 fp32 tiles per resident CTA
static int64_t field_bwd_workspace_bytes() { return (int64_t)sm_count() * 2 * kRingSlots * kTcTile * 32 * 4; }

static int launch_tc_bwd(const TcArgs &A, const FieldArgs &F, bool fused, const float *dout, float *dfeat,
                         int64_t dfeat_stride, float *dsh, int64_t dsh_stride, const pn_mlp_grads &dw, cudaStream_t st,
                         void *workspace = nullptr, int64_t workspace_bytes = 0) {
  const int smem = TS::BWD_END;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_tc_bwd_kernel<SRC_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_tc_bwd_kernel<SRC_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
#define PN_BWD4_ATTR(N, S, Q) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(field_bwd4_kernel<N, S, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
    PN_BWD4_ATTR(false, false, false); PN_BWD4_ATTR(false, true, false); PN_BWD4_ATTR(true, false, false); PN_BWD4_ATTR(true, true, false);
    PN_BWD4_ATTR(false, false, true); PN_BWD4_ATTR(false, true, true); PN_BWD4_ATTR(true, false, true); PN_BWD4_ATTR(true, true, true);
#undef PN_BWD4_ATTR
    PN_REQUIRE(e == cudaSuccess, PN_ECUDA, "cudaFuncSetAttribute(mlp_tc_bwd): %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  const int64_t tiles = ceil_div(A.in.n_points, kTcTile);
  const int64_t cap = (int64_t)sm_count() * 2;          // 2 CTAs/SM: 2 x 256 TMEM columns, 2 x 107 KB smem
  const int blocks = (int)(tiles < cap ? tiles : cap);
  if (fused && bwd_variant() == 3) PN_REQUIRE(A.C == 7 || ((uintptr_t)dout & 15) == 0, PN_EINVAL, "dout must be 16-byte aligned");
  if (fused && bwd_variant() == 3) {
    PN_REQUIRE(workspace != nullptr && ((uintptr_t)workspace & 15) == 0 && workspace_bytes >= field_bwd_workspace_bytes(), PN_EINVAL,
               "pn_field_bwd_bf16 needs a 16-byte aligned workspace of pn_field_bwd_workspace_bytes() = %lld bytes (got %lld)",
               (long long)field_bwd_workspace_bytes(), (long long)workspace_bytes);
    static int seg_min = -1;                     // PN_SCATTER_SEG_MIN: samples per ray from which the serial segment scatter runs
    if (seg_min < 0) {
      const char *e = getenv("PN_SCATTER_SEG_MIN");
      seg_min = e ? atoi(e) : 128;
    }
    const bool seg = A.in.samples_per_ray >= seg_min;
    float *ring = reinterpret_cast<float *>(workspace);
    const int variant = (A.normals ? 4 : 0) | (seg ? 2 : 0) | (A.in.act_q != nullptr ? 1 : 0);
#define PN_BWD4_LAUNCH(N, S, Q) field_bwd4_kernel<N, S, Q><<<blocks, kV4Threads, smem, st>>>(A, F, dout, dw, ring)
    switch (variant) {
      case 0: PN_BWD4_LAUNCH(false, false, false); break;
      case 1: PN_BWD4_LAUNCH(false, false, true); break;
      case 2: PN_BWD4_LAUNCH(false, true, false); break;
      case 3: PN_BWD4_LAUNCH(false, true, true); break;
      case 4: PN_BWD4_LAUNCH(true, false, false); break;
      case 5: PN_BWD4_LAUNCH(true, false, true); break;
      case 6: PN_BWD4_LAUNCH(true, true, false); break;
      default: PN_BWD4_LAUNCH(true, true, true); break;
    }
#undef PN_BWD4_LAUNCH
  } else if (fused)
    mlp_tc_bwd_kernel<SRC_TILE><<<blocks, kTcThreads, smem, st>>>(A, F, dout, dfeat, dfeat_stride, dsh, dsh_stride, dw);
  else
    mlp_tc_bwd_kernel<SRC_F32><<<blocks, kTcThreads, smem, st>>>(A, F, dout, dfeat, dfeat_stride, dsh, dsh_stride, dw);
  count_launch();
  return check_launch("mlp_tc_bwd_kernel");
}
}  // namespace pn

extern "C" int pn_mlp_fwd_bf16(const pn_mlp_weights *w, const pn_mlp_input *in, float *out, pn_stream_t stream) {
  bool normals = false;
  if (int e = check_mlp_args(w, in, &normals)) return e;
  PN_REQUIRE(out != nullptr, PN_EINVAL, "out is NULL");
  if (in->n_points == 0) return 0;
  TcArgs A;
  A.w = *w; A.in = *in; A.normals = normals; A.C = normals ? 7 : 4;
  PN_REQUIRE(A.C == 7 || ((uintptr_t)out & 15) == 0, PN_EINVAL, "out must be 16-byte aligned");
  FieldArgs F = {};
  return launch_tc_fwd(A, F, false, out, as_stream(stream));
}

extern "C" int pn_mlp_bwd_bf16(const pn_mlp_weights *w, const pn_mlp_input *in, const float *dout, float *dfeat,
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
  TcArgs A;
  A.w = *w; A.in = *in; A.normals = normals; A.C = normals ? 7 : 4;
  FieldArgs F = {};
  return launch_tc_bwd(A, F, false, dout, dfeat, dfeat_stride, dsh, dsh_stride, *dw, as_stream(stream));
}

static int fill_field(FieldArgs &F, const pn_hash_grid *grid, const float *const *tables, float *const *dtables,
                      const float *pts, const float *qparams) {
  if (int e = check_grid_args(grid)) return e;
  PN_REQUIRE(grid->n_levels == 16, PN_ESHAPE, "the fused field kernels are built for 16 levels (32 features), got %d",
             grid->n_levels);
  PN_REQUIRE(pts != nullptr, PN_EINVAL, "pts is NULL");
  F.G = make_grid_dev(*grid);
  for (int l = 0; l < PN_MAX_LEVELS; ++l) {
    F.T.t[l] = tables ? reinterpret_cast<const float2 *>(tables[l]) : nullptr;
    F.D.t[l] = dtables ? reinterpret_cast<float2 *>(dtables[l]) : nullptr;
    PN_REQUIRE(!tables || (tables[l] && ((uintptr_t)tables[l] & 15) == 0), PN_EINVAL,
               "tables[%d] is NULL or not 16-byte aligned", l);
    PN_REQUIRE(!dtables || (dtables[l] && ((uintptr_t)dtables[l] & 15) == 0), PN_EINVAL,
               "dtables[%d] is NULL or not 16-byte aligned", l);
  }
  F.pts = pts;
  F.qparams = qparams;
  static int dbg = -1;
  if (dbg < 0) {
    const char *e = getenv("PN_DEBUG_FLAGS");
    dbg = e ? atoi(e) : 0;
  }
  F.debug = dbg;
  static int sc_direct = -1, sc_maxlen = -1;
  if (sc_direct < 0) {
    const char *a = getenv("PN_SCATTER_DIRECT"), *b = getenv("PN_SCATTER_MAXLEN");
    sc_direct = a ? atoi(a) : 20;
    sc_maxlen = b ? atoi(b) : 32;
    if (sc_maxlen != 1 && sc_maxlen != 2 && sc_maxlen != 4 && sc_maxlen != 8 && sc_maxlen != 16) sc_maxlen = 32;
  }
  F.sc_direct = sc_direct;
  F.sc_maxlen = sc_maxlen;
  F.tlog = g_tlog_buf;
  F.tlog_cap = g_tlog_cap;
  return 0;
}

extern "C" int pn_field_fwd_bf16(const pn_hash_grid *grid, const float *const *tables, const float *qparams,
                                 const pn_mlp_weights *w, const float *pts, const float *dirs, int samples_per_ray,
                                 const float *act_q, int64_t n_points, float *out, uint8_t *keep, void *feat_tiles,
                                 pn_stream_t stream) {
  PN_REQUIRE(w && tables && dirs && out, PN_EINVAL, "NULL pointer argument");
  pn_mlp_input in = {};
  in.feat = pts;            // placeholder so that the shared argument check passes; the fused kernel never reads it
  in.feat_stride = 32;
  in.dirs = dirs; in.samples_per_ray = samples_per_ray; in.act_q = act_q; in.n_points = n_points;
  bool normals = false;
  PN_REQUIRE(((uintptr_t)pts & 15) == 0, PN_EINVAL, "pts must be 16-byte aligned");
  if (int e = check_mlp_args(w, &in, &normals)) return e;
  if (n_points == 0) return 0;
  TcArgs A;
  A.w = *w; A.in = in; A.normals = normals; A.C = normals ? 7 : 4;
  PN_REQUIRE(A.C == 7 || ((uintptr_t)out & 15) == 0, PN_EINVAL, "out must be 16-byte aligned");
  FieldArgs F = {};
  if (int e = fill_field(F, grid, tables, nullptr, pts, qparams)) return e;
  F.keep_out = keep;
  F.featb = reinterpret_cast<uint4 *>(feat_tiles);
  PN_REQUIRE(((uintptr_t)feat_tiles & 15) == 0, PN_EINVAL, "feat_tiles must be 16-byte aligned");
  return launch_tc_fwd(A, F, true, out, as_stream(stream));
}

extern "C" int pn_field_bwd_bf16(const pn_hash_grid *grid, float *const *dtables, const pn_mlp_weights *w,
                                 const void *feat_tiles, const float *pts, const float *dirs, int samples_per_ray,
                                 const float *act_q, const uint8_t *keep, const float *dout, int64_t n_points,
                                 const pn_mlp_grads *dw, void *workspace, int64_t workspace_bytes, pn_stream_t stream) {
  PN_REQUIRE(w && dtables && dirs && dout && dw && feat_tiles, PN_EINVAL, "NULL pointer argument");
  pn_mlp_input in = {};
  in.feat = pts;
  in.feat_stride = 32;
  in.dirs = dirs; in.samples_per_ray = samples_per_ray; in.act_q = act_q; in.keep = keep; in.n_points = n_points;
  bool normals = false;
  PN_REQUIRE(((uintptr_t)pts & 15) == 0 && ((uintptr_t)feat_tiles & 15) == 0, PN_EINVAL, "pts / feat_tiles alignment");
  if (int e = check_mlp_args(w, &in, &normals)) return e;
  if (n_points == 0) return 0;
  TcArgs A;
  A.w = *w; A.in = in; A.normals = normals; A.C = normals ? 7 : 4;
  FieldArgs F = {};
  if (int e = fill_field(F, grid, nullptr, dtables, pts, nullptr)) return e;
  F.featb = reinterpret_cast<uint4 *>(const_cast<void *>(feat_tiles));
  static int split = -1;
  if (split < 0) {                              // tuning knob (default: measured best on B200)
    const char *e = getenv("PN_SCATTER_SPLIT");
    split = e ? atoi(e) : 8;
    if (split < 1 || split > 15) split = 8;
  }
  F.scatter_split = split;
  return launch_tc_bwd(A, F, true, dout, nullptr, 32, nullptr, 16, *dw, as_stream(stream), workspace, workspace_bytes);
}

extern "C" int64_t pn_field_bwd_workspace_bytes(void) { return pn::field_bwd_workspace_bytes(); }

namespace pn {
int fill_packed(PackedDev &T, const pn_hash_grid *grid, const pn_packed_tables *packed);   // hash_encode.cu
}

extern "C" int pn_field_fwd_bf16_packed(const pn_hash_grid *grid, const pn_packed_tables *packed, const pn_mlp_weights *w,
                                        const float *pts, const float *dirs, int samples_per_ray, const float *act_q,
                                        int64_t n_points, float *out, uint8_t *keep, pn_stream_t stream) {
  PN_REQUIRE(w && packed && dirs && out, PN_EINVAL, "NULL pointer argument");
  pn_mlp_input in = {};
  in.feat = pts;            // placeholder for the shared argument check; never read
  in.feat_stride = 32;
  in.dirs = dirs; in.samples_per_ray = samples_per_ray; in.act_q = act_q; in.n_points = n_points;
  bool normals = false;
  PN_REQUIRE(((uintptr_t)pts & 15) == 0, PN_EINVAL, "pts must be 16-byte aligned");
  if (int e = check_mlp_args(w, &in, &normals)) return e;
  TcArgs A;
  A.w = *w; A.in = in; A.normals = normals; A.C = normals ? 7 : 4;
  PN_REQUIRE(A.C == 7 || ((uintptr_t)out & 15) == 0, PN_EINVAL, "out must be 16-byte aligned");
  FieldArgs F = {};
  if (int e = fill_field(F, grid, nullptr, nullptr, pts, nullptr)) return e;
  if (int e = fill_packed(F.PK, grid, packed)) return e;
  if (n_points == 0) return 0;
  F.keep_out = keep;
  return launch_tc_fwd(A, F, true, out, as_stream(stream), true);
}
