// tcgen05 / TMEM / mbarrier primitives for sm_100a, written as inline PTX (no CUTLASS dependency).
//
// Shared-memory operand tiles use ONE layout everywhere ("tile layout", no swizzle): a [R x C] bf16 tile
// is a grid of 8x8 core matrices of 128 contiguous bytes; element (r, c) lives at
//     (r/8) * RG + (c/8) * 128 + (r%8) * 16 + (c%8) * 2      with RG = (C/8) * 128 bytes.
// Read with a K-major descriptor (LBO = 128, SBO = RG) it is the operand X[r][c] with c = K;
// read with an MN-major descriptor (LBO = RG, SBO = 128) the SAME bytes are the operand with r = K and
// c = M/N.  That is what lets one copy of an activation tile serve as the A operand of the next layer
// (K-major), as the B operand of the input-gradient GEMM's weights (MN-major) and as both operands of
// the weight-gradient GEMM (MN-major, reduction over the 128 points of the tile).
#pragma once
#include <cuda_bf16.h>

#include "pn_common.cuh"

namespace pn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors -----------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (SWIZZLE_NONE): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type=0 [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// 32-bit instruction descriptor, kind::f16, A/B = bf16, D = f32.
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- tile layout -------------------------------------------------------------------------------------
// byte offset of the 16-byte chunk (row r, columns 8*c8 .. 8*c8+7) of a tile with C8 chunks per row
__host__ __device__ __forceinline__ constexpr uint32_t chunk_off(int r, int c8, int C8) {
  return (uint32_t)((r >> 3) * (C8 * 128) + c8 * 128 + (r & 7) * 16);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

__device__ __forceinline__ void st_chunk(uint8_t *tile, uint32_t off, const float *v) {
  uint4 u = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  *reinterpret_cast<uint4 *>(tile + off) = u;
}

__device__ __forceinline__ void ld_chunk(const uint8_t *tile, uint32_t off, float *v) {
  const uint4 u = *reinterpret_cast<const uint4 *>(tile + off);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

// ---- TMEM -----------------------------------------------------------------------------------------------
// one full warp allocates `cols` (power of two >= 32) columns and publishes the base address in *slot
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (the tensor core's operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread = lane of its warp's quarter)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 2 consecutive columns (one hash level's two feature gradients)
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float &a, float &b) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
  a = __uint_as_float(r0);
  b = __uint_as_float(r1);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- MMA ---------------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T, one K=16 step; issued by ONE thread.
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// One probe: the thread is suspended in hardware until the phase completes or a time limit passes.  The limit is given
// explicitly (nanoseconds): with the default, the ~9 waits per tile of 17 warps re-probed so often that mbarrier traffic
// was 70 % of the kernel's shared-memory wavefronts and the LSU data pipe sat at 82 % (ncu, round 2) — in a kernel
// whose real work on that pipe is reductions and gathers.  Completion still wakes the thread at once.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity), "r"(20000u)
      : "memory");
  return ok;
}
// Bounded spin: a protocol bug must surface as a launch failure (trap), never as a GPU that hangs until the
// watchdog of whoever launched us fires.  try_wait suspends in hardware for up to its time limit per probe,
// so 2^22 failed probes is at least tens of milliseconds (seconds with the 20 us limit below) — orders of magnitude beyond any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t spins = 0;
  while (!mbar_try_wait(a, parity)) {
    if (++spins == (1u << 20)) __trap();
  }
}
// one arrival of the executing thread (release semantics at CTA scope)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (all 32 lanes must execute it)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// warp index as a warp-uniform value (the compiler keeps what derives from it on the uniform datapath)
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
// bring `bytes` (multiple of 16) starting at a 16-byte aligned global address into L2, asynchronously
__device__ __forceinline__ void prefetch_l2(const void *gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
// named barrier over a subset of the CTA's warps (id 1..15; id 0 is __syncthreads)
template <int ID, int THREADS>
__device__ __forceinline__ void named_bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(THREADS) : "memory");
}
// register re-allocation between the roles of a warp-specialised CTA (all 4 warps of a warpgroup execute it)
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

}  // namespace tc
}  // namespace pn
