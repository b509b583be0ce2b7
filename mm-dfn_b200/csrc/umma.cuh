// Blackwell (sm_100a) tensor-core primitives used by the tcgen05 kernels of this library:
// TMEM allocation, shared-memory matrix descriptors, tcgen05.mma kind::tf32, tcgen05.commit /
// mbarrier hand-off, tcgen05.ld.  Descriptor bit layouts follow the SM100 UMMA definitions
// (SmemDescriptor / InstrDescriptor); everything here is inline PTX.
#pragma once
#include "common.cuh"

namespace mmdfn {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------
// one full warp; ncols power of two in [32, 512]; writes the base address to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane (warp w of the CTA owns lanes 32*(w%4)..+31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// two 16-column loads (main accumulator + correction accumulator) in flight together, ONE wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr_a, uint32_t taddr_b, float (&v)[16], float (&w)[16]) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr_a)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(taddr_b)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) {
    v[i] = __uint_as_float(r[i]);
    w[i] = __uint_as_float(q[i]);
  }
}

// "stage full" hand-off of a converter warp: every lane makes its generic-proxy writes visible to the async proxy, the
// warp converges, ONE lane arrives (barrier initialised with the number of converter WARPS) -- 8 arrivals per stage
// instead of 256 serialized shared-memory atomics
__device__ __forceinline__ void warp_arrive_full(uint64_t* bar) {
  fence_proxy_async_smem();
  __syncwarp();
  if ((threadIdx.x & 31) == 0)
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE, K-major canonical layout: an operand tile is a grid of
// 8-row x 16-byte "core matrices" (128 contiguous bytes each); `lbo` = byte distance between the two
// core matrices one MMA consumes along K (K = 8 tf32 = 32 bytes), `sbo` = byte distance between
// consecutive 8-row groups along M/N.  Fields are in 16-byte units; bits 46-47 = descriptor version 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Instruction descriptor for kind::tf32, fp32 accumulate; a_mn / b_mn = 1 selects the MN-major operand layout
// (bits 15 / 16), 0 the K-major one.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand is read from tensor memory (lane = row m, 32-bit column = k)
__device__ __forceinline__ void mma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Warp-convergent forms: the WHOLE warp executes the call with warp-uniform operands and one elected lane issues the
// instruction.  ptxas then keeps the operands in uniform registers; issuing from inside an `if (lane == 0)` region costs
// an R2UR / ELECT / BROADCAST sequence per instruction instead (~60 extra cycles per MMA on the issuing thread).
__device__ __forceinline__ void mma_tf32_ta_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One k-step of the 3-term tf32 split (main += a_hi b_hi; corr += a_lo b_hi; corr += a_hi b_lo) from shared-memory
// operands, issued by one elected lane of a converged warp.  dhi = high word shared by all four descriptors, the other
// four arguments are their low words (start-address field: base + byte offset / 16).
__device__ __forceinline__ void kstep3_elect(uint32_t tmem_main, uint32_t tmem_corr, uint32_t dhi, uint32_t a_hi, uint32_t a_lo,
                                             uint32_t b_hi, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e, t;\n\t.reg .b64 ah, al, bh, bl;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %8, 0;\n\t"
      "setp.eq.b32 t, %8, %8;\n\t"
      "mov.b64 ah, {%3, %2};\n\t"
      "mov.b64 al, {%4, %2};\n\t"
      "mov.b64 bh, {%5, %2};\n\t"
      "mov.b64 bl, {%6, %2};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %7, p;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], al, bh, %7, p;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], ah, bl, %7, t;\n\t}"
      ::"r"(tmem_main), "r"(tmem_corr), "r"(dhi), "r"(a_hi), "r"(a_lo), "r"(b_hi), "r"(b_lo), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)) : "memory");
}

// arrive on `bar` once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- 3xTF32 operand split ---------------------------------------------------------------------------
// x = hi + lo with hi = rna(x) (tf32, low 13 mantissa bits zero) and lo = x - hi exact in fp32; the tensor core
// drops the low 13 bits of lo, leaving a relative error ~2^-21: a_hi*b_hi + a_lo*b_hi + a_hi*b_lo ~ fp32 product.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  // rna on the bit pattern: + half an ulp of the 13 dropped bits, clear them (= cvt.rna.tf32.f32 for finite x, in two
  // integer operations instead of the five the conversion expands to with its NaN / infinity handling)
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

}  // namespace umma
}  // namespace mmdfn
