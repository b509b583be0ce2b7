// Layout probe (debug aid, not on any product path): one tcgen05.mma with A = identity on the first 8 rows/k
// (K-major, known-good) and the B operand region filled with its own word offsets.  D[k][n] then reads back the
// shared-memory word the tensor core uses as B(k, n) for a given descriptor (lbo, sbo) and major-ness.
#include "umma.cuh"

namespace mmdfn {
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(float* out, int N, int lbo, int sbo, int b_mn, int probe_a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* ident = reinterpret_cast<float*>(smem);                 // 16 KB: identity operand, K-major (SBO 528, LBO 128)
  float* probe = reinterpret_cast<float*>(smem + 16384);         // 48 KB: word-offset pattern
  for (int i = tid; i < 4096; i += 128) ident[i] = 0.f;
  for (int i = tid; i < 12288; i += 128) probe[i] = (float)i;
  __syncthreads();
  if (tid < 8) ident[(tid * 16 + (tid >> 2) * 128 + (tid & 3) * 4) / 4] = 1.0f;   // element (r=tid, k=tid) of row group 0
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_barrier_init(); }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint64_t d_ident = umma::smem_desc(umma::smem_u32(ident), 128, 528);
    const uint64_t d_probe = umma::smem_desc(umma::smem_u32(probe), (uint32_t)lbo, (uint32_t)sbo);
    if (!probe_a) {
      // D[m][n] = sum_k I[m][k] * P(k, n)  ->  rows 0..7 show P(k = m, n)
      umma::mma_tf32(tmem, d_ident, d_probe, umma::idesc_tf32(128, N, 0, b_mn), 0u);
    } else {
      // D[m][n] = sum_k P(m, k) * I[n][k]  ->  columns 0..7 show P(m, k = n)
      umma::mma_tf32(tmem, d_probe, d_ident, umma::idesc_tf32(128, N, b_mn, 0), 0u);
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after_sync();
  const int row = warp * 32 + lane;
  for (int cb = 0; cb < N; cb += 16) {
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cb, v);
    for (int q = 0; q < 16; q++) out[row * N + cb + q] = v[q];
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// Second probe: the A operand read from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc).  Thread m writes
// A(m, k) = 16 m + k, k = 0..15, into 16 consecutive TMEM columns of its lane with tcgen05.st.32x32b; B = 16 x 16
// identity (K-major, known-good layout); two K = 8 MMAs, the second with the A address advanced by 8 columns.
// out[m][n] must read back 16 m + n.
__global__ void __launch_bounds__(128, 1) umma_probe_ta_kernel(float* out, int a_col) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* ident = reinterpret_cast<float*>(smem);                 // identity operand, K-major (SBO 528, LBO 128)
  for (int i = tid; i < 4096; i += 128) ident[i] = 0.f;
  __syncthreads();
  if (tid < 16) ident[((tid >> 3) * 528 + (tid >> 2) * 128 + (tid & 7) * 16 + (tid & 3) * 4) / 4] = 1.0f;   // (n = tid, k = tid)
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_barrier_init(); }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  {
    uint32_t r[16];
    for (int k = 0; k < 16; k++) r[k] = __float_as_uint((float)(16 * tid + k));
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)a_col;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  if (tid == 0) {
    const uint32_t base = umma::smem_u32(ident);
    const uint32_t idesc = umma::idesc_tf32(128, 16);
    for (int j = 0; j < 2; j++) {
      const uint64_t bdesc = umma::smem_desc(base + j * 256, 128, 528);
      umma::mma_tf32_ta(tmem, tmem + (uint32_t)a_col + 8 * j, bdesc, idesc, j > 0 ? 1u : 0u);
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after_sync();
  {
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int q = 0; q < 16; q++) out[tid * 16 + q] = v[q];
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// Third probe: issue rate of back-to-back tcgen05.mma instructions (operands are zeros; only the timing matters).
// kind 0 = tf32 (K = 8), 1 = bf16 (kind::f16, K = 16); A from shared memory or tensor memory; `issuers` threads (lane 0
// of warps 0..issuers-1) issue concurrently, each into its own accumulator (columns 64 w, N <= 64 when issuers > 1).
// out[0] = cycles until every issuer's `reps` MMAs completed, out[1] = the same for 2 * reps.
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(long long* out, int kind, int N, int a_tmem, int reps, int issuers) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { for (int i = 0; i < 4; i++) umma::mbar_init(&bar[i], 1); umma::fence_barrier_init(); }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256;
    for (int c = 0; c < 128; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + c), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  uint32_t parity = 0;
  for (int pass = 0; pass < 2; pass++) {
    const int n = reps << pass;
    __syncthreads();
    const long long t0 = clock64();
    if (lane == 0 && warp < issuers) {
      const uint32_t base = umma::smem_u32(smem);
      const uint32_t idesc = (1u << 4) | ((kind == 0 ? 2u : 1u) << 7) | ((kind == 0 ? 2u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
      const uint64_t adesc = umma::smem_desc(base, 128, 528);
      const uint64_t bdesc = umma::smem_desc(base + 16384, 128, 528);
      const uint32_t tmem_d = tmem + (uint32_t)(issuers > 1 ? 64 * warp : 0);
      const uint32_t tmem_a = tmem + 256 + 32 * warp;
      if (a_tmem) {
        if (kind == 0) {
          for (int i = 0; i < n; i++) umma::mma_tf32_ta(tmem_d, tmem_a, bdesc, idesc, i > 0 ? 1u : 0u);
        } else {
          for (int i = 0; i < n; i++)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(i > 0 ? 1u : 0u) : "memory");
        }
      } else {
        if (kind == 0) {
          for (int i = 0; i < n; i++) umma::mma_tf32(tmem_d, adesc, bdesc, idesc, i > 0 ? 1u : 0u);
        } else {
          for (int i = 0; i < n; i++)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i > 0 ? 1u : 0u) : "memory");
        }
      }
      umma::mma_commit(&bar[warp]);
      umma::mbar_wait(&bar[warp], parity);
    }
    parity ^= 1;
    __syncthreads();
    if (tid == 0) out[pass] = clock64() - t0;
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}
}  // namespace mmdfn

extern "C" int mmdfn_umma_rate(long long* out, int kind, int N, int a_tmem, int reps, int issuers, void* stream) {
  using namespace mmdfn;
  if (!out) return MMDFN_ENULL;
  if (kind < 0 || kind > 1 || N < 16 || N > 256 || (N & 15) || reps < 1 || issuers < 1 || issuers > 4 || (issuers > 1 && N > 64)) return MMDFN_EINVAL;
  static bool configured = false;
  if (!configured) {
    MMDFN_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    configured = true;
  }
  umma_rate_kernel<<<1, 128, 65536, (cudaStream_t)stream>>>(out, kind, N, a_tmem, reps, issuers);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_umma_probe_ta(float* out, int a_col, void* stream) {
  using namespace mmdfn;
  if (!out) return MMDFN_ENULL;
  if (a_col < 16 || a_col > 240) return MMDFN_EINVAL;
  umma_probe_ta_kernel<<<1, 128, 16384, (cudaStream_t)stream>>>(out, a_col);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_umma_probe(float* out, int N, int lbo, int sbo, int mn_major, int probe_a, void* stream) {
  using namespace mmdfn;
  static bool configured = false;
  if (!configured) {
    MMDFN_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    configured = true;
  }
  umma_probe_kernel<<<1, 128, 65536, (cudaStream_t)stream>>>(out, N, lbo, sbo, mn_major, probe_a);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
