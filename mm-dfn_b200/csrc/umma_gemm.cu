// Dense GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM) with the
// 3-term TF32 split, so that results stay at fp32 accuracy (the 1e-4 logit budget and the acos() in the
// adjacency do not tolerate single-pass TF32).
//
//   C[M,N] = act(alpha * op(A) op(B) + beta*C + bias)        same contract as mmdfn_gemm
//
// One CTA (128 threads, two CTAs per SM) owns a 128 x 112 output tile.  Per 16-wide K chunk the CTA reads the
// fp32 operand rows from global memory (128-bit loads when aligned, two chunks in flight in registers),
// splits every value into tf32 hi/lo parts and writes them to shared memory directly in the UMMA
// SWIZZLE_NONE K-major core-matrix layout (this is also where transposed operands are transposed, so
// NT / NN / TN share one MMA configuration).  One thread then issues 3 MMAs per 8-wide k-step: hi*hi into
// the main TMEM accumulator, lo*hi and hi*lo into a second one (the tensor core's fp32 accumulate truncates,
// so keeping the 2^-11-scaled corrections apart leaves the main sum with a third of the update count; measured
// error equals the FFMA kernel's).  tcgen05.commit releases the shared-memory stage through an mbarrier; three
// stages overlap load + split with the tensor-core work.  Epilogue: tcgen05.ld (thread = one output row).
#include "umma.cuh"
#include "internal.cuh"

namespace mmdfn {

constexpr int UG_THREADS = 128;
constexpr int UG_BN = 112;         // output columns per CTA (N = 100 / 200 / 300 / 400 / 600 -> 1 / 2 / 3 / 4 / 6 tiles)
constexpr int UG_KC = 16;          // K elements per stage (2 k-steps of 8)
constexpr int UG_STAGES = 3;
constexpr int UG_LBO = 128;        // bytes between the two core matrices of one k-step
constexpr int UG_SBO = 528;        // bytes between 8-row groups: 4 core matrices (512 B) + 16 B pad (bank spread)
constexpr int UG_TMEM_COLS = 256;  // [0,112): hi*hi accumulator, [128,240): correction accumulator
constexpr int UG_CORR_COL = 128;
constexpr int UG_A_PART = 16 * UG_SBO;
constexpr int UG_B_PART = (UG_BN / 8) * UG_SBO;
constexpr int UG_STAGE_BYTES = 2 * (UG_A_PART + UG_B_PART);
constexpr int UG_SMEM = UG_STAGES * UG_STAGE_BYTES;      // 95 KB -> two CTAs per SM

struct UGemmArgs {
  const float* A; i64 lda;
  const float* B; i64 ldb;
  float* C; i64 ldc;
  const float* bias;
  int M, N, K;
  float alpha, beta;
  int act, splits;
};

struct OperandRegs {
  float4 v[4];
};

// ---- source rows are K-contiguous: element (r, k) at g[r*ld + k]; R rows (multiple of 8, <= 128) ------------
template <int R>
__device__ __forceinline__ void load_kmajor(OperandRegs& o, const float* __restrict__ g, i64 ld, int row0, int row_end,
                                            int k0, int k_end, bool vec) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r_in = lane & 7, c = lane >> 3;
  const int k = k0 + 4 * c;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int rg = warp + 4 * i;
    const int row = row0 + rg * 8 + r_in;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rg < R / 8 && row < row_end && k < k_end) {
      const float* p = g + (i64)row * ld + k;
      if (vec && k + 3 < k_end) {
        v = *reinterpret_cast<const float4*>(p);
      } else {
        v.x = p[0];
        if (k + 1 < k_end) v.y = p[1];
        if (k + 2 < k_end) v.z = p[2];
        if (k + 3 < k_end) v.w = p[3];
      }
    }
    o.v[i] = v;
  }
}

template <int R>
__device__ __forceinline__ void store_kmajor(const OperandRegs& o, uint8_t* hi, uint8_t* lo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r_in = lane & 7, c = lane >> 3;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int rg = warp + 4 * i;
    if (rg >= R / 8) continue;
    const int off = rg * UG_SBO + c * UG_LBO + r_in * 16;
    float4 h, l;
    umma::split_tf32(o.v[i].x, h.x, l.x);
    umma::split_tf32(o.v[i].y, h.y, l.y);
    umma::split_tf32(o.v[i].z, h.z, l.z);
    umma::split_tf32(o.v[i].w, h.w, l.w);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

// ---- source is MN-contiguous: element (r, k) at g[k*ld + r]  (transposed on the way into shared memory) ----
template <int R>
__device__ __forceinline__ void load_mnmajor(OperandRegs& o, const float* __restrict__ g, i64 ld, int row0, int row_end,
                                             int k0, int k_end, bool vec) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = 4 * lane;
  const int row = row0 + r;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int k = k0 + warp + 4 * i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R && k < k_end && row < row_end) {
      const float* q = g + (i64)k * ld + row;
      if (vec && row + 3 < row_end) {
        v = *reinterpret_cast<const float4*>(q);
      } else {
        v.x = q[0];
        if (row + 1 < row_end) v.y = q[1];
        if (row + 2 < row_end) v.z = q[2];
        if (row + 3 < row_end) v.w = q[3];
      }
    }
    o.v[i] = v;
  }
}

template <int R>
__device__ __forceinline__ void store_mnmajor(const OperandRegs& o, uint8_t* hi, uint8_t* lo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = 4 * lane;
  if (r >= R) return;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int kl = warp + 4 * i;                                   // k within the chunk
    const int koff = (kl >> 2) * UG_LBO + (kl & 3) * 4;
    const float e[4] = {o.v[i].x, o.v[i].y, o.v[i].z, o.v[i].w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int rr = r + q;
      const int off = (rr >> 3) * UG_SBO + (rr & 7) * 16 + koff;
      float h, l;
      umma::split_tf32(e[q], h, l);
      *reinterpret_cast<float*>(hi + off) = h;
      *reinterpret_cast<float*>(lo + off) = l;
    }
  }
}

// MODE 0: NT (A[M,K], B[N,K])   1: NN (A[M,K], B[K,N])   2: TN (A[K,M], B[K,N])
template <int MODE>
__global__ void __launch_bounds__(UG_THREADS, 2) umma_gemm_kernel(UGemmArgs p) {
  constexpr int BN = UG_BN;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_free[UG_STAGES];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN;

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, UG_TMEM_COLS);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < UG_STAGES; s++) umma::mbar_init(&bar_free[s], 1);
    umma::fence_barrier_init();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  int kb = 0, ke = p.K;
  if (p.splits > 1) {
    const int chunk = ((p.K + p.splits - 1) / p.splits + UG_KC - 1) / UG_KC * UG_KC;
    kb = blockIdx.z * chunk;
    ke = min(p.K, kb + chunk);
  }
  const int nchunks = ke > kb ? (ke - kb + UG_KC - 1) / UG_KC : 0;

  const bool a_vec = ((p.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
  const bool b_vec = ((p.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);
  auto load_chunk = [&](int c, OperandRegs& ra, OperandRegs& rb) {
    const int k0 = kb + c * UG_KC;
    if (MODE == 2) load_mnmajor<128>(ra, p.A, p.lda, m0, p.M, k0, ke, a_vec);
    else load_kmajor<128>(ra, p.A, p.lda, m0, p.M, k0, ke, a_vec);
    if (MODE == 0) load_kmajor<BN>(rb, p.B, p.ldb, n0, p.N, k0, ke, b_vec);
    else load_mnmajor<BN>(rb, p.B, p.ldb, n0, p.N, k0, ke, b_vec);
  };
  auto store_chunk = [&](int s, const OperandRegs& ra, const OperandRegs& rb) {
    uint8_t* st = smem + s * UG_STAGE_BYTES;
    if (MODE == 2) store_mnmajor<128>(ra, st, st + UG_A_PART);
    else store_kmajor<128>(ra, st, st + UG_A_PART);
    if (MODE == 0) store_kmajor<BN>(rb, st + 2 * UG_A_PART, st + 2 * UG_A_PART + UG_B_PART);
    else store_mnmajor<BN>(rb, st + 2 * UG_A_PART, st + 2 * UG_A_PART + UG_B_PART);
  };
  constexpr uint32_t IDESC = umma::idesc_tf32(128, BN);

  // one pipeline step: stage chunk c (already in registers), refill the registers with chunk c+2, hand the stage to the tensor core
  auto step = [&](int c, OperandRegs& ra, OperandRegs& rb) {
    const int s = c % UG_STAGES;
    if (c >= UG_STAGES) umma::mbar_wait(&bar_free[s], (uint32_t)(((c / UG_STAGES) - 1) & 1));   // MMAs of chunk c-3 released stage s
    store_chunk(s, ra, rb);
    if (c + 2 < nchunks) load_chunk(c + 2, ra, rb);      // two chunks in flight in registers
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after_sync();
      const uint32_t base = umma::smem_u32(smem + s * UG_STAGE_BYTES);
      const int kleft = ke - (kb + c * UG_KC);
      const int ksteps = kleft >= UG_KC ? UG_KC / 8 : (kleft + 7) / 8;
      for (int j = 0; j < ksteps; j++) {
        const uint64_t a_hi = umma::smem_desc(base + j * 2 * UG_LBO, UG_LBO, UG_SBO);
        const uint64_t a_lo = umma::smem_desc(base + UG_A_PART + j * 2 * UG_LBO, UG_LBO, UG_SBO);
        const uint64_t b_hi = umma::smem_desc(base + 2 * UG_A_PART + j * 2 * UG_LBO, UG_LBO, UG_SBO);
        const uint64_t b_lo = umma::smem_desc(base + 2 * UG_A_PART + UG_B_PART + j * 2 * UG_LBO, UG_LBO, UG_SBO);
        const uint32_t first = (c > 0 || j > 0) ? 1u : 0u;
        // the large hi*hi term and the 2^-11-scaled corrections accumulate in separate TMEM tiles, so the tensor
        // core's truncating fp32 accumulate touches the main sum once per k-step instead of three times
        umma::mma_tf32(tmem, a_hi, b_hi, IDESC, first);
        umma::mma_tf32(tmem + UG_CORR_COL, a_lo, b_hi, IDESC, first);
        umma::mma_tf32(tmem + UG_CORR_COL, a_hi, b_lo, IDESC, 1u);
      }
      umma::mma_commit(&bar_free[s]);
    }
  };

  OperandRegs ra0, rb0, ra1, rb1;
  if (nchunks > 0) load_chunk(0, ra0, rb0);
  if (nchunks > 1) load_chunk(1, ra1, rb1);
  for (int c = 0; c < nchunks; c += 2) {
    step(c, ra0, rb0);
    if (c + 1 < nchunks) step(c + 1, ra1, rb1);
  }
  if (nchunks > 0) {
    const int last = nchunks - 1;
    umma::mbar_wait(&bar_free[last % UG_STAGES], (uint32_t)((last / UG_STAGES) & 1));
  }
  umma::tc_fence_after_sync();

  // ---- epilogue: thread = output row (TMEM lane), 16 columns per tcgen05.ld ----
  const int row = m0 + warp * 32 + lane;
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const bool c_vec = (p.splits <= 1) && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll 1
  for (int cb = 0; cb < BN; cb += 16) {
    if (n0 + cb >= p.N) break;                      // warp-uniform
    float v[16];
    if (nchunks > 0) {
      float w[16];
      umma::tmem_ld16(taddr + cb, v);
      umma::tmem_ld16(taddr + UG_CORR_COL + cb, w);
#pragma unroll
      for (int q = 0; q < 16; q++) v[q] += w[q];
    } else {
#pragma unroll
      for (int q = 0; q < 16; q++) v[q] = 0.f;
    }
    if (row >= p.M) continue;
    float* crow = p.C + (i64)row * p.ldc + n0 + cb;
    if (p.splits > 1) {
#pragma unroll
      for (int q = 0; q < 16; q++)
        if (n0 + cb + q < p.N) atomicAdd(crow + q, p.alpha * v[q]);
      continue;
    }
#pragma unroll
    for (int q4 = 0; q4 < 16; q4 += 4) {
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int n = n0 + cb + q4 + q;
        float t = p.alpha * v[q4 + q];
        if (n < p.N) {
          if (p.beta != 0.f) t = fmaf(p.beta, crow[q4 + q], t);
          if (p.bias) t += p.bias[n];
        }
        if (p.act == 1) t = fmaxf(t, 0.f);
        o[q] = t;
      }
      if (c_vec && n0 + cb + q4 + 3 < p.N) {
        *reinterpret_cast<float4*>(crow + q4) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (n0 + cb + q4 + q < p.N) crow[q4 + q] = o[q];
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, UG_TMEM_COLS);
}

__global__ void ug_scale2d_kernel(float* C, i64 ldc, int M, int N, float beta) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (i64)M * N) return;
  const int m = (int)(idx / N), n = (int)(idx % N);
  float* c = C + (i64)m * ldc + n;
  *c = (beta == 0.f) ? 0.f : beta * *c;
}

template <int MODE>
static int launch_umma(const UGemmArgs& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    MMDFN_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, UG_SMEM));
    configured = true;
  }
  dim3 grid(ceil_div(p.M, 128), ceil_div(p.N, UG_BN), p.splits > 1 ? p.splits : 1);
  umma_gemm_kernel<MODE><<<grid, UG_THREADS, UG_SMEM, st>>>(p);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

int umma_gemm(bool ta, bool tb, int M, int N, int K, float alpha, const float* A, i64 lda, const float* B, i64 ldb,
              float beta, float* C, i64 ldc, const float* bias, int act, cudaStream_t st) {
  if (M < 0 || N < 0 || K < 0) return MMDFN_EINVAL;
  if (M == 0 || N == 0) return 0;
  if (!A || !B || !C) return MMDFN_ENULL;
  if (tb && ta) return MMDFN_EINVAL;
  UGemmArgs p{A, lda, B, ldb, C, ldc, bias, M, N, K, alpha, beta, act, 1};
  const i64 tiles = (i64)ceil_div(M, 128) * ceil_div(N, UG_BN);
  // split the contraction when the output has fewer tiles than two waves of 2 CTAs/SM (weight gradients)
  if (tiles < 296 && K >= 512 && bias == nullptr && act == 0) {
    i64 s = ceil_div64(592, tiles);
    const i64 smax = ceil_div(K, 128);
    p.splits = (int)(s < smax ? s : smax);
    if (p.splits < 1) p.splits = 1;
  }
  if (p.splits > 1) {
    ug_scale2d_kernel<<<(unsigned)ceil_div64((i64)M * N, 256), 256, 0, st>>>(C, ldc, M, N, beta);
    MMDFN_LAUNCH_CHECK();
  }
  if (!ta && tb) return launch_umma<0>(p, st);
  if (!ta && !tb) return launch_umma<1>(p, st);
  return launch_umma<2>(p, st);
}

}  // namespace mmdfn

extern "C" int mmdfn_gemm_tc(int transA, int transB, int M, int N, int K, float alpha, const float* A, long long lda,
                             const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias,
                             int act, void* stream) {
  return mmdfn::umma_gemm(transA != 0, transB != 0, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act,
                          (cudaStream_t)stream);
}
