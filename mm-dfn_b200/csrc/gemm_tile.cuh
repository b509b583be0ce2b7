// FP32 (FFMA) register-tiled GEMM main loop, shared by the dense GEMM entry point and by
// the grouped per-dialogue graph kernels (adjacency Gram, message aggregate, dA_hat).
// Exact fp32 accumulate: every contraction that feeds acos() or the 1e-4 logit budget
// stays in true fp32.
#pragma once
#include "common.cuh"

namespace mmdfn {

constexpr int GEMM_THREADS = 256;
constexpr int GEMM_BK = 16;

template <int BM, int BN>
struct GemmSmem {
  static constexpr int LDA = BM + 4;
  static constexpr int LDB = BN + 4;
  static constexpr int A_STAGE = GEMM_BK * LDA;
  static constexpr int B_STAGE = GEMM_BK * LDB;
  static constexpr int FLOATS = 2 * (A_STAGE + B_STAGE);
};

// Row / column owned by accumulator slot i / j of this thread.  8-wide slots are split in
// two groups of 4 half a tile apart so that shared loads stay conflict-free 128-bit and
// global stores stay coalesced.
template <int BM, int TM>
__device__ __forceinline__ int tile_row(int ty, int i) {
  return (TM == 8) ? ((i >> 2) * (BM / 2) + ty * 4 + (i & 3)) : (ty * TM + i);
}
template <int BN, int TN>
__device__ __forceinline__ int tile_col(int tx, int j) {
  return (TN == 8) ? ((j >> 2) * (BN / 2) + tx * 4 + (j & 3)) : (tx * TN + j);
}

// acc += op(A)[m0:m0+BM, k_begin:k_end] * op(B)[k_begin:k_end, n0:n0+BN]
//   op(A)(m,k) = TA ? A[k*lda + m] : A[m*lda + k]     (rows >= M read as 0)
//   op(B)(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]     (cols >= N read as 0)
// All 256 threads of the CTA must call it; ends with a __syncthreads().
template <int BM, int BN, int TM, int TN, bool TA, bool TB>
__device__ __forceinline__ void gemm_tile_accum(const float* __restrict__ A, i64 lda,
                                                const float* __restrict__ B, i64 ldb, int M, int N, int m0,
                                                int n0, int k_begin, int k_end, float (&acc)[TM][TN],
                                                float* smem) {
  static_assert((BM / TM) * (BN / TN) == GEMM_THREADS, "tile/thread shape");
  static_assert(TM == 4 || TM == 8, "TM");
  static_assert(TN == 4 || TN == 8, "TN");
  using S = GemmSmem<BM, BN>;
  constexpr int LA = BM * GEMM_BK / GEMM_THREADS;
  constexpr int LB = BN * GEMM_BK / GEMM_THREADS;
  float* As = smem;
  float* Bs = smem + 2 * S::A_STAGE;
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  float ra[LA], rb[LB];

  auto a_ml = [&](int i) { return TA ? (tid % BM) : (tid / GEMM_BK + (GEMM_THREADS / GEMM_BK) * i); };
  auto a_kl = [&](int i) { return TA ? (tid / BM + (GEMM_THREADS / BM) * i) : (tid % GEMM_BK); };
  auto b_nl = [&](int i) { return TB ? (tid / GEMM_BK + (GEMM_THREADS / GEMM_BK) * i) : (tid % BN); };
  auto b_kl = [&](int i) { return TB ? (tid % GEMM_BK) : (tid / BN + (GEMM_THREADS / BN) * i); };

  auto load_g = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; i++) {
      const int m = m0 + a_ml(i), k = k0 + a_kl(i);
      float v = 0.f;
      if (m < M && k < k_end) v = TA ? A[(i64)k * lda + m] : A[(i64)m * lda + k];
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < LB; i++) {
      const int n = n0 + b_nl(i), k = k0 + b_kl(i);
      float v = 0.f;
      if (n < N && k < k_end) v = TB ? B[(i64)n * ldb + k] : B[(i64)k * ldb + n];
      rb[i] = v;
    }
  };
  auto store_s = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; i++) As[buf * S::A_STAGE + a_kl(i) * S::LDA + a_ml(i)] = ra[i];
#pragma unroll
    for (int i = 0; i < LB; i++) Bs[buf * S::B_STAGE + b_kl(i) * S::LDB + b_nl(i)] = rb[i];
  };

  const int nk = (k_end - k_begin + GEMM_BK - 1) / GEMM_BK;
  if (nk <= 0) {
    __syncthreads();
    return;
  }
  load_g(k_begin);
  store_s(0);
  __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_g(k_begin + (kt + 1) * GEMM_BK);
    const float* as = As + buf * S::A_STAGE;
    const float* bs = Bs + buf * S::B_STAGE;
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; kk++) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(as + kk * S::LDA + tile_row<BM, TM>(ty, i));
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(bs + kk * S::LDB + tile_col<BN, TN>(tx, j));
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_s(buf ^ 1);
    __syncthreads();
  }
}

template <int TM, int TN>
__device__ __forceinline__ void zero_acc(float (&acc)[TM][TN]) {
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;
}

}  // namespace mmdfn
