// Host-side launchers shared between translation units (internal C++ API; the public
// surface is include/mmdfn_b200.h).  Every launcher only enqueues work on `st`.
#pragma once
#include "common.cuh"

namespace mmdfn {

// C[M,N] = act(alpha * op(A) op(B) + beta * C + bias[N]); row-major; see mmdfn_gemm.
int gemm(bool ta, bool tb, int M, int N, int K, float alpha, const float* A, i64 lda, const float* B, i64 ldb,
         float beta, float* C, i64 ldc, const float* bias, int act, cudaStream_t st);
// same contract on tcgen05 (3xTF32)
int umma_gemm(bool ta, bool tb, int M, int N, int K, float alpha, const float* A, i64 lda, const float* B, i64 ldb,
              float beta, float* C, i64 ldc, const float* bias, int act, cudaStream_t st);
// out[n] = beta*out[n] + sum_m A[m*lda + n]
int colsum(int M, int N, const float* A, i64 lda, float beta, float* out, cudaStream_t st);
int fill_zero(void* p, size_t bytes, cudaStream_t st);

// y = A_hat x on the block-compact adjacency (N3 = 3N rows, G columns)
int adj_spmm(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* adj_blk,
             const float* adj_diag, const float* x, int G, float* y, cudaStream_t st);
// the same product on tcgen05 (3xTF32) for G == 100, every block <= 128 rows, 16-byte aligned x / y (spmm_tc.cu)
int adj_spmm_tc(int B, int N, const int* dia_off, const i64* blk_off, const float* adj_blk, const float* adj_diag,
                const float* x, float* y, cudaStream_t st);
// experimental: the same on tcgen05 for any dialogue length (128-row tiles, streamed contraction; spmm_tc_long.cu)
int adj_spmm_tc_long(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* adj_blk,
                     const float* adj_diag, const float* x, float* y, cudaStream_t st);
// P_blk (+)= sym(dhi z^T) ; P_diag (+)= sym cross terms
int adj_grad_accum(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* dhi,
                   const float* z, int G, float* p_blk, float* p_diag, int accumulate, cudaStream_t st);

}  // namespace mmdfn
