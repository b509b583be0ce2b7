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
// second-generation tcgen05 kernel for 16-byte-aligned problems (umma_gemm2.cu); `splits` as chosen by umma_gemm
bool umma_gemm2_eligible(bool ta, bool tb, int M, int N, int K, const float* A, i64 lda, const float* B, i64 ldb);
int umma_gemm2(bool ta, bool tb, int M, int N, int K, float alpha, const float* A, i64 lda, const float* B, i64 ldb,
               float beta, float* C, i64 ldc, const float* bias, int act, int splits, cudaStream_t st);
// third-generation tcgen05 kernel (A operand in tensor memory, one wide CTA per SM; umma_gemm3.cu)
bool umma_gemm3_eligible(bool ta, bool tb, int M, int N, int K, const float* A, i64 lda, const float* B, i64 ldb);
int umma_gemm3_splits(int M, int N, int K, bool plain_epilogue);
void umma_gemm3_set_debug(int v);
void umma_gemm3_set_stamps(long long* device_buf);
int umma_gemm3(bool ta, bool tb, int M, int N, int K, float alpha, const float* A, i64 lda, const float* B, i64 ldb,
               float beta, float* C, i64 ldc, const float* bias, int act, int splits, cudaStream_t st);
int umma_gemm3_nt_pair(int M, int N1, int N2, int K, const float* A, i64 lda, const float* B1, const float* B2, i64 ldb,
                       float* C, i64 ldc, const float* bias1, const float* bias2, cudaStream_t st);
int umma_gemm3_nn_kpair(int M, int N, int K1, int K2, const float* A, i64 lda, const float* B1, const float* B2, i64 ldb,
                        float beta, float* C, i64 ldc, cudaStream_t st);
int umma_gemm3_npair(bool ta, int M, int N, int K, const float* A, i64 lda, const float* B1, const float* B2, i64 ldb,
                     float beta1, float* C1, float beta2, float* C2, i64 ldc, int splits, cudaStream_t st);
//   gemm_npair:     C1 = op(A) B1 + beta1 C1 ; C2 = op(A) B2 + beta2 C2   (NN or TN form, N columns each)
int gemm_npair(bool ta, int M, int N, int K, const float* A, i64 lda, const float* B1, const float* B2, i64 ldb, float beta1,
               float* C1, float beta2, float* C2, i64 ldc, cudaStream_t st);
// operand pairs: one launch on the third-generation tensor-core kernel where it applies, two plain gemm() calls otherwise
//   gemm_nt_pair:   C[:, :N1] = A B1^T + bias1 ; C[:, N1:N1+N2] = A B2^T + bias2
//   gemm_nn_kpair:  C = A[:, :K1] B1 + A[:, K1:K1+K2] B2 + beta C
int gemm_nt_pair(int M, int N1, int N2, int K, const float* A, i64 lda, const float* B1, const float* B2, i64 ldb, float* C,
                 i64 ldc, const float* bias1, const float* bias2, cudaStream_t st);
int gemm_nn_kpair(int M, int N, int K1, int K2, const float* A, i64 lda, const float* B1, const float* B2, i64 ldb, float beta,
                  float* C, i64 ldc, cudaStream_t st);
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


// ---- fused graph-conv layer (gcn_layer.cu): aggregate -> x folded weight -> fused epilogue, one launch per layer ----
long long gcn_layer_img_floats();
// folded matrices Mtop_l / Mbot_l (column blocks of two (100, 100 K) matrices) and the pre-split phase-B operand images
int gcn_layer_prep(int K, const float* const* convW, double lamda, double alpha, float* mtop_all, float* mbot_all,
                   float* img_f, float* img_b, cudaStream_t st);
// dconvW[l] (=, or += when accumulate) theta_l [dMtop_l ; dMbot_l]
int gcn_layer_unfold(int K, float* const* dconvW, double lamda, const float* dmtop_all, const float* dmbot_all,
                     int accumulate, cudaStream_t st);
// out = dropout(relu((A_hat zin) Mtop + r)) (+ q); flags = [relu and keep]
int gcn_layer_fwd(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* adj_blk,
                  const float* adj_diag, const float* zin, const float* wimg, const float* r, i64 ldr, const float* q,
                  const unsigned char* mask, float scale, unsigned char* flags, float* out, i64 ldo, cudaStream_t st);
// t_out = A_hat du ; out = t_out Mtop^T (+ add)
int gcn_layer_bwd(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* adj_blk,
                  const float* adj_diag, const float* du, i64 ldu, const float* wimg_t, float* t_out, i64 ldt,
                  const float* add, float* out, cudaStream_t st);

}  // namespace mmdfn
